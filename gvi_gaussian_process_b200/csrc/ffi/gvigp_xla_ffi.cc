// gvigp_xla_ffi.cc - XLA FFI (typed custom-call) handlers over the C ABI of include/gvigp.h, so that the reference's own
// host (JAX) can call the B200 kernels from inside jit-compiled code: BASELINE.json north_star "a thin C-ABI layer
// registered as JAX FFI custom calls".
//
// STATUS: written against the public xla/ffi/api/ffi.h of jaxlib >= 0.4.31; NOT compiled or run in this repository's
// image (no jax / jaxlib, hence no XLA headers here).  tests/test_abi.py type-checks it against tests/xla_ffi_stub (a stand-in for
// the part of that header used here): forwarded calls vs include/gvigp.h, handler signatures vs their bindings.  gvi_gaussian_process_b200/build.py:build_ffi() compiles it into
// libgvigp_ffi.so when a jaxlib with the header tree is importable, and gvi_gaussian_process_b200/jax_ffi.py registers
// the handlers and wraps them in the reference's method signatures.  Every handler is a plain forwarding call: buffers
// arrive as device pointers on XLA's stream, which is exactly what the C ABI takes; workspaces and the factor state are
// extra result buffers sized by the *_bytes queries on the Python side; scalar options are FFI attributes.
#include <cstdint>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

#include "../../../include/gvigp.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;
using S64Out = ffi::ResultBuffer<ffi::S64>;

namespace {
inline ffi::Error status(int rc) {
  return rc == GVI_OK ? ffi::Error::Success() : ffi::Error::Internal(gvi_last_error_string());
}
inline int64_t rows(const F64& b) { return b.dimensions().size() ? b.dimensions()[0] : 1; }
inline int cols(const F64& b) { return b.dimensions().size() > 1 ? (int)b.dimensions()[1] : 1; }

// ARDKernel._calculate_gram (src/kernels/standard/base.py:73-103, ard_kernel.py:58-66)
ffi::Error ArdGram(cudaStream_t stream, F64 x1, F64 x2, F64 log_scaling, F64 log_ls, F64Out out, F64Out workspace) {
  return status(gvi_ard_gram_f64(x1.typed_data(), rows(x1), x2.typed_data(), rows(x2), cols(x1), log_scaling.typed_data(),
                                 log_ls.typed_data(), (int)log_ls.element_count(), out->typed_data(), rows(x2),
                                 workspace->element_count() ? workspace->typed_data() : nullptr, stream));
}

// InnerProductKernel / PolynomialKernel Gram (non_stationary/inner_product_kernel.py:36-42, polynomial_kernel.py:40-52)
ffi::Error DotGram(cudaStream_t stream, F64 x1, F64 x2, F64 log_scaling, F64 log_constant, double degree,
                   bool has_constant, F64Out out) {
  return status(gvi_dot_gram_f64(x1.typed_data(), rows(x1), x2.typed_data(), rows(x2), cols(x1), log_scaling.typed_data(),
                                 has_constant ? log_constant.typed_data() : nullptr, degree, out->typed_data(), rows(x2),
                                 stream));
}

// cho_factor state of one GP over the inducing set (sparse_posterior_kernel.py:63-74, src/gps/base/base.py:516-521)
ffi::Error GpFactor(cudaStream_t stream, F64 z, F64 log_scaling, F64 log_ls, F64 log_noise, F64 y_z, double reg,
                    bool is_absolute, bool exact_gp, F64Out factor, F64Out workspace) {
  return status(gvi_gp_factor_f64(z.typed_data(), (int)rows(z), cols(z), log_scaling.typed_data(), log_ls.typed_data(),
                                  (int)log_ls.element_count(), reg, is_absolute ? 1 : 0,
                                  exact_gp ? log_noise.typed_data() : nullptr, exact_gp ? y_z.typed_data() : nullptr,
                                  factor->typed_data(), workspace->typed_data(), nullptr, stream));
}

// predictive mean / variance (sparse_posterior_kernel.py:90-96 in diagonal mode; base.py:519-525)
ffi::Error GpPredict(cudaStream_t stream, F64 factor, F64 x, F64 prior_mean, int64_t m, bool has_prior_mean,
                     F64Out mean, F64Out var, F64Out workspace) {
  return status(gvi_gp_predict_f64(factor.typed_data(), (int)m, cols(x), x.typed_data(), rows(x),
                                   has_prior_mean ? prior_mean.typed_data() : nullptr, nullptr, mean->typed_data(),
                                   var->typed_data(), workspace->typed_data(), stream));
}

// NegativeLogLikelihood + pointwise regularisers as sums (negative_log_likelihood.py:41-55, regularisations/projected/*)
ffi::Error Risk(cudaStream_t stream, F64 y, F64 m_q, F64 v_q, F64 m_p, F64 v_p, int64_t reg_kind, double alpha,
                F64Out out3, F64Out workspace) {
  return status(gvi_risk_f64((int)reg_kind, alpha, y.typed_data(), m_q.typed_data(), v_q.typed_data(),
                             reg_kind ? m_p.typed_data() : nullptr, reg_kind ? v_p.typed_data() : nullptr,
                             (int64_t)m_q.element_count(), out3->typed_data(), workspace->typed_data(), stream));
}

// GeneralisedVariationalInference._calculate_loss in one launch (generalised_variational_inference.py:73-94)
ffi::Error FusedStep(cudaStream_t stream, F64 factor_q, F64 factor_p, F64 x, F64 y, F64 m_q, F64 prior_mean_p,
                     int64_t m, int64_t reg_kind, double alpha, F64Out out3, F64Out workspace) {
  return status(gvi_fused_step_f64(factor_q.typed_data(), factor_p.typed_data(), (int)m, cols(x), x.typed_data(),
                                   y.typed_data(), m_q.typed_data(), prior_mean_p.typed_data(), nullptr, nullptr, rows(x),
                                   (int)reg_kind, alpha, out3->typed_data(), nullptr, nullptr, nullptr,
                                   workspace->typed_data(), stream));
}

// its gradient (jax.grad of calculate_loss, experiments/shared/trainer.py:83-89): the bwd of the custom_vjp
ffi::Error FusedStepGrad(cudaStream_t stream, F64 factor_q, F64 factor_p, F64 x, F64 y, F64 m_q, F64 prior_mean_p,
                         int64_t m, int64_t reg_kind, double alpha, F64Out out, F64Out dm, F64Out workspace) {
  return status(gvi_fused_step_grad_f64(factor_q.typed_data(), factor_p.typed_data(), (int)m, cols(x), x.typed_data(),
                                        y.typed_data(), m_q.typed_data(), prior_mean_p.typed_data(), nullptr, nullptr,
                                        rows(x), (int)reg_kind, alpha, 0, out->typed_data(), dm->typed_data(), nullptr,
                                        nullptr, nullptr, workspace->typed_data(), stream));
}

// ConditionalVarianceInducingPointsSelector.compute_inducing_points on permuted inputs (inducing_points_selection.py:113-176)
ffi::Error CondVarSelect(cudaStream_t stream, F64 x_perm, F64 log_scaling, F64 log_ls, int64_t m_sel, double jitter,
                         S64Out idx, F64Out trace, ffi::ResultBuffer<ffi::U8> workspace) {
  return status(gvi_cond_var_select_f64(x_perm.typed_data(), rows(x_perm), cols(x_perm), (int)m_sel,
                                        log_scaling.typed_data(), log_ls.typed_data(), (int)log_ls.element_count(), jitter,
                                        idx->typed_data(), trace->typed_data(), workspace->typed_data(), stream));
}
}  // namespace

#define GVI_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(gvi_ard_gram_ffi, ArdGram,
                              GVI_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(gvi_dot_gram_ffi, DotGram,
                              GVI_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Attr<double>("degree")
                                  .Attr<bool>("has_constant").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(gvi_gp_factor_ffi, GpFactor,
                              GVI_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Attr<double>("diagonal_regularisation").Attr<bool>("is_absolute")
                                  .Attr<bool>("exact_gp").Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(gvi_gp_predict_ffi, GpPredict,
                              GVI_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Attr<int64_t>("m").Attr<bool>("has_prior_mean")
                                  .Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(gvi_risk_ffi, Risk,
                              GVI_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Attr<int64_t>("reg_kind")
                                  .Attr<double>("alpha").Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(gvi_fused_step_ffi, FusedStep,
                              GVI_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Attr<int64_t>("m").Attr<int64_t>("reg_kind").Attr<double>("alpha").Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(gvi_fused_step_grad_ffi, FusedStepGrad,
                              GVI_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Attr<int64_t>("m").Attr<int64_t>("reg_kind").Attr<double>("alpha").Ret<F64>().Ret<F64>()
                                  .Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(gvi_cond_var_select_ffi, CondVarSelect,
                              GVI_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Attr<int64_t>("m_sel").Attr<double>("jitter")
                                  .Ret<ffi::Buffer<ffi::S64>>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());
