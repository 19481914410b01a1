/*
 * gvigp.h - C ABI of libgvigp.so: the B200 (sm_100a) implementation of the data-parallel GP hot path of
 * jswu18/gvi-gaussian-process.  The reference has no native/FFI layer (it is pure Python on JAX); each entry
 * point below names the reference *Python* override point it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - every function returns 0 on success and a negative GVI_E* code on failure; gvi_last_error_string()
 *     describes the last failure of the calling thread.  Nothing throws across the ABI.
 *   - all pointers are DEVICE pointers unless the name ends in _h; all arrays are float64, C-contiguous,
 *     row-major; indices are int64.
 *   - kernel hyper-parameters (log_scaling, log_lengthscales, log_observation_noise) are passed as device
 *     pointers so a training loop never has to synchronise to read them.
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*); no hidden synchronisation
 *     and no allocation: the caller owns outputs and workspaces (sizes from the *_bytes queries).
 */
#ifndef GVIGP_H
#define GVIGP_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define GVI_OK 0
#define GVI_EINVAL (-1)   /* bad argument (shape/size/null) - the Python shim raises AssertionError first */
#define GVI_ECUDA (-2)    /* CUDA runtime error, see gvi_last_error_string() */
#define GVI_EUNSUPPORTED (-3) /* size outside what this build supports */
#define GVI_ECOMM (-4)    /* communicator / NCCL error, see gvi_last_error_string() */

/* regulariser kinds for gvi_risk_f64 / gvi_fused_step_f64 (src/regularisations/...) */
#define GVI_REG_NONE 0
#define GVI_REG_PROJ_BHATTACHARYYA 1 /* projected/projected_bhattacharyya_regularisation.py:34-36 */
#define GVI_REG_PROJ_GAUSSIAN_WASSERSTEIN 2 /* projected/projected_gaussian_wasserstein_regularisation.py:34 */
#define GVI_REG_PROJ_HELLINGER 3 /* projected/projected_hellinger_regularisation.py:34-45 */
#define GVI_REG_PROJ_KL 4 /* projected/projected_kl_regularisation.py:34-38 */
#define GVI_REG_PROJ_RENYI 5 /* projected/projected_renyi_regularisation.py:36-45 */
#define GVI_REG_GWI_NO_EIG 6 /* gaussian_wasserstein_regularisation.py:143-160 (include_eigendecomposition=False) */
#define GVI_REG_SQUARED_DIFFERENCE 7 /* gaussian_squared_difference_regularisation.py:48-86 (full_covariance=False) */

int gvi_abi_version(void);
const char* gvi_last_error_string(void);
/* number of CUDA kernels this library has launched since it was loaded (bench.py's gpu_launches) */
int64_t gvi_launch_count(void);

/* ---- a1-a3: ARDKernel._calculate_gram (src/kernels/standard/base.py:73-103, ard_kernel.py:58-66) ----------
 * out[i*ld_out + j] = exp(log_scaling)^2 * exp(-0.5 * sum_d exp(log_ls[d]) * (x1[i,d]-x2[j,d])^2).
 * ls_len is D, or 1 (a scalar lengthscale broadcast over D, README.md:92-94). */
/* D >= 24 with a workspace: the x1.x2^T contraction runs on the FP64 tensor pipe (DMMA) with the squared-distance,
 * lengthscale and exp epilogue fused; otherwise (workspace NULL or small D) direct differences.
 * 5 <= D < 24 with a workspace (256 bytes; the size query returns 0 for outputs too small to pay for it): the same
 * expansion on the tensor pipe BEHIND A DEVICE-SIDE GUARD - a pass over the inputs bounds (max |u| + max |v|)^2 of the
 * scaled, centred rows; within 4096 (error of k <= 1.2e-11 relative) the tensor-pipe kernel writes the result, else
 * (short lengthscales, far-apart clusters, inf / NaN inputs) the direct-difference kernel does; no host sync.       */
size_t gvi_ard_gram_workspace_bytes(int64_t m1, int64_t m2, int D);
int gvi_ard_gram_f64(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                     const double* log_scaling, const double* log_ls, int ls_len,
                     double* out, int64_t ld_out, void* workspace, void* stream);
/* full_covariance=False path of KernelBase.calculate_gram (src/kernels/base.py:132-147): out[i] = k(x_i,x_i). */
int gvi_ard_gram_diag_f64(int64_t n, const double* log_scaling, double* out, void* stream);

/* ---- other kernels.calculate_gram: InnerProductKernel / PolynomialKernel through StandardKernelBase._calculate_gram
 * (src/kernels/non_stationary/inner_product_kernel.py:36-42, polynomial_kernel.py:40-52, standard/base.py:73-103):
 * out[i*ld_out + j] = (exp(log_scaling) * <x1_i, x2_j> + exp(log_constant))^degree.  log_constant == NULL drops the
 * constant (inner-product kernel: degree 1).  Integral degrees are evaluated by repeated multiplication (what
 * jnp.power does for a Python-int exponent), other degrees by pow. */
int gvi_dot_gram_f64(const double* x1, int64_t m1, const double* x2, int64_t m2, int D, const double* log_scaling,
                     const double* log_constant, double degree, double* out, int64_t ld_out, void* stream);
/* full_covariance=False path of KernelBase.calculate_gram for two DIFFERENT inputs (src/kernels/base.py:132-147):
 * out[i] = k(x1_i, x2_i).  kind GVI_KERNEL_ARD: p0 = log_scaling, p1 = log_lengthscales[p1_len] (degree ignored);
 * GVI_KERNEL_POLYNOMIAL: p0 = log_scaling, p1 = log_constant or NULL (p1_len ignored). */
#define GVI_KERNEL_ARD 0
#define GVI_KERNEL_POLYNOMIAL 1
int gvi_kernel_pairs_f64(int kind, const double* x1, const double* x2, int64_t n, int D, const double* p0,
                         const double* p1, int p1_len, double degree, double* out, void* stream);

/* ---- a4: add_diagonal_regulariser (src/utils/matrix_operations.py:7-28), in place on A[M,M] (ld doubles) -- */
int gvi_add_diagonal_regulariser_f64(double* A, int M, int64_t ld, double diagonal_regularisation,
                                     int is_absolute, void* stream);
/* ---- cho_factor call sites (sparse_posterior_kernel.py:68-74, src/gps/base/base.py:516-518): blocked
 * lower Cholesky in place (strict upper triangle left untouched); *info = 0 or 1+first non-PD column. ------- */
int gvi_chol_factor_f64(double* A, int M, int64_t ld, int* info, void* stream);

/* ---- a5/a7: factorisation state of one GP over the inducing set Z ------------------------------------------
 * Builds A = K_ZZ + jitter*I (+ exp(log_noise) I), its Cholesky L, L^-1 packed in DMMA-fragment order, and
 * alpha = A^-1 y_Z.  Replaces the per-call cho_factor of SparsePosteriorKernel._calculate_gram
 * (sparse_posterior_kernel.py:63-74) and GPBase.calculate_posterior_matrices_of_kernel (base.py:516-521).
 *   diagonal_regularisation/is_absolute : a4 semantics (pass 0.0 for the exact GP)
 *   log_noise : device scalar or NULL (adds exp(log_noise) to the diagonal; exact GP, base.py:473-476)
 *   y_z       : [M] or NULL (alpha = 0)                                                                     */
size_t gvi_gp_factor_bytes(int M, int D);
size_t gvi_gp_factor_workspace_bytes(int M);
int gvi_gp_factor_f64(const double* Z, int M, int D, const double* log_scaling, const double* log_ls,
                      int ls_len, double diagonal_regularisation, int is_absolute, const double* log_noise,
                      const double* y_z, void* factor_out, void* workspace, int* info, void* stream);

/* Same state with a FROZEN Cholesky: K(x,Z) tiles use (log_scaling, log_ls) but A = K_ZZ(chol_*) + jitter is built
 * from other, fixed parameters - FixedSparsePosteriorKernel (fixed_sparse_posterior_kernel.py:41-52, 78-99) and the
 * SVGP family (svgp/base.py:57-73).                                                                           */
int gvi_gp_factor_frozen_f64(const double* Z, int M, int D, const double* log_scaling, const double* log_ls,
                             int ls_len, const double* chol_log_scaling, const double* chol_log_ls, int chol_ls_len,
                             double diagonal_regularisation, int is_absolute, void* factor_out, void* workspace,
                             int* info, void* stream);
/* SVGP family: attach (E != NULL) or detach a lower-triangular matrix E [M,M] (row stride ld) to the state; the
 * predictive variance becomes k(x,x) - ||L^-1 k||^2 + ||E k||^2, i.e. + k(x,Z) (E^T E) k(Z,x)
 * (svgp/cholesky_svgp_kernel.py:133-150 with E = tril(P,-1) + diag(exp(d)); diagonal_svgp_kernel.py:117-129 with
 * E = diag(exp(d/2))).  The second triangular product runs in the same tile kernel on the same K tile.        */
int gvi_gp_factor_set_extra_f64(void* factor, int M, int D, const double* E, int64_t ld, void* stream);

/* ---- a5-a7: predictive mean / variance without any N x M intermediate in HBM -------------------------------
 * var[i]  = k(x_i,x_i) - || L^-1 k(Z,x_i) ||^2 (+ exp(log_noise_add) if given)   (sparse_posterior_kernel.py:90-96
 *           in diagonal mode; base.py:522-525 for the exact GP; base.py:205-211 for the added noise)
 * mean[i] = k(x_i,Z) alpha + prior_mean[i]                                    (base.py:519-521, 492)
 * mean_out / prior_mean / log_noise_add may be NULL; workspace may be NULL when its size query returns 0.
 * Any M: for M > 1152 the K tile is swept in 1024-column panels and the partial triangular products are parked
 * in the (L2-resident) workspace between sweeps - still no N x M array in HBM.  Any D: for D > 32 the rows
 * K(x_chunk, Z) are produced <= 28 416 points at a time (whole waves of both kernels) by the tensor-pipe Gram kernel
 * into a ring inside the workspace (<= 512 MB) and consumed by the tile kernel; gvi_fused_step_f64 does the same
 * with two rings, the gradient entry points take the GEMM route of grad_wide.cu for D > 32.
 * gvi_gp_predict_workspace_bytes(M, D) is enough for any N; the _n variant sizes the ring for one call's N.    */
size_t gvi_gp_predict_workspace_bytes(int M, int D);
size_t gvi_gp_predict_workspace_bytes_n(int M, int D, int64_t N);
int gvi_gp_predict_f64(const void* factor, int M, int D, const double* X, int64_t N,
                       const double* prior_mean, const double* log_noise_add,
                       double* mean_out, double* var_out, void* workspace, void* stream);

/* ---- a5-a7 over ANY base kernel: the caller supplies the Grams -------------------------------------------------------
 * SparsePosteriorKernel / the exact GP posterior are written against KernelBase.calculate_gram
 * (sparse_posterior_kernel.py:63-96, src/gps/base/base.py:495-526), so they work over every kernel of the reference
 * (polynomial, inner product, CustomKernel = the NNGP wrapper, CustomMappingKernel).  For kernels other than ARD the host
 * side materialises K_ZZ once and K(x_chunk, Z) chunk by chunk with that kernel's own Gram routine and hands them over:
 *   gvi_gp_factor_gram_f64 : same factor state as gvi_gp_factor_f64 from Kzz[M,M] (ld doubles; only read), D = 1 layout
 *                            (gvi_gp_factor_bytes(M, 1)); jitter / noise / y_z as in gvi_gp_factor_f64.
 *   gvi_gp_predict_gram_f64: var[i] = kdiag[i] - ||L^-1 Kxz[i,:]||^2 (+ exp(log_noise_add)), mean[i] = Kxz[i,:] alpha +
 *                            prior_mean[i] on the tensor pipe (the tile kernel reads its K tile from Kxz[N,M], ld
 *                            doubles, instead of evaluating the kernel).  kdiag[N] = k(x_i, x_i).
 *                            Workspace: gvi_gp_predict_workspace_bytes(M, 1). */
int gvi_gp_factor_gram_f64(const double* Kzz, int64_t ld, int M, double diagonal_regularisation, int is_absolute,
                           const double* log_noise, const double* y_z, void* factor_out, void* workspace, int* info,
                           void* stream);
int gvi_gp_predict_gram_f64(const void* factor, int M, const double* Kxz, int64_t ld, const double* kdiag, int64_t N,
                            const double* prior_mean, const double* log_noise_add, double* mean_out, double* var_out,
                            void* workspace, void* stream);

/* ---- a9-a11: risk and regulariser sums ----------------------------------------------------------------------
 * out[0] = sum_i NLL_i  (negative_log_likelihood.py:41-55), out[1] = sum_i D0_i (regulariser `kind`),
 * out[2] = N as double.  Sums, not means, so shards can be all-reduced; deterministic two-stage tree.
 * workspace: gvi_risk_workspace_bytes(N).  y/m_p/v_p may be NULL when the corresponding term is unused.     */
size_t gvi_risk_workspace_bytes(int64_t N);
int gvi_risk_f64(int reg_kind, double renyi_alpha, const double* y, const double* m_q, const double* v_q,
                 const double* m_p, const double* v_p, int64_t N, double* out3, void* workspace, void* stream);

/* Pointwise derivatives of the sums of gvi_risk_f64 w.r.t. the approximate GP's mean and variance (p constant):
 * dm_out[i] = w[0] dNLL_i/dm_q,i + w[1] dD0_i/dm_q,i and dv_out[i] likewise for v_q,i - the cotangents that
 * jax.grad of the reference's risks feeds back into predict_probability.  weights2 is a DEVICE array [2] (the
 * cotangents of out3[0], out3[1]) so a training step never synchronises.                                       */
int gvi_risk_grad_f64(int reg_kind, double renyi_alpha, const double* y, const double* m_q, const double* v_q,
                      const double* m_p, const double* v_p, int64_t N, const double* weights2, double* dm_out,
                      double* dv_out, void* stream);

/* ---- a12: fused Gram + predict(q) + predict(p) + NLL + regulariser for one shard of points ---------------
 * (GeneralisedVariationalInference._calculate_loss, src/generalised_variational_inference.py:73-94.)
 * One launch: per tile of points builds K(x,Z) for q and p in shared memory, runs both triangular products on
 * the FP64 tensor pipe and reduces the pointwise loss terms; only O(N) data crosses HBM.
 * factor_p may be NULL with reg_kind == GVI_REG_NONE.  prior-mode regulariser: pass factor_p = NULL and
 * m_p/v_p arrays instead.  out3 as in gvi_risk_f64.  var_q_out/mean_p_out/var_p_out optional (may be NULL). */
size_t gvi_fused_step_workspace_bytes(int64_t N, int M);
/* D > 32: the K rows of q and p are built <= 28 416 points at a time by the tensor-pipe Gram kernel into two rings
 * which the tile kernel consumes (2-3 launches per chunk inside the one call; no N x M array in HBM).  Such calls
 * need the larger workspace of gvi_fused_step_workspace_bytes_d (== the call above for D <= 32).                    */
size_t gvi_fused_step_workspace_bytes_d(int64_t N, int M, int D);
int gvi_fused_step_f64(const void* factor_q, const void* factor_p, int M, int D, const double* X,
                       const double* y, const double* m_q, const double* prior_mean_p,
                       const double* m_p_in, const double* v_p_in, int64_t N, int reg_kind, double renyi_alpha,
                       double* out3, double* var_q_out, double* mean_p_out, double* var_p_out,
                       void* workspace, void* stream);

/* ---- gradient of a12 (jax.grad of calculate_loss, experiments/shared/trainer.py:83-89) ---------------------------
 * Same inputs as gvi_fused_step_f64.  out[0..2] as out3; out[3] = d(sum_i loss_i)/dlog_scaling of q;
 * out[4+d] = d(sum_i loss_i)/dlog_lengthscales[d] of q (d < D); dm_out[i] = d(sum loss)/dm_q[i].  All in SUM form
 * (divide by the global N after the all-reduce); the regulariser GP p is constant.  The point-sized work
 * (a_i = L^-T L^-1 k_i per tile, S = sum_i g_i a_i a_i^T accumulated from the a_i tiles chunk by chunk) runs on the
 * FP64 tensor pipe.  out has 4 + D doubles.                                                                     */
size_t gvi_fused_step_grad_workspace_bytes(int64_t N, int M, int D);
int gvi_fused_step_grad_f64(const void* factor_q, const void* factor_p, int M, int D, const double* X,
                            const double* y, const double* m_q, const double* prior_mean_p, const double* m_p_in,
                            const double* v_p_in, int64_t N, int reg_kind, double renyi_alpha, int frozen_cholesky,
                            double* out, double* dm_out, double* var_q_out, double* mean_p_out, double* var_p_out,
                            void* workspace, void* stream);
/* frozen_cholesky = 1 for a state built with gvi_gp_factor_frozen_f64 (FixedSparsePosteriorKernel): A does not depend
 * on the trained parameters, so the <dK_ZZ, S> term (and the weighted-Gram pass) is skipped.
 *
 * SVGP family (state with an extra matrix E attached, kernel hyper-parameters frozen): value, dR/dm_q[N] and
 * dR/dE = 2 tril(E C), C = sum_i g_i k_i k_i^T from the same weighted-Gram kernel; dE_out is [M,M] row-major.
 * Workspace: gvi_fused_step_grad_workspace_bytes.                                                              */
int gvi_svgp_step_grad_f64(const void* factor_q, const void* factor_p, int M, int D, const double* X,
                           const double* y, const double* m_q, const double* prior_mean_p, const double* m_p_in,
                           const double* v_p_in, int64_t N, int reg_kind, double renyi_alpha, const double* E,
                           int64_t lde, double* out3, double* dm_out, double* dE_out, void* workspace, void* stream);

/* Vector-Jacobian product of gvi_gp_predict_f64's variance w.r.t. the kernel hyper-parameters: given
 * g_var[i] = dLoss/dvar_i of ANY downstream loss, out[3] = dLoss/dlog_scaling, out[4+d] = dLoss/dlog_lengthscales[d]
 * (out[0] = out[1] = 0, out[2] = N; same layout and workspace as gvi_fused_step_grad_f64).  This is what jax.grad
 * produces through SparsePosteriorKernel._calculate_gram (sparse_posterior_kernel.py:54-96) for losses that are not
 * the fused regression objective - e.g. the classification objective of section 8(f)-2.  D <= 32 runs in the tile
 * kernel; wider inputs (config #4: D = 784) take a GEMM formulation of the same contraction on the tensor pipe.  */
int gvi_gp_predict_grad_f64(const void* factor, int M, int D, const double* X, int64_t N, const double* g_var,
                            int frozen_cholesky, double* out, void* workspace, void* stream);

/* ---- section 8(f)-3: vector-Jacobian product of the EXACT GP predictive (mean AND variance) - the backward pass of
 * the exact-GP NLL training step on the inducing set (experiments/regression/trainers.py:19-85: jax.grad through
 * GPRegression.predict_probability, src/gps/base/base.py:384-526).  `factor` is a state built by gvi_gp_factor_f64
 * with y_z and log_noise (A = K_ZZ + exp(log_noise) I, alpha = A^-1 y_Z).  Given g_var[i] = dLoss/dvar_i and
 * h_mean[i] = dLoss/dmean_i (NULL = 0):  out[0] = out[1] = 0, out[2] = N, out[3] = dLoss/dlog_scaling,
 * out[4+d] = dLoss/dlog_lengthscales[d], out[4+D] = dLoss/dlog_observation_noise (5 + D doubles).  The gradient
 * w.r.t. the prior-mean parameters is the host framework's (dLoss/dm(x_i) = h_mean[i]).  Dense route: K(x,Z) in
 * row chunks, M^3-shaped products on the in-tree tensor-pipe GEMM (dgemm.cu), hand-written contraction kernels (D <= 32) or the GEMM formulation of
 * the contraction for wider inputs (grad_wide.cu).                                                               */
size_t gvi_exact_gp_predict_grad_workspace_bytes(int64_t N, int M, int D);
int gvi_exact_gp_predict_grad_f64(const void* factor, int M, int D, const double* X, int64_t N, const double* g_var,
                                  const double* h_mean, const double* log_noise, double* out, void* workspace,
                                  void* stream);

/* ---- section 8(f)-2: classification epilogue on [K,N] predictive means / variances ---------------------------
 * GPClassificationBase._calculate_s_matrix + _predict_probability (src/gps/base/classification_base.py:53-153):
 *   s[n,j]    = 1/sqrt(pi) sum_h weights[h] prod_{l != j} max(Phi((sqrt2 roots[h] sd[j,n] + mean[j,n] - mean[l,n])
 *               / sd[l,n]), cdf_lower_bound),   sd = sqrt(var)
 *   prob[n,j] = (1 - epsilon) s + epsilon / (K - 1) (1 - s)
 * mean, var are [K,N] row-major; roots/weights are the H Gauss-Hermite nodes (device arrays); prob_out is [N,K].
 * The _grad entry is the vector-Jacobian product: dprob[N,K] -> dmean_out[K,N], dvar_out[K,N].  2 <= K <= 64. */
int gvi_class_probabilities_f64(const double* mean, const double* var, int K, int64_t N, const double* roots,
                                const double* weights, int H, double epsilon, double cdf_lower_bound,
                                double* prob_out, void* stream);
int gvi_class_probabilities_grad_f64(const double* mean, const double* var, int K, int64_t N, const double* roots,
                                     const double* weights, int H, double epsilon, double cdf_lower_bound,
                                     const double* dprob, double* dmean_out, double* dvar_out, void* stream);
/* Risks / regularisers on class probabilities ([N,K] row-major), as SUMS over the shard's points:
 *   out4[0] = sum_i -log sum_k q_ik y_ik                     (negative_log_likelihood.py:57-78; the reference sums)
 *   out4[1] = sum_i (sum_k |p_ik - q_ik|^power)^(1/power)    (multinomial_wasserstein_regularisation.py:46-76; mean)
 *   out4[2] = N
 *   out4[3] = sum_i -sum_k y_ik log_softmax(q_i)_k           (cross_entropy.py:11-30: jax_metrics' Crossentropy takes
 *                                                             `preds` as logits; pinned by the reference's 1.306689)
 * y / prob_p may be NULL (their terms are 0).  The _grad entry writes
 * dprob_out[i,k] = w[0] d out4[0]/dq_ik + w[1] d out4[1]/dq_ik + w[3] d out4[3]/dq_ik, w = weights4, a DEVICE array
 * [4] (the cotangent of out4).                                                                                 */
size_t gvi_multinomial_risk_workspace_bytes(int64_t N);
int gvi_multinomial_risk_f64(const double* prob_q, const double* y, const double* prob_p, int K, int64_t N,
                             int power, double* out4, void* workspace, void* stream);
int gvi_multinomial_risk_grad_f64(const double* prob_q, const double* y, const double* prob_p, int K, int64_t N,
                                  int power, const double* weights4, double* dprob_out, void* stream);

/* Weighted Gram C = sum_i g[i] k(Z,x_i) k(x_i,Z) ([M,M] row-major) on the tensor pipe: for the SVGP family
 * (var_i = k_ii - k_i^T A^-1 k_i + k_i^T Sigma k_i, kernel hyper-parameters frozen) this is dLoss/dSigma given
 * g = dLoss/dvar, whatever the parameterisation of Sigma (cholesky_svgp_kernel.py:133-150, diagonal_svgp_kernel.py:
 * 117-129, log_svgp_kernel.py:108-124).  Workspace: gvi_fused_step_grad_workspace_bytes(N, M, D).  D <= 32.      */
int gvi_weighted_gram_f64(const void* factor, int M, int D, const double* X, int64_t N, const double* g,
                          double* C_out, void* workspace, void* stream);

/* ---- section 8(f)-4: device-resident optimiser step (experiments/shared/trainer.py:42-100; the optax 0.1.7
 * transformations of experiments/shared/resolvers/optimiser.py:6-16 with optax's defaults).  One elementwise launch over
 * the flat parameter vector: param, grad, mu, nu are device arrays [n]; count is the 1-based step index.
 *   GVI_OPT_ADAM      b1 .9  b2 .999  eps 1e-8   eps_root 0       p -= lr mu^ / (sqrt(nu^ + eps_root) + eps)
 *   GVI_OPT_ADABELIEF b1 .9  b2 .999  eps 1e-16  eps_root 1e-16   nu tracks (g - mu_prev)^2 (+ eps_root)
 *   GVI_OPT_RMSPROP   b2 = decay .9   eps 1e-8                    p -= lr g / sqrt(nu + eps)   (mu unused, may be NULL) */
#define GVI_OPT_ADAM 0
#define GVI_OPT_ADABELIEF 1
#define GVI_OPT_RMSPROP 2
int gvi_optimiser_step_f64(int kind, double* param, const double* grad, double* mu, double* nu, int64_t n,
                           int64_t count, double learning_rate, double b1, double b2, double eps, double eps_root,
                           void* stream);

/* ---- a13: ConditionalVarianceInducingPointsSelector.compute_inducing_points
 * (src/inducing_points_selection.py:113-176) on ALREADY PERMUTED inputs x_perm[N,D] (the permutation is the
 * host's job).  Writes pivot positions into x_perm to idx_out[M_sel] (int64) and
 * tr(K-Q) after every step to trace_out[M_sel-1] (NULL ok) for the threshold test (lines 167-172).
 * workspace: gvi_select_workspace_bytes(N, M_sel) = the C matrix [(M_sel-1), N] + d[N] + scratch.
 * Multi-GPU use: each rank owns a contiguous slice; see the *_step_* entry points.                           */
size_t gvi_select_workspace_bytes(int64_t N, int M_sel, int D);
int gvi_cond_var_select_f64(const double* x_perm, int64_t N, int D, int M_sel, const double* log_scaling,
                            const double* log_ls, int ls_len, double jitter, int64_t* idx_out,
                            double* trace_out, void* workspace, void* stream);
/* Sharded selector, one call per pivot step and rank (N_local points starting at global offset `offset`):
 *   init    : d = diag + jitter, local first candidate record
 *   update  : given the winning record (global idx, x_j[D], c_j[m]) applies step m to the local slice and
 *             writes the next local candidate record
 *   resolve : picks the winner among R gathered records with the reference tie rule (max d, then max index)
 * Record layout (doubles): [0]=d value, [1]=global index (as double, exact < 2^53), [2..2+D)=x_j,
 * [2+D .. 2+D+M_sel-1) = c_j (column of C), so record_len = gvi_select_record_len(D, M_sel).                 */
int64_t gvi_select_record_len(int D, int M_sel);
int gvi_select_init_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel,
                        const double* log_scaling, double jitter, double* record_out, void* workspace,
                        void* stream);
/* The same steps for ANY kernel (the reference's selector only calls kernel.calculate_gram,
 * src/inducing_points_selection.py:121-166): the caller supplies kdiag[i] = k(x_i, x_i) at init and, per step,
 * kcol[i] = k(x_i, x_j) for the winning pivot j (its position is idx_out_m of resolve, a device value: the column can
 * be produced without a host sync).  np.round(., 20), the jitter at j, the clip and the tie rules are applied here. */
int gvi_select_init_gram_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel,
                             const double* kdiag, double jitter, double* record_out, void* workspace, void* stream);
int gvi_select_update_gram_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel, int m,
                               const double* winner, const double* kcol, double jitter, double* record_out,
                               double* trace_partial_out, void* workspace, void* stream);
int gvi_select_resolve_f64(const double* records, int R, int D, int M_sel, int first, double* winner_out,
                           int64_t* idx_out_m, void* stream);
int gvi_select_update_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel, int m,
                          const double* winner, const double* log_scaling, const double* log_ls, int ls_len,
                          double jitter, double* record_out, double* trace_partial_out, void* workspace,
                          void* stream);

/* ---- measurement aid (no reference counterpart): the FP64 tensor-pipe (DMMA m8n8k4) issue rate of the current device.
 * Runs a register-resident mma loop for about target_ms milliseconds on `stream`, SYNCHRONISES it and writes TFLOP/s
 * (2 flop per multiply-add) to the HOST double *tflops_out_h: the roofline denominator of the tile kernels, measured on
 * the box the bench line comes from.  scratch: device buffer of gvi_fp64_tensor_peak_scratch_bytes().                */
size_t gvi_fp64_tensor_peak_scratch_bytes(void);
int gvi_fp64_tensor_peak_tflops(double target_ms, double* tflops_out_h, void* scratch, void* stream);

/* ---- the permutation of a13 on the device (src/inducing_points_selection.py:115-119: jax.random.permutation) --------
 * out[i] = jax.random.bits(key, (n,), uint32)[i] widened to int64: threefry2x32 of jax 0.4.13 (third-party, pinned in the
 * reference's poetry.lock; jax_threefry_partitionable=False) over the counters arange(n).  key0/key1 are the two
 * uint32 words of the (sub)key - key splitting is a handful of hashes and stays on the host.  One shuffle round of
 * jax.random.permutation = a stable sort of the current order by these keys.                                       */
int gvi_threefry_random_bits(uint32_t key0, uint32_t key1, int64_t n, int64_t* out, void* stream);

/* ---- multi-GPU: the communicator (SURVEY.md section 8(b) "gvi_comm_init/destroy (wraps ncclComm_t)", 8(e)) -----------
 * The reference is single-process JAX and has no counterpart; the split these calls serve is: the N points sharded
 * contiguously by rank, Z / theta replicated, only the risk / gradient SUMS and the selector's per-pivot candidate cross
 * NVLink.  One rank per process and device (the device current at gvi_comm_init).  A communicator wraps an ncclComm_t
 * (libnccl.so.2 is bound with dlopen at the first multi-rank init: a single-GPU host does not need NCCL) and owns a
 * small PEER WINDOW that every other rank maps with CUDA IPC; the sharded selector writes its per-pivot records
 * straight into the peers' windows from its kernels.  Where the mapping is not possible (or GVI_COMM_NO_PEER_WINDOW is
 * passed by every rank) the same kernels exchange through ncclAllGather enqueued from C on the caller's stream.
 *   bootstrap: rank 0 fills id128_h (HOST buffer, GVI_COMM_UNIQUE_ID_BYTES) with gvi_comm_unique_id_h, the host
 *   distributes the bytes by its own means (torch.distributed / MPI / a file), every rank calls gvi_comm_init
 *   (collective, synchronises).  world == 1 needs no id.  Calls on one communicator are collective and must be made
 *   in the same order by every rank; a communicator is thread-compatible, not thread-safe.                          */
typedef struct gvi_comm gvi_comm;
#define GVI_COMM_UNIQUE_ID_BYTES 128
#define GVI_COMM_NO_PEER_WINDOW 1
int gvi_comm_unique_id_h(void* id128_h);
int gvi_comm_init(gvi_comm** comm_out, int rank, int world, const void* id128_h, int flags);
int gvi_comm_destroy(gvi_comm* comm);
int gvi_comm_rank(const gvi_comm* comm);
int gvi_comm_size(const gvi_comm* comm);
int gvi_comm_has_peer_window(const gvi_comm* comm);
/* buf[count] <- sum over ranks, in place, on the caller's stream.  Up to 8192 doubles x ranks the partials are gathered
 * and summed in RANK ORDER, so every rank holds the same bits whatever the transport did (the [sum NLL, sum D0, N]
 * triple of gvi_fused_step_f64, the 4 + D gradient vector); longer vectors (flat parameter gradients) go to
 * ncclAllReduce.                                                                                                */
int gvi_comm_allreduce_sum_f64(gvi_comm* comm, double* buf, int64_t count, void* stream);
/* recv[world][count] <- every rank's send[count] */
int gvi_comm_allgather_f64(gvi_comm* comm, const double* send, double* recv, int64_t count, void* stream);

/* ---- a13 sharded: ONE call runs the whole greedy selection over points sharded across the communicator's ranks
 * (reference loop: src/inducing_points_selection.py:137-172).  x_local[N_local, D] is this rank's contiguous slice of
 * the PERMUTED inputs, starting at global position `offset` (N_local == 0 is allowed).  Every rank receives the same
 * idx_out[M_sel] (global positions into the permuted order; -1 entries = the exchange timed out) and trace_out[M_sel-1]
 * (tr(K - Q) over ALL points after each step, summed in rank order; NULL ok).  Per pivot one update launch and one
 * one-CTA launch; the candidate record of a slice (d, index, x_j, c_j[0..m]) goes to every peer's window from that
 * kernel and the next update kernel resolves the winner itself - no host code and no collective call between pivots.
 * workspace: gvi_select_sharded_workspace_bytes(N_local, M_sel, D) (the slice's C [(M_sel-1), N_local], d, scratch). */
size_t gvi_select_sharded_workspace_bytes(int64_t N_local, int M_sel, int D);
int gvi_cond_var_select_sharded_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel,
                                    const double* log_scaling, const double* log_ls, int ls_len, double jitter,
                                    int64_t* idx_out, double* trace_out, void* workspace, gvi_comm* comm, void* stream);
/* The same kernels and exchange protocol over R VIRTUAL shards of one device's points (mailboxes local): shard r owns
 * [bounds_h[r], bounds_h[r+1]) (HOST array of R+1 ascending positions from 0 to N; empty shards allowed).
 * idx_out is [R][M_sel]: every shard's own view of the pivots (they agree); trace_out [R][M_sel-1] or NULL.  Lets a
 * one-GPU host check the multi-rank protocol bit for bit.                                                        */
size_t gvi_select_virtual_shards_workspace_bytes(int64_t N, int M_sel, int D, int R);
int gvi_cond_var_select_virtual_shards_f64(const double* x_perm, int64_t N, int D, int M_sel,
                                           const double* log_scaling, const double* log_ls, int ls_len, double jitter,
                                           int R, const int64_t* bounds_h, int64_t* idx_out, double* trace_out,
                                           void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GVIGP_H */
