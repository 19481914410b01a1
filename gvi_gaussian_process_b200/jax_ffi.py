"""JAX side of the XLA-FFI adapter (csrc/ffi/gvigp_xla_ffi.cc): registration of the custom-call targets and the
functions a maintainer of the reference drops into its `src/` methods.

STATUS: jax / jaxlib are not installable in this repository's image, so this module has never been executed against JAX;
what IS tested here (tests/test_abi.py) is that the target table below matches the handler symbols the adapter defines,
that `register()` fails with a clear message when JAX is absent, and - with a recording stand-in for jax - that every wrapper
passes the operands, results, attribute names / types and buffer sizes the adapter's bindings and the C ABI declare.  The tested product surface is the C ABI under the
torch / NumPy host code of `gvi_gaussian_process_b200.src`.

Override points in the reference (file:line) and the wrapper that replaces each:
  ARDKernel._calculate_gram            src/kernels/standard/base.py:73-103            -> ard_gram
  PolynomialKernel / InnerProductKernel src/kernels/non_stationary/*.py                -> dot_gram
  SparsePosteriorKernel._calculate_gram (diagonal mode) sparse_posterior_kernel.py:54-96 -> gp_factor + gp_predict
  GPBase.calculate_posterior           src/gps/base/base.py:384-526                   -> gp_factor(exact_gp=True) + gp_predict
  NegativeLogLikelihood / projected regularisers                                      -> risk
  GeneralisedVariationalInference._calculate_loss  generalised_variational_inference.py:73-94 -> fused_loss (custom_vjp)
  ConditionalVarianceInducingPointsSelector.compute_inducing_points (after jax.random.permutation) -> cond_var_select
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
FFI_LIB_PATH = os.path.join(HERE, "libgvigp_ffi.so")

# custom-call target name -> handler symbol exported by libgvigp_ffi.so (XLA_FFI_DEFINE_HANDLER_SYMBOL in the adapter)
FFI_TARGETS = {
    "gvi_ard_gram": "gvi_ard_gram_ffi",
    "gvi_dot_gram": "gvi_dot_gram_ffi",
    "gvi_gp_factor": "gvi_gp_factor_ffi",
    "gvi_gp_predict": "gvi_gp_predict_ffi",
    "gvi_risk": "gvi_risk_ffi",
    "gvi_fused_step": "gvi_fused_step_ffi",
    "gvi_fused_step_grad": "gvi_fused_step_grad_ffi",
    "gvi_cond_var_select": "gvi_cond_var_select_ffi",
}

_registered = False


def _jax():
    try:
        import jax  # noqa: F401
        import jax.numpy as jnp  # noqa: F401
    except ImportError as e:  # pragma: no cover - the only branch this image can take
        raise ImportError("gvi_gaussian_process_b200.jax_ffi needs jax/jaxlib >= 0.4.31 (typed XLA FFI); this image has "
                          "none - use the torch front end gvi_gaussian_process_b200.src instead") from e
    ffi = getattr(jax, "ffi", None)
    if ffi is None:
        from jax.extend import ffi  # jax 0.4.31 - 0.4.38
    return jax, jnp, ffi


def register():
    """Load libgvigp_ffi.so (built by build.build_ffi) and register every handler for platform CUDA."""
    global _registered
    jax, _, ffi = _jax()
    if _registered:
        return
    if not os.path.exists(FFI_LIB_PATH):
        raise RuntimeError(f"{FFI_LIB_PATH} is missing: run gvi_gaussian_process_b200.build.build_ffi() first")
    from . import _lib

    _lib.load()  # libgvigp.so first: the adapter links against it
    lib = ctypes.CDLL(FFI_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for target, symbol in FFI_TARGETS.items():
        ffi.register_ffi_target(target, ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    jax.config.update("jax_enable_x64", True)
    _registered = True


def _sizes():
    from . import _lib

    return _lib.load()


# ---- wrappers (shapes follow the reference's methods) -----------------------------------------------------------------
def ard_gram(log_scaling, log_lengthscales, x1, x2):
    jax, jnp, ffi = _jax()
    m1, m2, d = x1.shape[0], x2.shape[0], x1.shape[1]
    ws = _sizes().gvi_ard_gram_workspace_bytes(m1, m2, d) // 8
    out, _ = ffi.ffi_call("gvi_ard_gram", (jax.ShapeDtypeStruct((m1, m2), jnp.float64),
                                           jax.ShapeDtypeStruct((ws,), jnp.float64)))(
        x1, x2, jnp.atleast_1d(log_scaling), jnp.atleast_1d(log_lengthscales))
    return out


def dot_gram(log_scaling, log_constant, degree, x1, x2):
    jax, jnp, ffi = _jax()
    has_constant = log_constant is not None
    return ffi.ffi_call("gvi_dot_gram", jax.ShapeDtypeStruct((x1.shape[0], x2.shape[0]), jnp.float64))(
        x1, x2, jnp.atleast_1d(log_scaling), jnp.atleast_1d(log_constant if has_constant else 0.0),
        degree=float(degree), has_constant=has_constant)


def gp_factor(z, log_scaling, log_lengthscales, diagonal_regularisation=1e-5, is_absolute=False, log_noise=None,
              y_z=None):
    """The factor state (a flat f64 buffer) of one GP over the inducing set z."""
    jax, jnp, ffi = _jax()
    m, d = z.shape
    lib = _sizes()
    exact = log_noise is not None
    factor, _ = ffi.ffi_call("gvi_gp_factor", (jax.ShapeDtypeStruct((lib.gvi_gp_factor_bytes(m, d) // 8,), jnp.float64),
                                               jax.ShapeDtypeStruct((lib.gvi_gp_factor_workspace_bytes(m) // 8,),
                                                                    jnp.float64)))(
        z, jnp.atleast_1d(log_scaling), jnp.atleast_1d(log_lengthscales),
        jnp.atleast_1d(log_noise if exact else 0.0), y_z if exact else jnp.zeros((m,)),
        diagonal_regularisation=float(diagonal_regularisation), is_absolute=bool(is_absolute), exact_gp=exact)
    return factor


def gp_predict(factor, m, x, prior_mean=None):
    jax, jnp, ffi = _jax()
    n, d = x.shape
    ws = _sizes().gvi_gp_predict_workspace_bytes_n(m, d, n) // 8
    has_prior = prior_mean is not None
    mean, var, _ = ffi.ffi_call("gvi_gp_predict", (jax.ShapeDtypeStruct((n,), jnp.float64),) * 2
                                + (jax.ShapeDtypeStruct((ws,), jnp.float64),))(
        factor, x, prior_mean if has_prior else jnp.zeros((n,)), m=int(m), has_prior_mean=has_prior)
    return mean, var


def risk(reg_kind, alpha, y, m_q, v_q, m_p=None, v_p=None):
    jax, jnp, ffi = _jax()
    n = m_q.shape[0]
    ws = _sizes().gvi_risk_workspace_bytes(n) // 8
    zeros = jnp.zeros((n,))
    out3, _ = ffi.ffi_call("gvi_risk", (jax.ShapeDtypeStruct((3,), jnp.float64),
                                        jax.ShapeDtypeStruct((ws,), jnp.float64)))(
        y, m_q, v_q, zeros if m_p is None else m_p, zeros if v_p is None else v_p, reg_kind=int(reg_kind),
        alpha=float(alpha))
    return out3


def fused_loss(m, reg_kind, alpha=0.5, diagonal_regularisation=1e-5):
    """Returns loss(log_scaling, log_lengthscales, m_q, z, factor_p, prior_mean_p, x, y): the GVI objective as ONE custom
    call, differentiable w.r.t. its first three arguments through a jax.custom_vjp whose backward is the CUDA gradient
    step (what jax.grad of calculate_loss computes in the reference)."""
    jax, jnp, ffi = _jax()
    lib = _sizes()

    @jax.custom_vjp
    def loss(log_scaling, log_lengthscales, m_q, z, factor_p, prior_mean_p, x, y):
        factor_q = gp_factor(z, log_scaling, log_lengthscales, diagonal_regularisation)
        ws = lib.gvi_fused_step_workspace_bytes_d(x.shape[0], m, x.shape[1]) // 8
        out3, _ = ffi.ffi_call("gvi_fused_step", (jax.ShapeDtypeStruct((3,), jnp.float64),
                                                  jax.ShapeDtypeStruct((ws,), jnp.float64)))(
            factor_q, factor_p, x, y, m_q, prior_mean_p, m=int(m), reg_kind=int(reg_kind), alpha=float(alpha))
        return (out3[0] + out3[1]) / out3[2]

    def fwd(log_scaling, log_lengthscales, m_q, z, factor_p, prior_mean_p, x, y):
        return loss(log_scaling, log_lengthscales, m_q, z, factor_p, prior_mean_p, x, y), (
            log_scaling, log_lengthscales, m_q, z, factor_p, prior_mean_p, x, y)

    def bwd(res, ct):
        log_scaling, log_lengthscales, m_q, z, factor_p, prior_mean_p, x, y = res
        n, d = x.shape
        factor_q = gp_factor(z, log_scaling, log_lengthscales, diagonal_regularisation)
        ws = lib.gvi_fused_step_grad_workspace_bytes(n, m, d) // 8
        out, dm, _ = ffi.ffi_call("gvi_fused_step_grad", (jax.ShapeDtypeStruct((4 + d,), jnp.float64),
                                                          jax.ShapeDtypeStruct((n,), jnp.float64),
                                                          jax.ShapeDtypeStruct((ws,), jnp.float64)))(
            factor_q, factor_p, x, y, m_q, prior_mean_p, m=int(m), reg_kind=int(reg_kind), alpha=float(alpha))
        scale = ct / out[2]
        d_ll = out[4:] * scale
        if jnp.ndim(log_lengthscales) == 0 or jnp.size(log_lengthscales) == 1:
            d_ll = jnp.sum(d_ll).reshape(jnp.shape(log_lengthscales))
        return (out[3] * scale, d_ll, dm * scale, None, None, None, None, None)

    loss.defvjp(fwd, bwd)
    return loss


def cond_var_select(x_perm, number_of_inducing_points, log_scaling, log_lengthscales, jitter=1e-12):
    """Pivot positions into x_perm (int64 [M]) and tr(K - Q) after every step ([M-1])."""
    jax, jnp, ffi = _jax()
    n, d = x_perm.shape
    ws = _sizes().gvi_select_workspace_bytes(n, number_of_inducing_points, d)
    idx, trace, _ = ffi.ffi_call("gvi_cond_var_select", (
        jax.ShapeDtypeStruct((number_of_inducing_points,), jnp.int64),
        jax.ShapeDtypeStruct((number_of_inducing_points - 1,), jnp.float64),
        jax.ShapeDtypeStruct((ws,), jnp.uint8)))(
        x_perm, jnp.atleast_1d(log_scaling), jnp.atleast_1d(log_lengthscales), m_sel=int(number_of_inducing_points),
        jitter=float(jitter))
    return idx, trace
