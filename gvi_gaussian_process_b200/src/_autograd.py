"""torch.autograd front end of the CUDA vector-Jacobian products (the role jax.grad plays in the reference's trainer,
experiments/shared/trainer.py:83-89).  Every backward here is a CUDA kernel of libgvigp; torch only chains them and
differentiates the host-framework mean function."""
from __future__ import annotations

import torch

from . import _lib_proxy as L


# The K per-class predictives of a multi-output GP run forward (and therefore backward) on side streams on purpose
# (_lib_proxy.run_concurrently); autograd synchronises the accumulation into the leaves correctly, the warning about
# the stream mismatch is expected.
if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


class PredictVariance(torch.autograd.Function):
    """var[N] = k(x,x) - ||L^-1 k(Z,x)||^2 of a factorised sparse posterior; differentiable w.r.t. the base kernel's
    (log_scaling, log_lengthscales).  `refactor()` rebuilds the factor state if another evaluation overwrote it
    between forward and backward."""

    @staticmethod
    def forward(ctx, factor, refactor, x, frozen, log_scaling, log_lengthscales):
        _, var = L.gp_predict(factor, x, want_mean=False)
        ctx.factor, ctx.refactor, ctx.x, ctx.frozen = factor, refactor, x, frozen
        ctx.version = factor.version
        ctx.shapes = (getattr(log_scaling, "shape", None), getattr(log_lengthscales, "shape", None))
        ctx.n_ll = log_lengthscales.numel() if isinstance(log_lengthscales, torch.Tensor) else 1
        return var

    @staticmethod
    def backward(ctx, g_var):
        factor = ctx.factor
        if factor.version != ctx.version:
            factor = ctx.refactor()
        out = L.gp_predict_grad(factor, ctx.x, g_var.contiguous(), ctx.frozen)
        d_ls = out[3].reshape(ctx.shapes[0]) if ctx.needs_input_grad[4] else None
        d_ll = None
        if ctx.needs_input_grad[5]:
            d_ll = out[4:].sum() if ctx.n_ll == 1 else out[4:]  # a scalar lengthscale is broadcast over D
            d_ll = d_ll.reshape(ctx.shapes[1])
        return None, None, None, None, d_ls, d_ll


class PointwiseRisk(torch.autograd.Function):
    """out3 = [sum NLL_i, sum D0_i, N] of gvi_risk_f64, differentiable w.r.t. (m_q, v_q)."""

    @staticmethod
    def forward(ctx, kind, alpha, y, m_p, v_p, m_q, v_q):
        ctx.args = (kind, alpha, y, m_p, v_p)
        ctx.save_for_backward(m_q, v_q)
        return L.risk(kind, alpha, y, m_q, v_q, m_p, v_p)

    @staticmethod
    def backward(ctx, g_out3):
        kind, alpha, y, m_p, v_p = ctx.args
        m_q, v_q = ctx.saved_tensors
        dm, dv = L.risk_grad(kind, alpha, y, m_q, v_q, m_p, v_p, g_out3[:2].contiguous())
        return None, None, None, None, None, dm, dv


class ClassProbabilities(torch.autograd.Function):
    """probabilities[N,K] from the [K,N] predictive means / variances (gvi_class_probabilities_f64)."""

    @staticmethod
    def forward(ctx, mean, var, roots, weights, epsilon, cdf_lower_bound):
        ctx.consts = (roots, weights, epsilon, cdf_lower_bound)
        ctx.save_for_backward(mean, var)
        return L.class_probabilities(mean, var, roots, weights, epsilon, cdf_lower_bound)

    @staticmethod
    def backward(ctx, dprob):
        roots, weights, epsilon, cdf_lower_bound = ctx.consts
        mean, var = ctx.saved_tensors
        dmean, dvar = L.class_probabilities_grad(mean, var, roots, weights, epsilon, cdf_lower_bound,
                                                 dprob.contiguous())
        return dmean, dvar, None, None, None, None


class MultinomialRisk(torch.autograd.Function):
    """out4 = [sum NLL, sum Wasserstein, N, sum cross-entropy] of gvi_multinomial_risk_f64, differentiable w.r.t. q."""

    @staticmethod
    def forward(ctx, prob_q, y, prob_p, power):
        ctx.args = (y, prob_p, power)
        ctx.save_for_backward(prob_q)
        return L.multinomial_risk(prob_q, y, prob_p, power)

    @staticmethod
    def backward(ctx, g_out4):
        y, prob_p, power = ctx.args
        (prob_q,) = ctx.saved_tensors
        return L.multinomial_risk_grad(prob_q, y, prob_p, power, g_out4.contiguous()), None, None, None


class ExactPredict(torch.autograd.Function):
    """(mean[N], var[N]) of an exact GP posterior (gvi_gp_predict_f64 on a state with alpha = A^-1 y), differentiable
    w.r.t. the prior mean values, log_observation_noise and the ARD kernel's (log_scaling, log_lengthscales):
    backward = gvi_exact_gp_predict_grad_f64."""

    @staticmethod
    def forward(ctx, factor, refactor, x, prior_mean, log_noise, log_scaling, log_lengthscales):
        mean, var = L.gp_predict(factor, x, prior_mean=prior_mean.contiguous())
        ctx.factor, ctx.refactor, ctx.x = factor, refactor, x
        ctx.version = factor.version
        ctx.log_noise = log_noise.detach().reshape(-1)[:1].contiguous()
        ctx.shapes = (log_noise.shape, getattr(log_scaling, "shape", None), getattr(log_lengthscales, "shape", None))
        ctx.n_ll = log_lengthscales.numel() if isinstance(log_lengthscales, torch.Tensor) else 1
        return mean, var

    @staticmethod
    def backward(ctx, h_mean, g_var):
        factor = ctx.factor
        if factor.version != ctx.version:
            factor = ctx.refactor()
        d = ctx.x.shape[1]
        if g_var is None:
            g_var = torch.zeros_like(h_mean)
        out = L.exact_gp_predict_grad(factor, ctx.x, g_var.contiguous(),
                                      None if h_mean is None else h_mean.contiguous(), ctx.log_noise)
        d_noise = out[4 + d].reshape(ctx.shapes[0]) if ctx.needs_input_grad[4] else None
        d_ls = out[3].reshape(ctx.shapes[1]) if ctx.needs_input_grad[5] else None
        d_ll = None
        if ctx.needs_input_grad[6]:
            d_ll = out[4:4 + d].sum() if ctx.n_ll == 1 else out[4:4 + d]
            d_ll = d_ll.reshape(ctx.shapes[2])
        return None, None, None, h_mean, d_noise, d_ls, d_ll


class SVGPVariance(torch.autograd.Function):
    """var[N] = k_ii - ||L^-1 k_i||^2 + k_i^T Sigma k_i of an SVGP-family kernel (frozen hyper-parameters), computed by
    the tile kernel with a lower-triangular E (E^T E = Sigma) attached to the factor state; differentiable w.r.t.
    Sigma: dLoss/dSigma = sum_i g_i k_i k_i^T (gvi_weighted_gram_f64)."""

    @staticmethod
    def forward(ctx, factor, x, sigma):
        _, var = L.gp_predict(factor, x, want_mean=False)
        ctx.factor, ctx.x = factor, x
        return var

    @staticmethod
    def backward(ctx, g_var):
        return None, None, L.weighted_gram(ctx.factor, ctx.x, g_var.contiguous())


# ---- kernels other than ARD: the predictive over caller-supplied Grams, differentiable w.r.t. the Grams ------------------
def tensor_leaves(obj):
    """The torch tensors inside a parameter container / dict / list (used to decide whether a backward is wanted)."""
    from .module import ModuleParameters

    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, ModuleParameters):
        yield from tensor_leaves(obj.dict_shallow())
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from tensor_leaves(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from tensor_leaves(v)


class DotGram(torch.autograd.Function):
    """(exp(log_scaling) x1 x2^T + exp(log_constant))^degree: forward = the CUDA Gram kernel (gvi_dot_gram_f64); backward
    w.r.t. the two scalars AND the inputs (the features of a CustomMappingKernel) as library GEMMs + elementwise ops."""

    @staticmethod
    def forward(ctx, x1, x2, log_scaling, log_constant, degree):
        x1, x2 = x1.contiguous(), x2.contiguous()
        out = torch.empty(x1.shape[0], x2.shape[0], dtype=torch.float64, device=x1.device)
        L.dot_gram(x1, x2, log_scaling.detach().reshape(-1)[:1].contiguous(),
                   None if log_constant is None else log_constant.detach().reshape(-1)[:1].contiguous(), degree, out)
        ctx.save_for_backward(x1, x2, log_scaling, log_constant if log_constant is not None else log_scaling)
        ctx.has_constant, ctx.degree = log_constant is not None, float(degree)
        return out

    @staticmethod
    def backward(ctx, g):
        x1, x2, log_scaling, log_constant = ctx.saved_tensors
        scale = torch.exp(log_scaling.detach()).reshape(())
        const = torch.exp(log_constant.detach()).reshape(()) if ctx.has_constant else None
        dot = x1 @ x2.T
        if ctx.degree != 1.0:
            base = scale * dot + (const if const is not None else 0.0)
            g = g * (ctx.degree * base ** (ctx.degree - 1.0))
        d_ls = ((g * dot).sum() * scale).reshape(log_scaling.shape) if ctx.needs_input_grad[2] else None
        d_lc = (g.sum() * const).reshape(log_constant.shape) if ctx.has_constant and ctx.needs_input_grad[3] else None
        d_x1 = scale * (g @ x2) if ctx.needs_input_grad[0] else None
        d_x2 = scale * (g.T @ x1) if ctx.needs_input_grad[1] else None
        return d_x1, d_x2, d_ls, d_lc, None


class PredictGram(torch.autograd.Function):
    """(mean[n], var[n]) = (K_xz alpha + prior_mean, k_diag - ||L^-1 K_xz^T||^2) through gvi_gp_predict_gram_f64 on a state
    built by gvi_gp_factor_gram_f64, differentiable w.r.t. K_zz, K_xz, k_diag, the prior mean and log_observation_noise.
    The backward is the dense vector-Jacobian product (what reverse-mode differentiation of cho_solve gives) as library
    GEMMs against the L^-1 of the CUDA factorisation; the kernel's own Gram routine is differentiated by whoever built
    K (DotGram above, or the host framework for CustomKernel / feature maps).
      A = K_zz + jitter I (+ exp(log_noise) I),  B = K_xz A^-1,  alpha = A^-1 y_z
      dK_xz = -2 diag(g) B + h alpha^T,  dk_diag = g,  dA = B^T diag(g) B - (A^-1 K_xz^T h) alpha^T
      dK_zz = dA (+ reg / M tr(dA) I for a trace-relative jitter),  dlog_noise = exp(log_noise) tr(dA)."""

    @staticmethod
    def forward(ctx, factor, linv, alpha, reg, is_abs, k_zz, k_xz, k_diag, prior_mean, log_noise):
        n = k_xz.shape[0]
        mean = torch.empty(n, dtype=torch.float64, device=k_xz.device)
        var = torch.empty(n, dtype=torch.float64, device=k_xz.device)
        k_xz_c = k_xz.contiguous()
        L.gp_predict_gram(factor, k_xz_c, k_diag.contiguous(), None if prior_mean is None else prior_mean.contiguous(),
                          None, mean, var)
        ctx.save_for_backward(linv, k_xz_c)
        ctx.alpha, ctx.reg, ctx.is_abs = alpha, float(reg), bool(is_abs)
        ctx.log_noise = None if log_noise is None else log_noise.detach()
        ctx.noise_shape = None if log_noise is None else log_noise.shape
        ctx.has_prior = prior_mean is not None
        return mean, var

    @staticmethod
    def backward(ctx, h, g):
        linv, k_xz = ctx.saved_tensors
        m = linv.shape[0]
        d_kxz = torch.zeros_like(k_xz)
        d_a = torch.zeros(m, m, dtype=torch.float64, device=k_xz.device)
        if g is not None:
            b = (k_xz @ linv.T) @ linv
            gb = g.reshape(-1, 1) * b
            d_kxz -= 2.0 * gb
            d_a += b.T @ gb
        if h is not None and ctx.alpha is not None:
            d_kxz += h.reshape(-1, 1) * ctx.alpha.reshape(1, -1)
            w = linv.T @ (linv @ (k_xz.T @ h))
            d_a -= torch.outer(w, ctx.alpha)
        d_kzz = d_a if ctx.is_abs else d_a + (ctx.reg / m) * torch.trace(d_a) * torch.eye(
            m, dtype=torch.float64, device=k_xz.device)
        d_noise = None
        if ctx.log_noise is not None and ctx.needs_input_grad[9]:
            d_noise = (torch.exp(ctx.log_noise).reshape(()) * torch.trace(d_a)).reshape(ctx.noise_shape)
        return (None, None, None, None, None, d_kzz if ctx.needs_input_grad[5] else None,
                d_kxz if ctx.needs_input_grad[6] else None, g if ctx.needs_input_grad[7] else None,
                h if ctx.has_prior and ctx.needs_input_grad[8] else None, d_noise)


class SVGPVarianceGram(torch.autograd.Function):
    """SVGPVariance for a regulariser kernel other than ARD: the variance comes from the tile kernel over a supplied
    K_xz chunk (state with E attached), dLoss/dSigma = K_xz^T diag(g) K_xz is one library GEMM on the same chunk."""

    @staticmethod
    def forward(ctx, factor, k_xz, k_diag, sigma):
        n = k_xz.shape[0]
        var = torch.empty(n, dtype=torch.float64, device=k_xz.device)
        L.gp_predict_gram(factor, k_xz, k_diag, None, None, None, var)
        ctx.save_for_backward(k_xz)
        return var

    @staticmethod
    def backward(ctx, g_var):
        (k_xz,) = ctx.saved_tensors
        return None, None, None, k_xz.T @ (g_var.reshape(-1, 1) * k_xz)
