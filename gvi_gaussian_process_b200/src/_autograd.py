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
