"""torch-float64 autograd twin of gp_oracle (TEST INFRASTRUCTURE): the same reference arithmetic written with
differentiable torch CPU ops, so that d(loss)/d{log_scaling, log_lengthscales, m_q} - which the reference obtains
with jax.grad of GeneralisedVariationalInference.calculate_loss (experiments/shared/trainer.py:83-89) - can be
checked.  Pinned indirectly: its forward values equal gp_oracle's and its gradients equal central finite differences
of gp_oracle's loss (tests/test_twin_consistency.py, CPU), and gp_oracle is pinned to the reference's goldens.
Gradients themselves are unpinned in the reference (no test).
"""
from __future__ import annotations

import math

import torch


def ard_gram(log_scaling, log_ls, x1, x2):
    # src/kernels/standard/ard_kernel.py:58-66
    w = torch.exp(log_ls).reshape(-1)
    if w.numel() == 1:
        w = w.expand(x1.shape[1])
    diff = x1[:, None, :] - x2[None, :, :]
    return torch.exp(log_scaling) ** 2 * torch.exp(-0.5 * (diff * diff * w).sum(-1))


def add_diagonal_regulariser(k, reg, is_absolute):
    # src/utils/matrix_operations.py:25-28
    m = k.shape[0]
    eye = torch.eye(m, dtype=k.dtype)
    return k + (reg if is_absolute else reg * torch.trace(k) / m) * eye


def sparse_posterior_diag(log_scaling, log_ls, z, x, reg=1e-5, is_absolute=False):
    # sparse_posterior_kernel.py:63-96 in diagonal mode
    a = add_diagonal_regulariser(ard_gram(log_scaling, log_ls, z, z), reg, is_absolute)
    chol = torch.linalg.cholesky(a)
    kxz = ard_gram(log_scaling, log_ls, x, z)
    sol = torch.cholesky_solve(kxz.T, chol)
    return torch.exp(log_scaling) ** 2 - (kxz * sol.T).sum(-1)


def exact_gp_posterior(log_scaling, log_ls, prior_mean, z, yz, log_noise, x):
    # src/gps/base/base.py:495-526
    c = ard_gram(log_scaling, log_ls, z, z) + torch.exp(log_noise) * torch.eye(z.shape[0], dtype=z.dtype)
    chol = torch.linalg.cholesky(c)
    kzx = ard_gram(log_scaling, log_ls, z, x)
    mean = kzx.T @ torch.cholesky_solve(yz.reshape(-1, 1), chol).reshape(-1) + prior_mean
    var = torch.exp(log_scaling) ** 2 - (kzx * torch.cholesky_solve(kzx, chol)).sum(0)
    return mean, var


def gaussian_nll(y, m, v):
    sd = torch.sqrt(v)
    return (0.5 * ((y - m) / sd) ** 2 + torch.log(sd) + 0.5 * math.log(2 * math.pi)).mean()


def regulariser(kind, m_p, c_p, m_q, c_q, alpha=0.5):
    dm2 = (m_p - m_q) ** 2
    if kind == "bhattacharyya":
        f = 0.125 * dm2 / (c_p + c_q) + 0.5 * torch.log(((c_p + c_q) / 2) / torch.sqrt(c_p * c_q))
    elif kind == "gaussian_wasserstein":
        f = dm2 + c_p + c_q - 2 * torch.sqrt(c_p * c_q)
    elif kind == "hellinger":
        f = 1 - torch.sqrt((2 * torch.sqrt(c_p) * torch.sqrt(c_q)) / (c_p + c_q)) * torch.exp(-0.25 * dm2 / (c_p + c_q))
    elif kind == "kl":
        f = torch.log(torch.sqrt(c_q) / torch.sqrt(c_p)) + (c_p + dm2) / (2 * c_q) - 0.5
    elif kind == "renyi":
        mix = alpha * c_p + (1 - alpha) * c_q
        f = torch.log(torch.sqrt(c_p) / torch.sqrt(c_q)) + torch.log(c_p / mix) / (2 * (alpha - 1)) + 0.5 * alpha * dm2 / mix
    elif kind == "gwi_no_eig":
        f = dm2 + c_p + c_q
    elif kind == "squared_difference":
        f = dm2 + (c_p - c_q) ** 2
    else:
        raise ValueError(kind)
    return f.mean()


def gvi_loss_and_grads(x, y, z, m_q, ls_q, ll_q, m_p, v_p, kind, alpha=0.5, reg=1e-5, is_absolute=False):
    """loss = NLL + regulariser(kind) with p fixed; returns loss and d/d{log_scaling, log_lengthscales, m_q}."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    x, y, z, m_p, v_p = t(x), t(y), t(z), t(m_p), t(v_p)
    ls = t(ls_q).clone().requires_grad_(True)
    ll = t(ll_q).clone().requires_grad_(True)
    mq = t(m_q).clone().requires_grad_(True)
    v_q = sparse_posterior_diag(ls, ll, z, x, reg, is_absolute)
    loss = gaussian_nll(y, mq, v_q) + regulariser(kind, m_p, v_p, mq, v_q, alpha)
    loss.backward()
    return float(loss.detach()), float(ls.grad), ll.grad.numpy().copy(), mq.grad.numpy().copy()


def fixed_sparse_posterior_diag(ls_b, ll_b, ls_r, ll_r, z, x, reg=1e-5, is_absolute=False):
    # fixed_sparse_posterior_kernel.py:41-52, 78-99 (Cholesky from the frozen regulariser kernel)
    a = add_diagonal_regulariser(ard_gram(ls_r, ll_r, z, z), reg, is_absolute)
    chol = torch.linalg.cholesky(a)
    kxz = ard_gram(ls_b, ll_b, x, z)
    return torch.exp(ls_b) ** 2 - (kxz * torch.cholesky_solve(kxz.T, chol).T).sum(-1)


def svgp_diag(ls_r, ll_r, z, x, sigma, reg=1e-5, is_absolute=False):
    # cholesky_svgp_kernel.py:133-150 / diagonal_svgp_kernel.py:117-129 in diagonal mode
    a = add_diagonal_regulariser(ard_gram(ls_r, ll_r, z, z), reg, is_absolute)
    chol = torch.linalg.cholesky(a)
    kxz = ard_gram(ls_r, ll_r, x, z)
    return (torch.exp(ls_r) ** 2 - (kxz * torch.cholesky_solve(kxz.T, chol).T).sum(-1)
            + ((kxz @ sigma) * kxz).sum(-1))


# ---- section 8(f)-2: classification ------------------------------------------------------------------------------
def classification_probabilities(means, variances, roots, weights, epsilon=0.01, cdf_lower_bound=1e-10):
    """classification_base.py:53-153 with differentiable torch ops: [K,N],[K,N] -> [N,K]."""
    sd = torch.sqrt(variances)
    r = torch.as_tensor(roots, dtype=torch.float64)
    w = torch.as_tensor(weights, dtype=torch.float64)
    cdf_input = (math.sqrt(2.0) * r[:, None, None, None] * sd[None, :, None, :] + means[None, :, None, :]
                 - means[None, None, :, :]) / sd[None, None, :, :]
    log_cdf = torch.log(torch.clamp(torch.special.ndtr(cdf_input), min=cdf_lower_bound))
    diag = torch.diagonal(log_cdf, dim1=1, dim2=2).permute(0, 2, 1)  # (h, n, k) -> (h, k, n)
    comp = w[:, None, None] * torch.exp(log_cdf.sum(dim=2) - diag)
    s = ((1.0 / math.sqrt(math.pi)) * comp.sum(dim=0)).T
    k = means.shape[0]
    return (1.0 - epsilon) * s + (epsilon / (k - 1)) * (1.0 - s)


def multinomial_nll(prob, y):
    return -torch.log((prob * y).sum(dim=1)).sum()


def multinomial_wasserstein(prob_p, prob_q, power=2):
    return torch.pow(torch.pow(torch.abs(prob_p - prob_q), power).sum(dim=1), 1.0 / power).mean()


def cross_entropy_from_probabilities(prob, y):
    return (-(y * torch.log_softmax(prob, dim=1)).sum(dim=1)).mean()
