"""CPU-only chain of trust for the GRADIENT oracle.

The GPU gradient tests compare the CUDA path with torch autograd through `oracle/torch_twin.py`.  That twin is only as
good as its agreement with `oracle/gp_oracle.py` (the NumPy restatement pinned on the reference's goldens), so here:
  (1) every twin function's forward value equals the NumPy oracle's on seeded inputs (<= 1e-12 relative), and
  (2) the twin's autograd gradients equal central finite differences of the NUMPY oracle's loss (<= 2e-6 relative to the
      gradient norm - the accuracy of an O(h^2) difference in float64 with h = 1e-5), for every regulariser, both jitter
      modes and through the exact-GP / SVGP / classification paths.
Nothing here touches the CUDA library.
"""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as O
from oracle import torch_twin as T

KINDS = ["bhattacharyya", "gaussian_wasserstein", "hellinger", "kl", "renyi", "gwi_no_eig", "squared_difference"]


def t64(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def problem(seed=0, n=37, m=9, d=3):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    z = rng.standard_normal((m, d))
    y = np.sin(x.sum(1)) + 0.1 * rng.standard_normal(n)
    yz = np.sin(z.sum(1))
    ls, ll = 0.3, np.log(np.array([0.5, 0.8, 1.3])[:d])
    return rng, x, z, y, yz, ls, ll


def np_regulariser(kind, m_p, c_p, m_q, c_q, alpha=0.5):
    if kind == "gwi_no_eig":
        return O.gwi_no_eig(m_p, c_p, m_q, c_q)
    if kind == "squared_difference":
        return O.gaussian_squared_difference(m_p, c_p, m_q, c_q)
    return O.projected_regularisation(kind, m_p, c_p, m_q, c_q, alpha)


def np_loss(x, y, z, m_q, ls, ll, m_p, v_p, kind, alpha, reg, is_absolute):
    gram = lambda a, b: O.ard_gram(ls, ll, a, b)
    diag = lambda a: O.ard_gram_diag(ls, ll, a)
    v_q = O.sparse_posterior_diag(gram, diag, z, x, reg, is_absolute)
    return O.gvi_loss(O.gaussian_nll(y, m_q, v_q), np_regulariser(kind, m_p, v_p, m_q, v_q, alpha))


def central(f, v, h=1e-5):
    v = np.array(v, dtype=np.float64)
    flat = v.reshape(-1)
    g = np.empty_like(flat)
    for i in range(flat.size):
        old = flat[i]
        flat[i] = old + h
        fp = f(v if v.ndim else float(flat[0]))
        flat[i] = old - h
        fm = f(v if v.ndim else float(flat[0]))
        flat[i] = old
        g[i] = (fp - fm) / (2 * h)
    return g.reshape(v.shape)


def test_forward_values_of_the_twin_equal_the_oracle():
    rng, x, z, y, yz, ls, ll = problem()
    assert rel(T.ard_gram(t64(ls), t64(ll), t64(x), t64(z)).numpy(), O.ard_gram(ls, ll, x, z)) < 1e-14
    # scalar lengthscale broadcast (README passes a 0-d log_lengthscales)
    assert rel(T.ard_gram(t64(ls), t64(-0.2), t64(x), t64(z)).numpy(), O.ard_gram(ls, np.array(-0.2), x, z)) < 1e-14
    k = O.ard_gram(ls, ll, z, z)
    for is_abs in (False, True):
        assert rel(T.add_diagonal_regulariser(t64(k), 1e-3, is_abs).numpy(),
                   O.add_diagonal_regulariser(k, 1e-3, is_abs)) < 1e-15
        got = T.sparse_posterior_diag(t64(ls), t64(ll), t64(z), t64(x), 1e-5, is_abs).numpy()
        want = O.sparse_posterior_diag(lambda a, b: O.ard_gram(ls, ll, a, b), lambda a: O.ard_gram_diag(ls, ll, a),
                                       z, x, 1e-5, is_abs)
        assert np.max(np.abs(got - want)) < 1e-11  # the variance is a difference of O(1) numbers
    pm = rng.standard_normal(x.shape[0])
    mean_t, var_t = T.exact_gp_posterior(t64(ls), t64(ll), t64(pm), t64(z), t64(yz), t64(np.log(0.7)), t64(x))
    mean_o, var_o = O.exact_gp_posterior(lambda a, b: O.ard_gram(ls, ll, a, b), lambda a: O.ard_gram_diag(ls, ll, a),
                                         pm, z, yz, np.log(0.7), x)
    assert rel(mean_t.numpy(), mean_o) < 1e-12 and np.max(np.abs(var_t.numpy() - var_o)) < 1e-12
    m_q, v_q = rng.standard_normal(x.shape[0]), rng.uniform(0.2, 2.0, x.shape[0])
    m_p, v_p = rng.standard_normal(x.shape[0]), rng.uniform(0.2, 2.0, x.shape[0])
    assert rel(float(T.gaussian_nll(t64(y), t64(m_q), t64(v_q))), O.gaussian_nll(y, m_q, v_q)) < 1e-14
    for kind in KINDS:
        got = float(T.regulariser(kind, t64(m_p), t64(v_p), t64(m_q), t64(v_q), 0.3))
        assert rel(got, np_regulariser(kind, m_p, v_p, m_q, v_q, 0.3)) < 1e-13, kind


def test_forward_values_of_the_widened_twins_equal_the_oracle():
    rng, x, z, y, yz, ls, ll = problem(seed=3)
    gram_r = lambda a, b: O.ard_gram(ls, ll, a, b)
    ls_b, ll_b = -0.1, ll + 0.2
    gram_b = lambda a, b: O.ard_gram(ls_b, ll_b, a, b)
    want = np.diag(O.fixed_sparse_posterior_gram(gram_b, gram_r, z, x))
    got = T.fixed_sparse_posterior_diag(t64(ls_b), t64(ll_b), t64(ls), t64(ll), t64(z), t64(x)).numpy()
    assert np.max(np.abs(got - want)) < 1e-11
    m = z.shape[0]
    sigma = O.cholesky_sigma(0.1 * rng.standard_normal((m, m)), 0.1 * rng.standard_normal(m))
    want = np.diag(O.svgp_gram(gram_r, z, sigma, x))
    got = T.svgp_diag(t64(ls), t64(ll), t64(z), t64(x), t64(sigma)).numpy()
    assert np.max(np.abs(got - want)) < 1e-11
    # classification: [K, N] means / variances -> [N, K] probabilities and the three multinomial risks
    kk, n = 4, 11
    means, variances = rng.standard_normal((kk, n)), rng.uniform(0.3, 1.5, (kk, n))
    roots, weights = np.polynomial.hermite.hermgauss(50)
    prob_o = O.classification_probabilities(means, variances)
    prob_t = T.classification_probabilities(t64(means), t64(variances), roots, weights).numpy()
    assert rel(prob_t, prob_o) < 1e-12
    labels = np.eye(kk)[rng.integers(0, kk, n)]
    other = O.classification_probabilities(means + 0.3, variances * 1.2)
    assert rel(float(T.multinomial_nll(t64(prob_o), t64(labels))), O.multinomial_nll(prob_o, labels)) < 1e-13
    assert rel(float(T.multinomial_wasserstein(t64(other), t64(prob_o))), O.multinomial_wasserstein(other, prob_o)) < 1e-13
    assert rel(float(T.cross_entropy_from_probabilities(t64(prob_o), t64(labels))),
               O.cross_entropy_from_probabilities(prob_o, labels)) < 1e-13


@pytest.mark.parametrize("is_absolute", [False, True])
@pytest.mark.parametrize("kind", KINDS)
def test_twin_gradients_equal_central_differences_of_the_numpy_oracle(kind, is_absolute):
    rng, x, z, y, yz, ls, ll = problem(seed=1)
    n = x.shape[0]
    m_q = 0.5 * rng.standard_normal(n)
    m_p, v_p = rng.standard_normal(n), rng.uniform(0.3, 1.5, n)
    alpha, reg = 0.5, 1e-3
    loss, g_ls, g_ll, g_mq = T.gvi_loss_and_grads(x, y, z, m_q, ls, ll, m_p, v_p, kind, alpha, reg, is_absolute)
    f = lambda ls_, ll_, mq_: np_loss(x, y, z, mq_, ls_, ll_, m_p, v_p, kind, alpha, reg, is_absolute)
    assert rel(loss, f(ls, ll, m_q)) < 1e-12
    fd_ls = central(lambda v: f(v, ll, m_q), ls)
    fd_ll = central(lambda v: f(ls, v, m_q), ll)
    fd_mq = central(lambda v: f(ls, ll, v), m_q)
    scale = max(abs(g_ls), float(np.max(np.abs(g_ll))), 1e-12)
    assert abs(g_ls - float(fd_ls)) <= 2e-6 * scale
    assert np.max(np.abs(g_ll - fd_ll)) <= 2e-6 * scale
    assert np.max(np.abs(g_mq - fd_mq)) <= 2e-6 * max(float(np.max(np.abs(g_mq))), 1e-12)


def test_relative_jitter_makes_the_variance_homogeneous_in_the_amplitude():
    """SURVEY.md section 8 'free self-check': with relative jitter v_i scales with exp(2 log_scaling), so
    d v_i / d log_scaling = 2 v_i exactly; with absolute jitter it does not."""
    rng, x, z, y, yz, ls, ll = problem(seed=2)
    for is_abs, holds in ((False, True), (True, False)):
        lst = t64(ls).clone().requires_grad_(True)
        v = T.sparse_posterior_diag(lst, t64(ll), t64(z), t64(x), 1e-2, is_abs)
        (g,) = torch.autograd.grad(v.sum(), lst)
        tot = float(v.sum().detach())
        err = abs(float(g) - 2.0 * tot) / abs(2.0 * tot)
        assert (err < 1e-10) == holds


def test_exact_gp_posterior_twin_gradients_equal_central_differences():
    rng, x, z, y, yz, ls, ll = problem(seed=4)
    pm = np.zeros(x.shape[0])
    log_noise = np.log(0.5)
    w_m, w_v = rng.standard_normal(x.shape[0]), rng.standard_normal(x.shape[0])

    def f(ls_, ll_, ln_):
        mean, var = O.exact_gp_posterior(lambda a, b: O.ard_gram(ls_, ll_, a, b),
                                         lambda a: O.ard_gram_diag(ls_, ll_, a), pm, z, yz, ln_, x)
        return float(w_m @ mean + w_v @ var)

    lst, llt, lnt = (t64(v).clone().requires_grad_(True) for v in (ls, ll, log_noise))
    mean, var = T.exact_gp_posterior(lst, llt, t64(pm), t64(z), t64(yz), lnt, t64(x))
    (t64(w_m) @ mean + t64(w_v) @ var).backward()
    fd = [central(lambda v: f(v, ll, log_noise), ls), central(lambda v: f(ls, v, log_noise), ll),
          central(lambda v: f(ls, ll, v), log_noise)]
    got = [lst.grad.numpy(), llt.grad.numpy(), lnt.grad.numpy()]
    scale = max(float(np.max(np.abs(g))) for g in got)
    for g, d in zip(got, fd):
        assert np.max(np.abs(g - d)) <= 2e-6 * scale


def test_classification_twin_gradients_equal_central_differences():
    rng = np.random.default_rng(5)
    kk, n = 3, 7
    means, variances = rng.standard_normal((kk, n)), rng.uniform(0.4, 1.2, (kk, n))
    labels = np.eye(kk)[rng.integers(0, kk, n)]
    roots, weights = np.polynomial.hermite.hermgauss(50)
    f = lambda mu, var: O.multinomial_nll(O.classification_probabilities(mu, var), labels)
    mt, vt = t64(means).clone().requires_grad_(True), t64(variances).clone().requires_grad_(True)
    T.multinomial_nll(T.classification_probabilities(mt, vt, roots, weights), t64(labels)).backward()
    fd_m = central(lambda v: f(v, variances), means)
    fd_v = central(lambda v: f(means, v), variances)
    scale = max(float(mt.grad.abs().max()), float(vt.grad.abs().max()))
    assert np.max(np.abs(mt.grad.numpy() - fd_m)) <= 2e-6 * scale
    assert np.max(np.abs(vt.grad.numpy() - fd_v)) <= 2e-6 * scale
