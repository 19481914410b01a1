"""Generates tests/golden/*.npz from the CPU oracle on seeded inputs (configs #1 and #2 of BASELINE.json, and config #4 in
miniature for the classification path).

The reference itself cannot be imported here (needs jax==0.4.13, absent), so these fixtures are
ORACLE outputs; the oracle is pinned to the reference's own golden values in tests/test_oracle_goldens.py.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import gp_oracle as O  # noqa: E402
from oracle import jax_prng as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, n, d, m, log_scaling, log_ls, x, y, seed, lam=1e-5):
    rng = np.random.default_rng(seed)
    gram = lambda a, b: O.ard_gram(log_scaling, log_ls, a, b)
    diag = lambda a: O.ard_gram_diag(log_scaling, log_ls, a)
    key = P.prng_key(0)
    _, sub = P.split(key)
    perm = P.permutation(sub, n)
    piv = O.conditional_variance_select(x[perm], m, gram, diag, use_argsort=False)
    idx = perm[piv]
    z, yz = x[idx], y[idx]
    w1, w2 = rng.standard_normal((d, 10)), rng.standard_normal(10)
    m_q = np.tanh(x @ w1) @ w2 * 0.3
    # q: sparse posterior kernel with slightly different hyper-parameters; p: exact GP on (z, yz), noise 1
    ls_q, ll_q = log_scaling - 0.1, np.atleast_1d(log_ls) + 0.05
    gq = lambda a, b: O.ard_gram(ls_q, ll_q, a, b)
    dq = lambda a: O.ard_gram_diag(ls_q, ll_q, a)
    v_q = O.sparse_posterior_diag(gq, dq, z, x, lam, False)
    m_p, v_p = O.exact_gp_posterior(gram, diag, np.zeros(n), z, yz, 0.0, x)
    out = dict(x=x, y=y, idx=idx, m_q=m_q, v_q=v_q, m_p=m_p, v_p=v_p, log_scaling=log_scaling,
               log_ls=np.atleast_1d(log_ls), ls_q=ls_q, ll_q=ll_q, lam=lam,
               gram_xz=gram(x[:64], z), nll=O.gaussian_nll(y, m_q, v_q))
    for kind in O.D0:
        out["reg_" + kind] = O.projected_regularisation(kind, m_p, v_p, m_q, v_q, 0.5)
    out["reg_gwi_no_eig"] = O.gwi_no_eig(m_p, v_p, m_q, v_q)
    out["reg_squared_difference"] = O.gaussian_squared_difference(m_p, v_p, m_q, v_q)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "nll", out["nll"], "renyi", out["reg_renyi"], "idx[:5]", idx[:5])


def make_classification(name, n, d, k, m, seed, lam=1e-5):
    """config #4 in miniature (MNIST-shaped classification): inputs in [0, 1]^d, K classes, per-class ARD sparse
    posteriors over their own inducing sets (q), exact GP classifier on a combined inducing set (p)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 1.0, (n, d))
    labels = rng.integers(0, k, n)
    y = np.eye(k)[labels]
    zs = np.stack([x[rng.permutation(n)[:m]] for _ in range(k)])
    zi = rng.permutation(n)[:m]
    ls_q, ll_q = rng.uniform(-0.2, 0.2, k), np.log(rng.uniform(0.5, 1.5, (k, d)) * 4.0 / d)
    ls_p, ll_p = rng.uniform(-0.2, 0.2, k), np.log(rng.uniform(0.5, 1.5, (k, d)) * 4.0 / d)
    noise_p = np.log(rng.uniform(0.3, 1.0, k))
    w = rng.standard_normal((d, k)) / np.sqrt(d)
    m_q = (x @ w).reshape(k, -1)  # MeanBase.predict reshapes (n, k) -> (k, n) (means/base.py:97-100)
    v_q = np.stack([O.sparse_posterior_diag(lambda a, b, i=i: O.ard_gram(ls_q[i], ll_q[i], a, b),
                                            lambda a, i=i: O.ard_gram_diag(ls_q[i], ll_q[i], a), zs[i], x, lam, False)
                    for i in range(k)])
    gf = [(lambda a, b, i=i: O.ard_gram(ls_p[i], ll_p[i], a, b)) for i in range(k)]
    df = [(lambda a, i=i: O.ard_gram_diag(ls_p[i], ll_p[i], a)) for i in range(k)]
    m_p, v_p = O.multi_output_exact_gp_posterior(gf, df, np.zeros((k, n)), x[zi], y[zi], noise_p, x)
    prob_q = O.classification_probabilities(m_q, v_q)
    prob_p = O.classification_probabilities(m_p, v_p)
    out = dict(x=x, y=y, zs=zs, zi=zi, ls_q=ls_q, ll_q=ll_q, ls_p=ls_p, ll_p=ll_p, noise_p=noise_p, w=w, lam=lam,
               m_q=m_q, v_q=v_q, m_p=m_p, v_p=v_p, prob_q=prob_q, prob_p=prob_p,
               nll=O.multinomial_nll(prob_q, y), cross_entropy=O.cross_entropy_from_probabilities(prob_q, y),
               reg_multinomial_wasserstein=O.multinomial_wasserstein(prob_p, prob_q),
               reg_gwi_no_eig=O.gwi_no_eig(m_p, v_p, m_q, v_q))
    for kind in O.D0:
        out["reg_" + kind] = O.projected_regularisation(kind, m_p, v_p, m_q, v_q, 0.5)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "nll", out["nll"], "kl", out["reg_kl"], "wasserstein", out["reg_multinomial_wasserstein"])


if __name__ == "__main__":
    make_classification("config4_mini_classification", 400, 16, 4, 24, seed=4)
    rng = np.random.default_rng(0)
    # config #1: README toy sine (README.md:65-72, 92-94): N=100, D=1, scaling=10, weight=10, M=10
    x = np.linspace(-1, 1, 100).reshape(-1, 1)
    y = (np.sin(np.pi * x) + 0.1 * rng.standard_normal(x.shape)).reshape(-1)
    make("config1_readme", 100, 1, 10, np.log(10.0), np.log(10.0), x, y, seed=1)
    # config #2: toy curves (experiments/toy_curves/curves.py:29-37): N=1e4, D=1, M=100, scaling=1.  The
    # lengthscale weight is 400 (lengthscale 0.05 ~ the inducing spacing) instead of the config's initial value 1 so
    # that cond(K_ZZ + jitter) stays ~1e3 and the 1e-10 parity tolerance is meaningful (SURVEY.md section 7,
    # "parity at 1e-10 is a conditioning question"); the ill-conditioned setting is reported in DESIGN.md.
    x = np.linspace(-2, 2, 10000).reshape(-1, 1)
    y = (2 * np.sin(2 * np.pi * x) + 0.5 * rng.standard_normal(x.shape)).reshape(-1)
    make("config2_toy_curves", 10000, 1, 100, 0.0, np.log(400.0), x, y, seed=2)
