"""Pins the CPU oracle against every golden vector the reference's own tests hold for the hot path
(SURVEY.md section 8c).  Values and the reference test locations are cited per case."""
import numpy as np
import pytest

from oracle import gp_oracle as O
from oracle import jax_prng as P

ones_gram = lambda a, b: np.ones((np.atleast_2d(a).shape[0], np.atleast_2d(b).shape[0]))  # mockers/kernel.py:12-16
ones_diag = lambda a: np.ones((np.atleast_2d(a).shape[0],))


def eye_gram(a, b):  # mockers/kernel.py:19-24
    n1, n2 = np.atleast_2d(a).shape[0], np.atleast_2d(b).shape[0]
    return np.eye(max(n1, n2))[:n1, :n2]


X_TRAIN = np.array([[1.0, 3.0, 2.0], [1.5, 1.5, 9.5]])
Y_TRAIN = np.array([1.0, 1.5])
X = np.array([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5]])


def test_ard_scalar_exact():
    # tests/kernels/test_kernels.py:97-142 (exact ==)
    assert O.ard_kernel_scalar(0.5, [-0.5, 0.0, 0.5], [1.0, 2.0, 3.0], [1.5, 2.5, 3.5]) == 1.8095777100611745
    assert O.ard_kernel_scalar(np.log(1.0), -0.5, [6.0], [6.0]) == 1.0
    assert O.ard_gram(0.5, [-0.5, 0.0, 0.5], [[1.0, 2.0, 3.0]], [[1.5, 2.5, 3.5]])[0, 0] == pytest.approx(
        1.8095777100611745, rel=1e-15)


def test_inner_product_and_polynomial_kernel_goldens():
    # tests/kernels/test_kernels.py:16-47 (inner product) and 50-94 (polynomial), jnp.isclose in the reference
    assert np.isclose(O.inner_product_gram(np.log(2.4), [[1.0, 2.0, 3.0]], [[1.5, 2.5, 3.5]])[0, 0], 40.8, rtol=1e-12)
    assert np.isclose(O.inner_product_gram(np.log(2.2), [[6.0]], [[6.0]])[0, 0], 79.2, rtol=1e-12)
    assert np.isclose(O.inner_product_gram(np.log(2.0), [[6.0]], [[6.0]])[0, 0], 72.0, rtol=1e-12)
    assert np.isclose(O.polynomial_gram(0.5, 3.2, 3, [[1.0, 2.0, 3.0]], [[1.5, 2.5, 3.5]])[0, 0], 73403079.4946829,
                      rtol=1e-12)
    assert np.isclose(O.polynomial_gram(0.5, 3.2, 1, [[6.0]], [[6.0]])[0, 0], 884.81980837, rtol=1e-9)
    assert np.isclose(O.polynomial_gram(0.5, 3.2, 1.5, [[6.0]], [[6.0]])[0, 0], 26319.78000332, rtol=1e-9)
    # diagonal mode = pairs
    x = np.array([[1.0, 2.0], [0.5, -1.0], [3.0, 0.25]])
    z = np.array([[0.1, 0.2], [1.5, 1.0], [-2.0, 4.0]])
    g = O.polynomial_gram(0.1, -0.3, 2, x, z)
    np.testing.assert_allclose(O.gram_pairs(lambda a, b: O.polynomial_gram(0.1, -0.3, 2, a, b), x, z), np.diag(g),
                               rtol=1e-15)


def test_gram_shapes():
    # tests/kernels/test_kernels.py:145-211
    assert O.ard_gram(0.0, [0.0, 0.0, 0.0], [[1.0, 2.0, 3.0]], X).shape == (1, 2)
    assert O.ard_gram_diag(0.0, [0.0, 0.0, 0.0], X).shape == (2,)


def test_add_diagonal_regulariser():
    # tests/test_matrix_operations.py:95-145
    m = np.array([[1.0, 0.5], [0.5, 3.0]])
    np.testing.assert_allclose(O.add_diagonal_regulariser(m, 0.1, True), m + 0.1 * np.eye(2))
    np.testing.assert_allclose(O.add_diagonal_regulariser(m, 0.1, False), m + 0.1 * 2.0 * np.eye(2))


def test_sparse_posterior_gram_identity_kernel():
    # tests/kernels/test_sparse_posterior_kernels.py:15-64: diag 9.9999e-06, off-diagonal 0
    z = np.array([[5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    x = np.array([[5.3, 5.0, 6.0], [2.5, 4.5, 2.5]])
    r = O.sparse_posterior_gram(eye_gram, z, x, x, 1e-5, False)
    np.testing.assert_allclose(r, np.array([[9.9999e-06, 0.0], [0.0, 9.9999e-06]]), atol=1e-10)
    d = O.sparse_posterior_diag(eye_gram, ones_diag, z, x, 1e-5, False)
    np.testing.assert_allclose(d, [9.9999e-06, 9.9999e-06], atol=1e-10)


def test_exact_gp_posterior_goldens():
    # tests/gps/test_regression.py:14-56,100-200 (array_equal)
    mean, var = O.exact_gp_posterior(ones_gram, ones_diag, np.ones(2), X_TRAIN, Y_TRAIN, np.log(1.0), X)
    assert np.array_equal(mean, np.array([1.8333333333333335, 1.8333333333333335]))
    assert np.array_equal(var, np.array([0.33333333333333326, 0.33333333333333326]))


def test_approximate_gp_drops_noise():
    # tests/gps/test_regression.py:58-97,203-249: mean 1, var 1 although log_observation_noise = log 1 is passed
    mean, var = O.approximate_gp_predict(np.ones(2), ones_diag(X))
    assert np.array_equal(mean, np.ones(2)) and np.array_equal(var, np.ones(2))


def test_gaussian_nll_goldens():
    # tests/test_empirical_risks.py:20-144.  That module never enables x64, so its goldens are float32 values
    # compared with jnp.isclose (rtol 1e-5); the same tolerance is used here (hand value: 1.4112990555).
    mean, var = O.exact_gp_posterior(ones_gram, ones_diag, np.ones(2), X_TRAIN, Y_TRAIN, np.log(1.0), X)
    assert O.gaussian_nll(np.ones(2), mean, var) == pytest.approx(1.4112988, rel=1e-5)
    x12 = np.tile(X, (6, 1))
    mean, var = O.exact_gp_posterior(ones_gram, ones_diag, np.ones(12), X_TRAIN, Y_TRAIN, np.log(1.0), x12)
    assert O.gaussian_nll(np.ones(12), mean, var) == pytest.approx(1.4112988, rel=1e-5)
    assert O.gaussian_nll(np.array([1.2, 4.2]), np.ones(2), np.ones(2)) == pytest.approx(3.48893853, rel=1e-5)


REG_GOLDENS = {  # tests/regularisations/projected/test_projected_*.py:120-176
    "bhattacharyya": 0.13702468,
    "gaussian_wasserstein": 0.87307724,
    "hellinger": 0.18301035,
    "kl": 0.56319503,
    "renyi": 0.4042577,
}


@pytest.mark.parametrize("kind", sorted(REG_GOLDENS))
def test_projected_regulariser_goldens(kind):
    m_p, c_p = O.exact_gp_posterior(ones_gram, ones_diag, np.ones(2), X_TRAIN, Y_TRAIN, np.log(1.0), X)
    m_q, c_q = O.approximate_gp_predict(np.ones(2), ones_diag(X))
    assert O.projected_regularisation(kind, m_p, c_p, m_q, c_q, 0.5) == pytest.approx(REG_GOLDENS[kind], rel=1e-5)
    # "same GP => 0" cases
    assert O.projected_regularisation(kind, m_p, c_p, m_p, c_p, 0.5) == pytest.approx(0.0, abs=1e-12)


def test_gwi_and_squared_difference_goldens():
    m_p, c_p = O.exact_gp_posterior(ones_gram, ones_diag, np.ones(2), X_TRAIN, Y_TRAIN, np.log(1.0), X)
    m_q, c_q = np.ones(2), np.ones(2)
    # tests/regularisations/test_gaussian_squared_difference.py:120-176 uses the default full covariances:
    # posterior cov of the ones-kernel GP is 1/3 everywhere, q's "prior" cov is 1 everywhere
    full_p, full_q = np.full((1, 2, 2), 1.0 / 3.0), np.ones((1, 2, 2))
    assert O.gaussian_squared_difference(m_p, full_p, m_q, full_q) == pytest.approx(1.13888889, rel=1e-5)
    # tests/regularisations/test_gaussian_wasserstein.py:124-181 (with eigendecomposition): 0.87307723
    val = O.gwi_with_eig(m_p, full_p[0], m_q, full_q[0], full_p[0], full_q[0], 1e-8, False)
    assert val == pytest.approx(0.87307723, rel=1e-5)
    # the partial (no-eig) form is what the CUDA path implements; it differs by the cross term only
    assert O.gwi_no_eig(m_p, c_p, m_q, c_q) == pytest.approx(np.mean((m_p - m_q) ** 2) + 1.0 / 3.0 + 1.0)


def test_gvi_sum():
    assert O.gvi_loss(2.3, 1.2) == 3.5  # tests/test_generalised_variational_inference.py


def test_prng_pins():
    # tests/test_inducing_points_selection.py:128,139 (greedy) and :151,162 (random choice)
    key = P.prng_key(0)
    _, sub = P.split(key)
    assert list(P.permutation(sub, 4)) == [2, 3, 0, 1]
    assert list(P.permutation(sub, 3)) == [2, 0, 1]
    assert list(P.choice_with_replacement(key, 4, 2)) == [1, 2]
    assert list(P.choice_with_replacement(key, 3, 2)) == [0, 2]


def test_threefry_known_answers():
    """Third-party pins of the hash itself (jax 0.4.13 is not under /root/reference): the three known-answer vectors of
    the Threefry-2x32 (20 rounds) reference implementation (Random123 test_threefry.cpp), which jax's own
    tests/random_test.py::testThreefry2x32 asserts, and the documented output of split(PRNGKey(0))."""
    kat = [((0x0, 0x0), (0x0, 0x0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, count, want in kat:
        got = P.threefry_2x32(np.array(key, dtype=np.uint32), np.array(count, dtype=np.uint32))
        assert tuple(int(v) for v in got) == want
    assert P.split(P.prng_key(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]


@pytest.mark.parametrize("use_argsort", [True, False])
def test_selector_goldens(use_argsort):
    # tests/test_inducing_points_selection.py:116-182: PRNGKey(0), all-ones Gram, M=2 -> indices [2, 1]
    key = P.prng_key(0)
    _, sub = P.split(key)
    x4 = np.array([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5], [4.5, 2.5, 3.5], [5.5, 6.5, 7.5]])
    for x in (x4, np.arange(3 * 784, dtype=float).reshape(3, 784)):
        n = x.shape[0]
        perm = P.permutation(sub, n)
        piv = O.conditional_variance_select(x[perm], 2, ones_gram, ones_diag, 1e-12, 0.0, use_argsort)
        assert list(perm[piv]) == [2, 1]
    perm = P.permutation(sub, 4)
    piv = O.conditional_variance_select(x4[perm], 2, ones_gram, ones_diag)
    np.testing.assert_array_equal(x4[perm][piv], np.array([[4.5, 2.5, 3.5], [1.5, 2.5, 3.5]]))


def test_selector_argsort_and_maxrule_agree_on_real_kernel():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((300, 2))
    g = lambda a, b: O.ard_gram(0.1, [0.3, -0.2], a, b)
    d = lambda a: O.ard_gram_diag(0.1, [0.3, -0.2], a)
    a = O.conditional_variance_select(x, 12, g, d, use_argsort=True)
    b = O.conditional_variance_select(x, 12, g, d, use_argsort=False)
    assert np.array_equal(a, b) and len(set(a.tolist())) == 12


def test_selector_equals_the_dense_definition_of_greedy_conditional_variance():
    """Independent statement of what the pivoted-Cholesky recursion computes (no reference test pins the selector on a
    non-degenerate kernel): pivot m+1 maximises k(x,x) - k(x,S) K_SS^-1 k(S,x) over the points not in S = pivots[:m+1]."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((200, 2))
    g = lambda a, b: O.ard_gram(0.0, [0.5, 0.1], a, b)
    d = lambda a: O.ard_gram_diag(0.0, [0.5, 0.1], a)
    piv = O.conditional_variance_select(x, 10, g, d)
    k = g(x, x)
    s = [int(np.argmax(np.diag(k)))]
    for _ in range(9):
        ks = k[:, s]
        cond = np.diag(k) - np.einsum("ij,ji->i", ks, np.linalg.solve(k[np.ix_(s, s)], ks.T))
        cond[s] = -np.inf
        order = np.argsort(cond)
        assert cond[order[-1]] - cond[order[-2]] > 1e-9  # the choice is decided far above rounding
        s.append(int(order[-1]))
    assert piv.tolist() == s


def test_svgp_goldens():
    # tests/kernels/test_svgp_kernels.py:20-160 (initial parameters) and :202-300 (Grams), identity mock kernel
    x_train = np.array([[1.0, 2.0, 3.0], [5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    z = np.array([[5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    x = np.array([[5.3, 5.0, 6.0], [2.5, 4.5, 2.5]])
    lower, log_diag = O.svgp_cholesky_initial_parameters(eye_gram, z, x_train, np.log(1.0))
    assert np.array_equal(log_diag, np.array([-0.34657359027997275, -0.34657359027997275]))
    assert np.array_equal(lower, np.zeros((2, 2)))
    assert np.allclose(O.svgp_diagonal_initial_parameters(eye_gram, z, x_train, np.log(1.0)), -0.69314724, rtol=1e-6)
    sigma = O.cholesky_sigma(np.zeros((2, 2)), np.array([-0.34657359, -0.34657359]))
    g = O.svgp_gram(eye_gram, z, sigma, x, x)
    assert np.allclose(g, np.array([[0.5000099999000007, 0.0], [0.0, 0.5000099999000007]]))
    g = O.svgp_gram(eye_gram, z, np.diag(np.exp([-0.69314724, -0.69314724])), x, x)
    assert np.allclose(g, np.array([[0.50001, 0.0], [0.0, 0.50001]]))


def test_fixed_sparse_posterior_golden():
    # tests/kernels/test_sparse_posterior_kernels.py:66-114 (fixed variant, identity mock for both kernels)
    z = np.array([[5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    x = np.array([[5.3, 5.0, 6.0], [2.5, 4.5, 2.5]])
    r = O.fixed_sparse_posterior_gram(eye_gram, eye_gram, z, x, x, 1e-5, False)
    np.testing.assert_allclose(r, np.array([[9.9999e-06, 0.0], [0.0, 9.9999e-06]]), atol=1e-10)


# ---------------------------------------------------------------------------------------------------------------
# classification.  The reference's diagonal mode vmaps 1-point posteriors
# (src/gps/base/base.py:384-412), so position-dependent mock kernels (the identity mock) are evaluated point by point.
# ---------------------------------------------------------------------------------------------------------------
K_CLS = 4
Y_CLS = np.array([[0.5, 0.1, 0.2, 0.2], [0.1, 0.2, 0.3, 0.4]])
NOISE_A = np.log(np.array([1.0, 0.5, 0.2, 0.9]))
NOISE_B = np.log(np.array([0.1, 0.2, 0.4, 1.8]))
eye_diag = ones_diag


def pointwise_exact(gram_fns, diag_fns, x_train, y_train, noise, xs):
    ms, vs = [], []
    for i in range(xs.shape[0]):
        m, v = O.multi_output_exact_gp_posterior(gram_fns, diag_fns, np.ones((K_CLS, 1)), x_train, y_train, noise,
                                                 xs[i : i + 1])
        ms.append(m)
        vs.append(v)
    return np.concatenate(ms, axis=1), np.concatenate(vs, axis=1)


def test_classification_prediction_goldens():
    # tests/gps/test_classification.py:23-88 (array_equal in the reference; LAPACK/libm orderings allow 2 ulp here)
    m, v = pointwise_exact([ones_gram] * K_CLS, [ones_diag] * K_CLS, X, Y_CLS, NOISE_A, X_TRAIN)
    p = O.classification_probabilities(m, v)
    golden = np.array([0.29579238035540445, 0.19449721942556136, 0.2159886961681648, 0.29372170534428094])
    np.testing.assert_allclose(p, np.tile(golden, (2, 1)), rtol=1e-14)
    # :91-128 approximate GP, ones kernel -> uniform
    m, v = O.multi_output_prior([ones_diag] * K_CLS, np.ones((K_CLS, 2)), -np.inf, X_TRAIN)
    np.testing.assert_allclose(O.classification_probabilities(m, v), 0.25, rtol=1e-12)
    # :131-206 tempered exact GP on the identity mock
    temper = np.log(np.array([0.5, 1.0, 3.0, 4.0]))
    gf = [(lambda a, b, t=t: np.exp(t) * eye_gram(a, b)) for t in temper]
    df = [O.tempered_diag(eye_diag, t) for t in temper]
    m, v = pointwise_exact(gf, df, X, Y_CLS, NOISE_A, X_TRAIN)
    np.testing.assert_allclose(O.classification_probabilities(m, v),
                               np.tile([0.25269439, 0.19932216, 0.22324345, 0.32474002], (2, 1)), rtol=2e-7)
    # :209-265 tempered approximate GP: the observation noise passed in is DROPPED (generate_parameters)
    m, v = O.multi_output_prior([O.tempered_diag(ones_diag, t) for t in temper], np.ones((K_CLS, 2)), -np.inf,
                                X_TRAIN)
    np.testing.assert_allclose(O.classification_probabilities(m, v),
                               np.tile([0.1690632, 0.20674086, 0.29816303, 0.32603688], (2, 1)), rtol=2e-7)


def test_classification_risk_goldens():
    # tests/test_empirical_risks.py:147-212 (exact GP, identity mock): multinomial NLL 2.63190702
    m, v = pointwise_exact([eye_gram] * K_CLS, [eye_diag] * K_CLS, X_TRAIN, Y_CLS, NOISE_B, X)
    p = O.classification_probabilities(m, v)
    assert np.isclose(O.multinomial_nll(p, Y_CLS), 2.63190702, rtol=1e-8)
    # :215-269 cross entropy 1.306689: jax_metrics treats the probabilities as logits
    labels = np.array([[0, 0, 0, 1], [1, 0, 0, 0]])
    assert np.isclose(O.cross_entropy_from_probabilities(p, labels), 1.306689, rtol=1e-6)
    # :272-316 approximate GP NLL 2.77258869 = 2 log 4; :319-370 cross entropy 1.3862944 = log 4
    m, v = O.multi_output_prior([ones_diag] * K_CLS, np.ones((K_CLS, 2)), -np.inf, X)
    q = O.classification_probabilities(m, v)
    assert np.isclose(O.multinomial_nll(q, Y_CLS), 2.77258869, rtol=1e-8)
    assert np.isclose(O.cross_entropy_from_probabilities(q, np.array([[0, 0, 0, 1], [0, 1, 0, 0]])), 1.3862944,
                      rtol=1e-7)


def test_classification_regulariser_goldens():
    # tests/regularisations/**: K = 4 cases (line ~316 of each file); p = exact GP (ones mock), q = approximate GP
    m_p, v_p = pointwise_exact([ones_gram] * K_CLS, [ones_diag] * K_CLS, X_TRAIN, Y_CLS, NOISE_B, X)
    m_q, v_q = O.multi_output_prior([ones_diag] * K_CLS, np.ones((K_CLS, 2)), -np.inf, X)
    for kind, golden in [("bhattacharyya", 0.24135331), ("gaussian_wasserstein", 0.4287477),
                         ("hellinger", 0.208868), ("kl", 0.61610398), ("renyi", 0.49202457)]:
        assert np.isclose(O.projected_regularisation(kind, m_p, v_p, m_q, v_q), golden, rtol=2e-6), kind
    assert np.isclose(O.gaussian_squared_difference(m_p, v_p, m_q, v_q), 0.71837243, rtol=1e-7)
    # tests/regularisations/test_multinomial_wasserstein.py:130-197
    p = O.classification_probabilities(m_p, v_p)
    q = O.classification_probabilities(m_q, v_q)
    assert np.isclose(O.multinomial_wasserstein(p, q), 0.12691447, rtol=1e-7)
    assert O.multinomial_wasserstein(q, q, power=4) == 0.0


def test_log_svgp_goldens():
    # tests/kernels/test_svgp_kernels.py:150-199 (initial log-El) and :266-327 (Gram), identity mock kernel
    x_train = np.array([[1.0, 2.0, 3.0], [5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    z = np.array([[5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    x = np.array([[5.3, 5.0, 6.0], [2.5, 4.5, 2.5]])
    init = O.log_svgp_initial_parameters(eye_gram, z, x_train, np.log(1.0))
    np.testing.assert_allclose(init, np.array([[-0.34657362, -11.512925], [-11.512925, -0.34657362]]), rtol=1e-6)
    g = O.svgp_gram(eye_gram, z, O.log_svgp_sigma(init), x, x)
    np.testing.assert_allclose(g, np.array([[2.5000998e-01, 1.4142140e-05], [1.4142140e-05, 2.5000998e-01]]),
                               rtol=1e-6)


def test_committed_classification_fixture_is_current():
    """tests/golden/config4_mini_classification.npz must be what the oracle produces today (stale-fixture guard)."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config4_mini_classification.npz"))
    np.testing.assert_allclose(O.classification_probabilities(g["m_q"], g["v_q"]), g["prob_q"], rtol=1e-13)
    assert np.isclose(O.multinomial_nll(g["prob_q"], g["y"]), float(g["nll"]), rtol=1e-13)
    assert np.isclose(O.multinomial_wasserstein(g["prob_p"], g["prob_q"]), float(g["reg_multinomial_wasserstein"]),
                      rtol=1e-13)
    i = 1
    gram = lambda a, b: O.ard_gram(g["ls_q"][i], g["ll_q"][i], a, b)
    diag = lambda a: O.ard_gram_diag(g["ls_q"][i], g["ll_q"][i], a)
    np.testing.assert_allclose(O.sparse_posterior_diag(gram, diag, g["zs"][i], g["x"], float(g["lam"]), False),
                               g["v_q"][i], rtol=1e-12)


def test_adam_rule_equals_an_independent_implementation():
    """The reference holds no golden for an optimiser step (optax is third-party); the oracle's Adam rule is checked against
    torch.optim.Adam, an independent implementation of the same published update (bias-corrected moments, eps outside the
    square root).  AdaBelief / RMSprop have no second implementation in this image: they stay 'parity unpinned'."""
    import torch

    rng = np.random.default_rng(2)
    p0 = rng.standard_normal(17)
    grads = [rng.standard_normal(17) * (10.0 ** rng.integers(-3, 2)) for _ in range(6)]
    w = torch.tensor(p0.copy(), dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([w], lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    p, mu, nu = p0.copy(), np.zeros(17), np.zeros(17)
    for count, g in enumerate(grads, start=1):
        w.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
        p, mu, nu = O.optimiser_step("adam", p, g, mu, nu, count, 1e-2)
        np.testing.assert_allclose(p, w.detach().numpy(), rtol=1e-13, atol=1e-15)
