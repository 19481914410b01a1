"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI / the src-compatible
shim, against the CPU oracle on the same seeded inputs, against the committed golden fixtures, and - at
BASELINE.json's full size - through size-independent properties.

Tolerances (BASELINE.json north_star): Gram / mean / variance / risk 1e-10 relative (variance normwise:
max|delta| / max|ref|), selector indices bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O
from oracle import jax_prng as OP

pytestmark = pytest.mark.gpu

TOL = 1e-10
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def S():
    """The src-compatible shim, imported lazily so collection works without a GPU."""
    import types

    from gvi_gaussian_process_b200 import _lib, distributed
    from gvi_gaussian_process_b200.src import _lib_proxy
    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
    from gvi_gaussian_process_b200.src.gps import ApproximateGPRegression, GPRegression
    from gvi_gaussian_process_b200.src.inducing_points_selection import (
        ConditionalVarianceInducingPointsSelector, RandomInducingPointsSelector)
    from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean
    from gvi_gaussian_process_b200.src.regularisations import (
        GaussianSquaredDifferenceRegularisation, GaussianWassersteinRegularisation)
    from gvi_gaussian_process_b200.src.regularisations import projected
    from gvi_gaussian_process_b200.src.utils import jax_prng

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.load()  # fails loudly if libgvigp.so is missing
    return types.SimpleNamespace(**locals())


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


def rel_elem(a, ref):
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), 1e-300)))


def rel_norm(a, ref):
    return float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))


# ---------------------------------------------------------------------------------------------------------------
# a1-a3 Gram
# ---------------------------------------------------------------------------------------------------------------
def test_ard_scalar_golden(S):
    # reference tests/kernels/test_kernels.py:97-142 pins 1.8095777100611745 with ==; CUDA's exp() is not the same
    # libm, so the GPU value is required to agree to 2 ulp (the oracle reproduces the value exactly on the CPU).
    k = S.ARDKernel(number_of_dimensions=3)
    p = k.Parameters.construct(log_scaling=0.5, log_lengthscales=np.array([-0.5, 0.0, 0.5]))
    g = host(k.calculate_gram(p, np.array([[1.0, 2.0, 3.0]]), np.array([[1.5, 2.5, 3.5]])))
    assert g.shape == (1, 1)
    assert abs(g[0, 0] - 1.8095777100611745) <= 2 * np.spacing(1.8095777100611745)
    k1 = S.ARDKernel(number_of_dimensions=1)
    p1 = k1.Parameters.construct(log_scaling=np.log(1.0), log_lengthscales=-0.5)
    assert host(k1.calculate_gram(p1, np.array([[6.0]]), np.array([[6.0]])))[0, 0] == 1.0


@pytest.mark.parametrize("m1,m2,d", [(1, 2, 3), (37, 300, 1), (300, 257, 8), (65, 33, 40), (1000, 1, 5)])
def test_ard_gram_parity(S, m1, m2, d):
    rng = np.random.default_rng(m1 * 1000 + m2)
    x1, x2 = rng.standard_normal((m1, d)), rng.standard_normal((m2, d))
    ls, ll = 0.3, rng.normal(-1.0, 0.5, d)
    k = S.ARDKernel(number_of_dimensions=d)
    p = k.Parameters.construct(log_scaling=ls, log_lengthscales=ll)
    g = host(k.calculate_gram(p, x1, x2))
    assert g.shape == (m1, m2)
    assert rel_elem(g, O.ard_gram(ls, ll, x1, x2)) <= TOL
    # x2=None -> Gram with itself; diagonal mode returns [m1] (src/kernels/base.py:132-147)
    gs = host(k.calculate_gram(p, x1))
    assert rel_elem(gs, O.ard_gram(ls, ll, x1, x1)) <= TOL
    gd = host(k.calculate_gram(p, x1, x1, full_covariance=False))
    assert gd.shape == (m1,) and rel_elem(gd, O.ard_gram_diag(ls, ll, x1)) <= 1e-15


@pytest.mark.parametrize("m1,m2,d,offset", [(300, 257, 784, 0.0), (129, 128, 24, 0.0), (200, 130, 100, 1000.0)])
def test_ard_gram_tensor_pipe_contraction(S, m1, m2, d, offset):
    """D >= 24: x1.x2^T on the FP64 tensor pipe with the fused distance/exp epilogue (gram_dmma.cu) against the
    direct-difference oracle, incl. inputs far from the origin (UCI inputs are not standardised) - the common
    shift keeps the expansion |x|^2 + |z|^2 - 2 x.z from cancelling."""
    rng = np.random.default_rng(d)
    x1, x2 = offset + rng.uniform(0, 1, (m1, d)), offset + rng.uniform(0, 1, (m2, d))
    ls, ll = 0.2, np.log(1.0 / d) + 0.2 * rng.standard_normal(d)
    k = S.ARDKernel(number_of_dimensions=d)
    p = k.Parameters.construct(log_scaling=ls, log_lengthscales=ll)
    g = host(k.calculate_gram(p, x1, x2))
    ref = O.ard_gram(ls, ll, x1, x2)
    assert rel_elem(g, ref) <= TOL
    # and the direct-difference CUDA kernel agrees as well
    out = torch.empty(m1, m2, dtype=torch.float64, device="cuda")
    S._lib_proxy.ard_gram(dev(x1), dev(x2), dev([ls]), dev(ll), out, tensor_pipe=False)
    assert rel_elem(host(out), ref) <= TOL


@pytest.mark.parametrize("m1,m2,d", [(2051, 517, 8), (1000, 300, 5), (1500, 1030, 13), (777, 600, 20)])
def test_ard_gram_guarded_tensor_pipe_small_d(S, m1, m2, d):
    """5 <= D < 24 (csrc/gram_mma.cu): the x.z part of the exponent runs on the tensor pipe when the guard
    (R1 + R2)^2 <= 4096 over the scaled, centred inputs holds, else the direct-difference kernel does the work -
    decided on the device.  Checked here: the fast path against the oracle (<= 1e-10 elementwise, inputs far from the
    origin included: the centring removes a common offset), the tripped guard (short lengthscales: the result must be
    the direct kernel's, bit for bit), ragged sizes, an odd leading dimension, and a NaN input (-> NaN row)."""
    rng = np.random.default_rng(m1 + d)
    ls = 0.3
    k = S.ARDKernel(number_of_dimensions=d)
    for offset, log_ls, fast in ((0.0, np.log(1.0 / d), True), (1000.0, np.log(1.0 / d), True), (0.0, 9.0, False)):
        x1, x2 = offset + rng.standard_normal((m1, d)), offset + rng.standard_normal((m2, d))
        ll = log_ls + 0.3 * rng.standard_normal(d)
        ref = O.ard_gram(ls, ll, x1, x2)
        g = host(k.calculate_gram(k.Parameters.construct(log_scaling=ls, log_lengthscales=ll), x1, x2))
        assert rel_elem(g, ref) <= TOL, (offset, log_ls)
        direct = torch.empty(m1, m2, dtype=torch.float64, device="cuda")
        S._lib_proxy.ard_gram(dev(x1), dev(x2), dev([ls]), dev(ll), direct, tensor_pipe=False)
        assert rel_elem(host(direct), ref) <= TOL
        same_bits = np.array_equal(g, host(direct))
        assert same_bits == (not fast), "the guard decides which kernel writes the result"
        if fast:  # the bound the guard promises (1.2e-11) against the direct kernel
            assert np.max(np.abs(g - host(direct)) / np.maximum(np.abs(host(direct)), 1e-300)) <= 2e-11
    # odd leading dimension (scalar stores) and a NaN coordinate: its row is NaN, the others are untouched
    x1, x2 = rng.standard_normal((m1, d)), rng.standard_normal((m2, d))
    ll = np.full(d, np.log(1.0 / d))
    out = torch.full((m1, m2 + 1), -7.0, dtype=torch.float64, device="cuda")
    lib = S._lib.load()
    ws = torch.zeros(lib.gvi_ard_gram_workspace_bytes(m1, m2, d) // 8, dtype=torch.float64, device="cuda")
    assert ws.numel() > 0
    a1, a2, als, all_ = dev(x1), dev(x2), dev([ls]), dev(ll)
    rc = lib.gvi_ard_gram_f64(a1.data_ptr(), m1, a2.data_ptr(), m2, d, als.data_ptr(), all_.data_ptr(), d, out.data_ptr(),
                              m2 + 1, ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    o = host(out)
    assert rel_elem(o[:, :m2], O.ard_gram(ls, ll, x1, x2)) <= TOL and np.all(o[:, m2] == -7.0)
    x1[3, d - 1] = np.nan
    g = host(k.calculate_gram(k.Parameters.construct(log_scaling=ls, log_lengthscales=ll), x1, x2))
    assert np.all(np.isnan(g[3])) and not np.any(np.isnan(np.delete(g, 3, axis=0)))


def test_gram_shape_asserts(S):
    k = S.ARDKernel(number_of_dimensions=3)
    p = k.Parameters.construct(log_scaling=0.0, log_lengthscales=np.zeros(3))
    with pytest.raises(AssertionError):
        k.calculate_gram(p, np.zeros((2, 3)), np.zeros((2, 4)))
    with pytest.raises(AssertionError):
        k.calculate_gram(p, np.zeros((2, 3)), np.zeros((3, 3)), full_covariance=False)
    with pytest.raises(AssertionError):
        k.calculate_gram({"log_scaling": 0.0, "log_lengthscales": np.zeros(2)}, np.zeros((2, 3)))


# ---------------------------------------------------------------------------------------------------------------
# a4 + Cholesky
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m", [1, 31, 32, 33, 200, 1000])
def test_add_diag_and_cholesky(S, m):
    rng = np.random.default_rng(m)
    b = rng.standard_normal((m, m + 5))
    a = b @ b.T / m + np.eye(m)
    for is_abs in (True, False):
        ad = dev(a)
        S._lib_proxy.add_diagonal_regulariser(ad, 0.25, is_abs)
        assert rel_elem(host(ad), O.add_diagonal_regulariser(a, 0.25, is_abs)) <= 1e-14
    ad = dev(a)
    info = S._lib_proxy.chol_factor(ad)
    assert int(info.item()) == 0
    lref = np.linalg.cholesky(a)
    assert np.max(np.abs(np.tril(host(ad)) - lref)) <= 1e-12 * np.max(np.abs(lref))
    # a matrix that is not positive definite reports the failing column (and yields NaN, like the reference)
    bad = np.eye(m)
    bad[m // 2, m // 2] = -1.0
    info = S._lib_proxy.chol_factor(dev(bad))
    assert int(info.item()) == m // 2 + 1


# ---------------------------------------------------------------------------------------------------------------
# a5-a7 predictive
# ---------------------------------------------------------------------------------------------------------------
def make_problem(n, m, d, seed, ll_base=0.0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    z = rng.standard_normal((m, d))
    w = rng.standard_normal(d)
    y = np.sin(x @ w) + 0.1 * rng.standard_normal(n)
    yz = np.sin(z @ w) + 0.1 * rng.standard_normal(m)
    ls_q, ll_q = 0.1, ll_base + 0.1 * rng.standard_normal(d)
    ls_p, ll_p = -0.2, ll_base + 0.1 * rng.standard_normal(d)
    m_q = 0.3 * np.tanh(x @ rng.standard_normal(d))
    return dict(x=x, z=z, y=y, yz=yz, ls_q=ls_q, ll_q=ll_q, ls_p=ls_p, ll_p=ll_p, m_q=m_q)


def oracle_predictions(pr, lam=1e-5, prior_const=0.25, log_noise=0.0):
    gq = lambda a, b: O.ard_gram(pr["ls_q"], pr["ll_q"], a, b)
    dq = lambda a: O.ard_gram_diag(pr["ls_q"], pr["ll_q"], a)
    gp = lambda a, b: O.ard_gram(pr["ls_p"], pr["ll_p"], a, b)
    dp = lambda a: O.ard_gram_diag(pr["ls_p"], pr["ll_p"], a)
    n = pr["x"].shape[0]
    v_q = O.sparse_posterior_diag(gq, dq, pr["z"], pr["x"], lam, False)
    m_p, v_p = O.exact_gp_posterior(gp, dp, np.full(n, prior_const), pr["z"], pr["yz"], log_noise, pr["x"])
    return v_q, m_p, v_p


def build_models(S, pr, lam=1e-5, prior_const=0.25, log_noise=0.0):
    d = pr["x"].shape[1]
    kernel = S.ARDKernel(number_of_dimensions=d)
    mq_table = dev(pr["m_q"])
    x_dev = dev(pr["x"])
    # the approximate mean is "any function evaluated by the host framework": here a lookup of precomputed values
    approx_mean = S.CustomMean(mean_function=lambda params, x: params[: x.shape[0]])
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=pr["z"], diagonal_regularisation=lam)
    approx_gp = S.ApproximateGPRegression(mean=approx_mean, kernel=sparse)
    approx_params = approx_gp.Parameters.construct(
        mean=approx_mean.Parameters.construct(custom=mq_table),
        kernel=sparse.Parameters.construct(
            base_kernel=kernel.Parameters.construct(log_scaling=pr["ls_q"], log_lengthscales=pr["ll_q"])))
    const_mean = S.ConstantMean()
    exact_gp = S.GPRegression(mean=const_mean, kernel=kernel, x=pr["z"], y=pr["yz"])
    exact_params = exact_gp.Parameters.construct(
        log_observation_noise=log_noise, mean=const_mean.Parameters.construct(constant=prior_const),
        kernel=kernel.Parameters.construct(log_scaling=pr["ls_p"], log_lengthscales=pr["ll_p"]))
    return approx_gp, approx_params, exact_gp, exact_params, x_dev


@pytest.mark.parametrize("n,m,d", [(1, 2, 1), (23, 10, 1), (1000, 100, 3), (2500, 333, 8), (777, 1024, 8),
                                   (500, 1500, 5), (300, 2048, 16), (200, 64, 32), (260, 4096, 8), (4000, 1153, 2)])
def test_predictive_parity(S, n, m, d):
    pr = make_problem(n, m, d, seed=n + m)
    v_q_ref, m_p_ref, v_p_ref = oracle_predictions(pr)
    approx_gp, approx_params, exact_gp, exact_params, x_dev = build_models(S, pr)
    q = approx_gp.predict_probability(approx_params, x_dev)
    assert q.mean.shape == (n,) and q.covariance.shape == (n,)
    assert rel_norm(host(q.covariance), v_q_ref) <= TOL
    assert rel_elem(host(q.mean), pr["m_q"]) <= 1e-15
    p = exact_gp.predict_probability(exact_params, x_dev)
    assert p.mean.shape == (1, n) and p.covariance.shape == (n,)  # [1,N] mean quirk, base.py:410-412
    assert rel_norm(host(p.mean).reshape(-1), m_p_ref) <= TOL
    assert rel_norm(host(p.covariance), v_p_ref) <= TOL
    assert np.all(host(q.covariance) > -1e-12) and np.all(host(p.covariance) > 0)


@pytest.mark.parametrize("n,m,d,ll_shift", [(4001, 200, 8, 0.0), (3000, 1100, 5, 0.0), (2500, 96, 16, 0.0), (1500, 64, 27, 0.0),
                                           (2000, 150, 8, 7.0)])
def test_tile_kernel_guarded_tensor_pipe_kernel_evaluations(S, n, m, d, ll_shift):
    """Phase 1 of the tile kernel evaluates k(x_i, z_j) through |u|^2/2 + |v|^2/2 - u.v on the tensor pipe for every point
    whose guard (|u_i| + |v|max)^2 <= 512 holds and by direct differences for the others (csrc/predict.cu).  Checked:
    parity with the oracle for a data set with OUTLIERS (some points 40 lengthscales away: mixed tiles, the fix-up path),
    for short lengthscales (ll_shift = 7: no point passes, whole tiles take the direct code), and that the value a point
    gets does not depend on the tile it falls into - the same points predicted in a different order / batch give the
    same bits."""
    rng = np.random.default_rng(n + d)
    pr = make_problem(n, m, d, seed=n + m)
    out_rows = rng.choice(n, size=37, replace=False)
    pr["x"][out_rows] *= 40.0
    pr["ll_q"] = pr["ll_q"] + ll_shift
    pr["ll_p"] = pr["ll_p"] + ll_shift
    v_q_ref, m_p_ref, v_p_ref = oracle_predictions(pr)
    approx_gp, approx_params, exact_gp, exact_params, x_dev = build_models(S, pr)
    q = approx_gp.predict_probability(approx_params, x_dev)
    p = exact_gp.predict_probability(exact_params, x_dev)
    assert rel_norm(host(q.covariance), v_q_ref) <= TOL
    assert rel_norm(host(p.mean).reshape(-1), m_p_ref) <= TOL and rel_norm(host(p.covariance), v_p_ref) <= TOL
    # tile independence: reversed order and a 1000-point slice (8- / 16-point tiles) reproduce the same bits per point
    rev = torch.flip(x_dev, dims=[0]).contiguous()
    p_rev = exact_gp.predict_probability(exact_params, rev)
    assert np.array_equal(host(p_rev.covariance)[::-1], host(p.covariance))
    assert np.array_equal(host(p_rev.mean).reshape(-1)[::-1], host(p.mean).reshape(-1))
    sl = x_dev[777:1777].contiguous()
    p_sl = exact_gp.predict_probability(exact_params, sl)
    assert np.array_equal(host(p_sl.covariance), host(p.covariance)[777:1777])


@pytest.mark.parametrize("n,m,d", [(5000, 300, 33), (700, 1300, 784), (30000, 64, 40)])
def test_predictive_parity_wide_inputs(S, n, m, d):
    """D > 32 (config #4's MNIST-like width): K(x,Z) rows come from the tensor-pipe Gram kernel in chunks of <= 28 416
    points (n = 30 000: two chunks, the second ragged)."""
    rng = np.random.default_rng(d)
    x, z = rng.uniform(0, 1, (n, d)), rng.uniform(0, 1, (m, d))
    pr = dict(x=x, z=z, y=rng.standard_normal(n), yz=rng.standard_normal(m), ls_q=0.1,
              ll_q=np.log(40.0 / d) + 0.1 * rng.standard_normal(d), ls_p=-0.2,
              ll_p=np.log(40.0 / d) + 0.1 * rng.standard_normal(d), m_q=0.3 * rng.standard_normal(n))
    v_q_ref, m_p_ref, v_p_ref = oracle_predictions(pr)
    approx_gp, approx_params, exact_gp, exact_params, x_dev = build_models(S, pr)
    q = approx_gp.predict_probability(approx_params, x_dev)
    p = exact_gp.predict_probability(exact_params, x_dev)
    assert rel_norm(host(q.covariance), v_q_ref) <= TOL
    assert rel_norm(host(p.mean).reshape(-1), m_p_ref) <= TOL and rel_norm(host(p.covariance), v_p_ref) <= TOL
    # the objective is ONE call of gvi_fused_step_f64 for wide inputs too (r02): K rows of q and p from the tensor-pipe
    # Gram kernel through two rings, <= 28 416 points per chunk
    risk = S.NegativeLogLikelihood(gp=approx_gp)
    reg = S.projected.ProjectedKLRegularisation(gp=approx_gp, regulariser=exact_gp, regulariser_parameters=exact_params,
                                                mode="posterior")
    gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=risk)
    assert gvi._fusable()
    ref = O.gaussian_nll(pr["y"], pr["m_q"], v_q_ref) + O.projected_regularisation("kl", m_p_ref, v_p_ref, pr["m_q"],
                                                                                  v_q_ref)
    fused = float(gvi.calculate_loss(approx_params, x_dev, dev(pr["y"])))
    assert abs(fused - ref) <= TOL * abs(ref)
    unfused = float(gvi._calculate_loss_term_by_term(approx_params, x_dev, dev(pr["y"])))
    assert abs(fused - unfused) <= 1e-12 * abs(ref)
    # prior-mode regulariser (no second pass: m_p, v_p arrays) through the same wide route
    reg_prior = S.projected.ProjectedKLRegularisation(gp=approx_gp, regulariser=exact_gp,
                                                      regulariser_parameters=exact_params, mode="prior")
    gvi_prior = S.GeneralisedVariationalInference(regularisation=reg_prior, empirical_risk=risk)
    assert gvi_prior._fusable()
    a = float(gvi_prior.calculate_loss(approx_params, x_dev, dev(pr["y"])))
    b = float(gvi_prior._calculate_loss_term_by_term(approx_params, x_dev, dev(pr["y"])))
    assert abs(a - b) <= 1e-12 * abs(b)


def test_sparse_posterior_full_gram_and_identity_golden(S):
    # reference tests/kernels/test_sparse_posterior_kernels.py:15-64 uses an identity mock kernel: K_ZZ = I,
    # k(x_i, z_j) = delta_ij -> diag 1 - 1/(1 + 1e-5) = 9.9999e-06, off-diagonal 0.  An ARD kernel with a huge
    # inverse lengthscale evaluated at x = Z is that identity kernel.
    z = np.array([[5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    k = S.ARDKernel(number_of_dimensions=3)
    sp = S.SparsePosteriorKernel(base_kernel=k, inducing_points=z, diagonal_regularisation=1e-5,
                                 is_diagonal_regularisation_absolute_scale=False)
    p = sp.generate_parameters({"base_kernel": {"log_scaling": 0.0, "log_lengthscales": np.full(3, 20.0)}})
    full = host(sp.calculate_gram(p, z, z, full_covariance=True))
    assert np.allclose(full, np.array([[9.9999e-06, 0.0], [0.0, 9.9999e-06]]), rtol=1e-5, atol=1e-8)
    diag = host(sp.calculate_gram(p, z, z, full_covariance=False))
    assert np.allclose(diag, [9.9999e-06, 9.9999e-06], rtol=1e-5, atol=1e-8)
    # full Gram against the oracle on a random problem
    pr = make_problem(50, 20, 3, seed=9)
    sp = S.SparsePosteriorKernel(base_kernel=k, inducing_points=pr["z"])
    p = sp.generate_parameters({"base_kernel": {"log_scaling": pr["ls_q"], "log_lengthscales": pr["ll_q"]}})
    x2 = pr["x"][:17]
    g = host(sp.calculate_gram(p, pr["x"], x2))
    ref = O.sparse_posterior_gram(lambda a, b: O.ard_gram(pr["ls_q"], pr["ll_q"], a, b), pr["z"], pr["x"], x2)
    assert rel_norm(g, ref) <= TOL


# ---------------------------------------------------------------------------------------------------------------
# reference goldens through the real product path: an all-ones Gram is ARD with exp(log_ls) = 0
# ---------------------------------------------------------------------------------------------------------------
X_TRAIN = np.array([[1.0, 3.0, 2.0], [1.5, 1.5, 9.5]])
Y_TRAIN = np.array([1.0, 1.5])
X2 = np.array([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5]])
ONES = {"log_scaling": 0.0, "log_lengthscales": np.full(3, -np.inf)}


def ones_kernel_models(S):
    k = S.ARDKernel(number_of_dimensions=3)
    mean = S.ConstantMean()
    exact = S.GPRegression(mean=mean, kernel=k, x=X_TRAIN, y=Y_TRAIN)
    approx = S.ApproximateGPRegression(mean=mean, kernel=k)
    params = {"log_observation_noise": np.log(1.0), "mean": {"constant": 1.0}, "kernel": ONES}
    return exact, approx, params


def test_reference_goldens_exact_and_approximate_gp(S):
    exact, approx, params = ones_kernel_models(S)
    p = exact.predict_probability(params, X2)
    # tests/gps/test_regression.py:14-56 pins these with array_equal; the GPU inverse-Cholesky route is allowed 2 ulp
    assert np.allclose(host(p.mean), 1.8333333333333335, rtol=5e-16, atol=0)
    assert np.allclose(host(p.covariance), 0.33333333333333326, rtol=1e-15, atol=0)
    q = approx.predict_probability(params, X2)  # noise dropped: tests/gps/test_regression.py:203-249
    assert np.array_equal(host(q.mean), np.ones(2)) and np.array_equal(host(q.covariance), np.ones(2))


def test_reference_goldens_nll(S):
    exact, approx, params = ones_kernel_models(S)
    nll = S.NegativeLogLikelihood(gp=exact)
    assert np.isclose(float(nll.calculate_empirical_risk(params, X2, np.ones(2))), 1.4112988, rtol=1e-5)
    assert np.isclose(float(nll.calculate_empirical_risk(params, np.tile(X2, (6, 1)), np.ones(12))), 1.4112988,
                      rtol=1e-5)
    nll_q = S.NegativeLogLikelihood(gp=approx)
    assert np.isclose(float(nll_q.calculate_empirical_risk(params, X2, np.array([1.2, 4.2]))), 3.48893853,
                      rtol=1e-5)


REG_CASES = [
    ("ProjectedBhattacharyyaRegularisation", {}, 0.13702468),
    ("ProjectedGaussianWassersteinRegularisation", {}, 0.87307724),
    ("ProjectedHellingerRegularisation", {}, 0.18301035),
    ("ProjectedKLRegularisation", {}, 0.56319503),
    ("ProjectedRenyiRegularisation", {"alpha": 0.5}, 0.4042577),
]


@pytest.mark.parametrize("cls,kw,golden", REG_CASES)
def test_reference_goldens_projected_regularisers(S, cls, kw, golden):
    # tests/regularisations/projected/test_projected_*.py:120-176: p = (1.8333, 1/3) posterior, q = (1, 1)
    exact, approx, params = ones_kernel_models(S)
    reg = getattr(S.projected, cls)(gp=approx, regulariser=exact, regulariser_parameters=params, mode="posterior",
                                    **kw)
    assert np.isclose(float(reg.calculate_regularisation(params, X2)), golden, rtol=1e-5)
    same = getattr(S.projected, cls)(gp=approx, regulariser=approx, regulariser_parameters=params, mode="prior", **kw)
    assert abs(float(same.calculate_regularisation(params, X2))) <= 1e-12


def test_gvi_is_risk_plus_regularisation(S):
    exact, approx, params = ones_kernel_models(S)
    risk = S.NegativeLogLikelihood(gp=approx)
    reg = S.projected.ProjectedKLRegularisation(gp=approx, regulariser=exact, regulariser_parameters=params,
                                                mode="posterior")
    gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=risk)
    y = np.array([1.2, 4.2])
    total = float(gvi.calculate_loss(params, X2, y))
    assert np.isclose(total, 3.48893853 + 0.56319503, rtol=1e-5)
    assert np.isclose(total, float(risk.calculate_empirical_risk(params, X2, y))
                      + float(reg.calculate_regularisation(params, X2)), rtol=1e-14)


# ---------------------------------------------------------------------------------------------------------------
# a9-a12 risks, regularisers, fused step
# ---------------------------------------------------------------------------------------------------------------
ALL_REGS = [("bhattacharyya", "ProjectedBhattacharyyaRegularisation"),
            ("gaussian_wasserstein", "ProjectedGaussianWassersteinRegularisation"),
            ("hellinger", "ProjectedHellingerRegularisation"), ("kl", "ProjectedKLRegularisation"),
            ("renyi", "ProjectedRenyiRegularisation")]


def oracle_reg(kind, m_p, v_p, m_q, v_q, alpha=0.5):
    if kind == "gwi_no_eig":
        return O.gwi_no_eig(m_p, v_p, m_q, v_q)
    if kind == "squared_difference":
        return O.gaussian_squared_difference(m_p, v_p, m_q, v_q)
    return O.projected_regularisation(kind, m_p, v_p, m_q, v_q, alpha)


@pytest.mark.parametrize("n", [1, 2, 3, 1001, 300000])
def test_risk_kernel_all_kinds(S, n):
    rng = np.random.default_rng(n)
    y, m_q, m_p = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    v_q, v_p = rng.uniform(0.05, 2.0, n), rng.uniform(0.05, 2.0, n)
    for kind, code in S._lib.REG_KINDS.items():
        if kind == "none":
            continue
        alpha = 0.3 if kind == "renyi" else 0.5
        out = host(S._lib_proxy.risk(code, alpha, dev(y), dev(m_q), dev(v_q), dev(m_p), dev(v_p)))
        assert out[2] == n
        assert abs(out[0] / n - O.gaussian_nll(y, m_q, v_q)) <= TOL * abs(O.gaussian_nll(y, m_q, v_q))
        ref = oracle_reg(kind, m_p, v_p, m_q, v_q, alpha)
        assert abs(out[1] / n - ref) <= TOL * abs(ref), kind
    # unaligned views take the scalar path
    out = host(S._lib_proxy.risk(4, 0.5, dev(y)[1:].contiguous() if n > 1 else dev(y), *(
        [t[1:] for t in (dev(m_q), dev(v_q), dev(m_p), dev(v_p))] if n > 1 else
        [dev(m_q), dev(v_q), dev(m_p), dev(v_p)])))
    sl = slice(1, None) if n > 1 else slice(None)
    ref = O.projected_regularisation("kl", m_p[sl], v_p[sl], m_q[sl], v_q[sl])
    assert abs(out[1] / out[2] - ref) <= TOL * abs(ref)


@pytest.mark.parametrize("n,m,d", [(50, 10, 1), (3001, 257, 8)])
def test_fused_step_matches_oracle_and_unfused(S, n, m, d):
    pr = make_problem(n, m, d, seed=n)
    v_q_ref, m_p_ref, v_p_ref = oracle_predictions(pr)
    approx_gp, approx_params, exact_gp, exact_params, x_dev = build_models(S, pr)
    nll_ref = O.gaussian_nll(pr["y"], pr["m_q"], v_q_ref)
    risk = S.NegativeLogLikelihood(gp=approx_gp)
    y_dev = dev(pr["y"])
    regs = [(kind, getattr(S.projected, cls)) for kind, cls in ALL_REGS]
    regs += [("gwi_no_eig", S.GaussianWassersteinRegularisation)]
    regs += [("squared_difference", lambda **kw: S.GaussianSquaredDifferenceRegularisation(full_covariance=False, **kw))]
    for mode in ("posterior", "prior"):
        for kind, ctor in regs:
            reg = ctor(gp=approx_gp, regulariser=exact_gp, regulariser_parameters=exact_params, mode=mode)
            gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=risk)
            assert gvi._fusable()
            loss = float(gvi.calculate_loss(approx_params, x_dev, y_dev))
            if mode == "posterior":
                reg_ref = oracle_reg(kind, m_p_ref, v_p_ref, pr["m_q"], v_q_ref)
            else:
                m_pr, v_pr = O.exact_gp_prior(lambda a: O.ard_gram_diag(pr["ls_p"], pr["ll_p"], a), np.full(n, 0.25),
                                              0.0, pr["x"])
                reg_ref = oracle_reg(kind, m_pr, v_pr, pr["m_q"], v_q_ref)
            assert abs(loss - (nll_ref + reg_ref)) <= TOL * abs(nll_ref + reg_ref), (mode, kind)
            unfused = float(risk.calculate_empirical_risk(approx_params, x_dev, y_dev)) + float(
                reg.calculate_regularisation(approx_params, x_dev))
            assert abs(loss - unfused) <= 1e-12 * abs(unfused), (mode, kind)


def test_unsupported_variants_raise(S):
    pr = make_problem(10, 4, 2, seed=1)
    approx_gp, approx_params, exact_gp, exact_params, _ = build_models(S, pr)
    with pytest.raises(NotImplementedError):
        S.GaussianWassersteinRegularisation(gp=approx_gp, regulariser=exact_gp, regulariser_parameters=exact_params,
                                            include_eigendecomposition=True)
    with pytest.raises(NotImplementedError):
        S.GaussianSquaredDifferenceRegularisation(gp=approx_gp, regulariser=exact_gp,
                                                  regulariser_parameters=exact_params)
    with pytest.raises(AssertionError):  # not a kernel at all (any KernelBase is accepted: tests/test_gpu_other_kernels.py)
        S.SparsePosteriorKernel(base_kernel=object(), inducing_points=pr["z"])


# ---------------------------------------------------------------------------------------------------------------
# golden fixtures (configs #1 and #2), end to end through the README-style API
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,m", [("config1_readme", 10), ("config2_toy_curves", 100)])
def test_golden_fixture_end_to_end(S, name, m):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    x, y = g["x"], g["y"]
    d = x.shape[1]
    kernel = S.ARDKernel(number_of_dimensions=d)
    kp = kernel.Parameters.construct(log_scaling=float(g["log_scaling"]), log_lengthscales=g["log_ls"])
    sel = S.ConditionalVarianceInducingPointsSelector()
    z, idx = sel.compute_inducing_points(key=S.jax_prng.PRNGKey(0), training_inputs=x,
                                         number_of_inducing_points=m, kernel=kernel, kernel_parameters=kp)
    assert np.array_equal(host(idx), g["idx"])          # indices bit-exact
    assert np.array_equal(host(z), x[g["idx"]])
    assert rel_elem(host(kernel.calculate_gram(kp, x[:64], z)), g["gram_xz"]) <= TOL
    yz = y[g["idx"]]
    mean0 = S.ConstantMean()
    exact = S.GPRegression(mean=mean0, kernel=kernel, x=z, y=yz)
    exact_params = exact.Parameters.construct(log_observation_noise=0.0,
                                              mean=mean0.Parameters.construct(constant=0.0), kernel=kp)
    p = exact.predict_probability(exact_params, x)
    assert rel_norm(host(p.mean).reshape(-1), g["m_p"]) <= TOL and rel_norm(host(p.covariance), g["v_p"]) <= TOL
    mq = dev(g["m_q"])
    amean = S.CustomMean(mean_function=lambda params, xx: params)
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=z, diagonal_regularisation=float(g["lam"]))
    approx = S.ApproximateGPRegression(mean=amean, kernel=sparse)
    ap = approx.Parameters.construct(
        mean=amean.Parameters.construct(custom=mq),
        kernel=sparse.Parameters.construct(
            base_kernel=kernel.Parameters.construct(log_scaling=float(g["ls_q"]), log_lengthscales=g["ll_q"])))
    q = approx.predict_probability(ap, x)
    assert rel_norm(host(q.covariance), g["v_q"]) <= TOL
    risk = S.NegativeLogLikelihood(gp=approx)
    nll = float(risk.calculate_empirical_risk(ap, x, y))
    assert abs(nll - float(g["nll"])) <= TOL * abs(float(g["nll"]))
    for kind, cls in ALL_REGS:
        reg = getattr(S.projected, cls)(gp=approx, regulariser=exact, regulariser_parameters=exact_params,
                                        mode="posterior")
        gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=risk)
        ref = float(g["nll"]) + float(g["reg_" + kind])
        assert abs(float(gvi.calculate_loss(ap, x, y)) - ref) <= TOL * abs(ref), kind
    gw = S.GaussianWassersteinRegularisation(gp=approx, regulariser=exact, regulariser_parameters=exact_params,
                                             mode="posterior")
    assert abs(float(gw.calculate_regularisation(ap, x)) - float(g["reg_gwi_no_eig"])) <= TOL * float(
        g["reg_gwi_no_eig"])


# ---------------------------------------------------------------------------------------------------------------
# a13 selector
# ---------------------------------------------------------------------------------------------------------------
def test_selector_reference_pins(S):
    # reference tests/test_inducing_points_selection.py:116-182: PRNGKey(0), all-ones Gram -> indices [2, 1]
    k = S.ARDKernel(number_of_dimensions=3)
    sel = S.ConditionalVarianceInducingPointsSelector()
    x4 = np.array([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5], [4.5, 2.5, 3.5], [5.5, 6.5, 7.5]])
    z, idx = sel.compute_inducing_points(S.jax_prng.PRNGKey(0), x4, 2, k, ONES)
    assert host(idx).tolist() == [2, 1]
    assert np.array_equal(host(z), np.array([[4.5, 2.5, 3.5], [1.5, 2.5, 3.5]]))
    k784 = S.ARDKernel(number_of_dimensions=784)
    img = np.arange(3 * 784, dtype=float).reshape(3, 28, 28, 1)
    z, idx = sel.compute_inducing_points(S.jax_prng.PRNGKey(0), img, 2, k784,
                                         {"log_scaling": 0.0, "log_lengthscales": np.full(784, -np.inf)})
    assert host(idx).tolist() == [2, 1] and tuple(z.shape) == (2, 28, 28, 1)
    rnd = S.RandomInducingPointsSelector()
    assert host(rnd.compute_inducing_points(S.jax_prng.PRNGKey(0), x4, 2)[1]).tolist() == [1, 2]
    assert host(rnd.compute_inducing_points(S.jax_prng.PRNGKey(0), img, 2)[1]).tolist() == [0, 2]


@pytest.mark.parametrize("n,m,d,seed", [(2, 2, 1, 0), (500, 40, 1, 1), (5000, 128, 3, 2), (20000, 200, 8, 3)])
def test_selector_parity_bit_exact(S, n, m, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    # Bit-exact pivots are only defined while the residual variances stay far above rounding noise: once
    # tr(K - Q) ~ 1e-12 the arg-max order is decided by the last bits of exp()/dot and differs between ANY two
    # implementations (it would between JAX-CPU and JAX-GPU too).  In 1-D a short lengthscale keeps 40 pivots
    # well-posed.
    ls, ll = 0.2, rng.normal(4.0 if d == 1 else -0.5, 0.3, d)
    k = S.ARDKernel(number_of_dimensions=d)
    kp = k.Parameters.construct(log_scaling=ls, log_lengthscales=ll)
    key = S.jax_prng.PRNGKey(seed)
    z, idx = S.ConditionalVarianceInducingPointsSelector().compute_inducing_points(key, x, m, k, kp)
    _, sub = OP.split(OP.prng_key(seed))
    perm = OP.permutation(sub, n)
    piv = O.conditional_variance_select(x[perm], m, lambda a, b: O.ard_gram(ls, ll, a, b),
                                        lambda a: O.ard_gram_diag(ls, ll, a), use_argsort=(n <= 5000))
    assert np.array_equal(host(idx), perm[piv])
    assert np.array_equal(host(z), x[perm[piv]])
    assert len(set(host(idx).tolist())) == m


def test_selector_duplicates_and_threshold(S):
    # Triplicated rows: the three copies of a point carry bit-identical d, so every pivot is a three-way exact tie
    # decided by the highest-index rule (reference lines 161-164); picking one copy zeroes the others.  Only the 6
    # distinct points are well-posed pivots (afterwards every d is rounding noise ~1e-12).
    rng = np.random.default_rng(4)
    base = rng.standard_normal((6, 2))
    x = np.concatenate([base, base, base])
    k = S.ARDKernel(number_of_dimensions=2)
    kp = k.Parameters.construct(log_scaling=0.0, log_lengthscales=np.zeros(2))
    gram = lambda a, b: O.ard_gram(0.0, np.zeros(2), a, b)
    diag = lambda a: O.ard_gram_diag(0.0, np.zeros(2), a)
    z, idx = S.ConditionalVarianceInducingPointsSelector().compute_inducing_points(S.jax_prng.PRNGKey(3), x, 6, k, kp)
    _, sub = OP.split(OP.prng_key(3))
    perm = OP.permutation(sub, 18)
    piv = O.conditional_variance_select(x[perm], 6, gram, diag)
    assert np.array_equal(host(idx), perm[piv])
    assert sorted((host(idx) % 6).tolist()) == list(range(6))
    # threshold: tr(K - Q) drops below 1e-3 at step m = 5 (all distinct points used) -> indices[:5], with a warning
    with pytest.warns(UserWarning, match="Terminating selection"):
        z, idx_t = S.ConditionalVarianceInducingPointsSelector(threshold=1e-3).compute_inducing_points(
            S.jax_prng.PRNGKey(3), x, 10, k, kp)
    import warnings as _w
    with _w.catch_warnings():
        _w.simplefilter("ignore")
        piv_t = O.conditional_variance_select(x[perm], 10, gram, diag, threshold=1e-3)
    assert len(piv_t) == 5 and np.array_equal(host(idx_t), perm[piv_t])
    assert tuple(z.shape) == (5, 2)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json config #3 at full size: size-independent properties + sampled oracle rows
# ---------------------------------------------------------------------------------------------------------------
def test_config3_full_size_properties(S):
    n, m, d = 1_000_000, 1024, 8
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, d))
    w = rng.standard_normal(d)
    y = np.sin(x @ w) + 0.1 * rng.standard_normal(n)
    zi = rng.permutation(n)[:m]
    pr = dict(x=x, z=x[zi], y=y, yz=y[zi], ls_q=0.0, ll_q=np.full(d, np.log(1.0 / d)), ls_p=0.0,
              ll_p=np.full(d, np.log(1.0 / d)), m_q=0.3 * np.tanh(x @ rng.standard_normal(d)))
    approx_gp, approx_params, exact_gp, exact_params, x_dev = build_models(S, pr)
    y_dev = dev(y)
    risk = S.NegativeLogLikelihood(gp=approx_gp)
    reg = S.projected.ProjectedRenyiRegularisation(gp=approx_gp, regulariser=exact_gp,
                                                   regulariser_parameters=exact_params, mode="posterior", alpha=0.5)
    gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=risk)
    full = host(gvi.loss_sums(approx_params, x_dev, y_dev))
    # (1) shard additivity: the loss sums of two contiguous shards add up to the full-batch sums
    approx_mean_tbl = approx_params.mean.custom
    halves = []
    for lo, hi in ((0, 400_003), (400_003, n)):
        approx_params.mean.custom = approx_mean_tbl[lo:hi].contiguous()
        halves.append(host(gvi.loss_sums(approx_params, x_dev[lo:hi], y_dev[lo:hi])))
    approx_params.mean.custom = approx_mean_tbl
    both = halves[0] + halves[1]
    assert both[2] == n == full[2]
    assert np.all(np.abs(both[:2] - full[:2]) <= 1e-12 * np.abs(full[:2]))
    # (2) run-to-run reproducibility (fixed-order reductions)
    assert np.array_equal(full, host(gvi.loss_sums(approx_params, x_dev, y_dev)))
    # (3) 0 <= var <= k(x,x); homogeneity: with relative jitter var scales exactly with exp(2 log_scaling)
    q = approx_gp.predict_probability(approx_params, x_dev)
    v = host(q.covariance)
    assert v.min() >= -1e-12 and v.max() <= 1.0 + 1e-12
    approx_params.kernel.base_kernel.log_scaling = np.log(3.0)
    v3 = host(approx_gp.predict_probability(approx_params, x_dev).covariance)
    approx_params.kernel.base_kernel.log_scaling = 0.0
    assert np.max(np.abs(v3 - 9.0 * v)) <= TOL * 9.0 * v.max()
    # (4) 2000 sampled rows against the oracle (the oracle materialises K(X_s, Z) like the reference)
    rows = np.sort(rng.permutation(n)[:2000])
    prs = dict(pr, x=x[rows], y=y[rows], m_q=pr["m_q"][rows])
    v_q_ref, m_p_ref, v_p_ref = oracle_predictions(prs)
    p = exact_gp.predict_probability(exact_params, x_dev)
    assert rel_norm(v[rows], v_q_ref) <= TOL
    assert rel_norm(host(p.mean).reshape(-1)[rows], m_p_ref) <= TOL
    assert rel_norm(host(p.covariance)[rows], v_p_ref) <= TOL


def test_selector_full_size_properties(S):
    n, m, d = 1_000_000, 48, 8
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, d))
    ls, ll = 0.0, np.full(d, np.log(1.0 / d))
    k = S.ARDKernel(number_of_dimensions=d)
    kp = k.Parameters.construct(log_scaling=ls, log_lengthscales=ll)
    z, idx = S.ConditionalVarianceInducingPointsSelector().compute_inducing_points(S.jax_prng.PRNGKey(11), x, m, k, kp)
    idx = host(idx)
    assert len(set(idx.tolist())) == m and idx.min() >= 0 and idx.max() < n
    assert np.array_equal(host(z), x[idx])
    _, sub = OP.split(OP.prng_key(11))
    perm = OP.permutation(sub, n)
    piv = O.conditional_variance_select(x[perm], m, lambda a, b: O.ard_gram(ls, ll, a, b),
                                        lambda a: O.ard_gram_diag(ls, ll, a), use_argsort=False)
    assert np.array_equal(idx, perm[piv])


# ---------------------------------------------------------------------------------------------------------------
# gradients (jax.grad of calculate_loss in the reference; oracle = torch-f64 autograd twin), tolerance 1e-8
# ---------------------------------------------------------------------------------------------------------------
GRAD_TOL = 1e-8


@pytest.mark.parametrize("n,m,d,lam,is_abs", [(60, 12, 1, 1e-5, False), (700, 65, 3, 1e-5, False),
                                              (1500, 200, 8, 1e-3, True), (900, 129, 8, 1e-5, False)])
def test_gradients_match_autograd_twin(S, n, m, d, lam, is_abs):
    from oracle import torch_twin as TT

    # 1-D problems with a long lengthscale drive v_q to ~1e-9 and the NLL to ~1e5: there the loss itself is only
    # defined to ~1e-9 relative in ANY implementation (conditioning), so 1-D uses a short one
    pr = make_problem(n, m, d, seed=7 * n + m, ll_base=3.0 if d == 1 else 0.5)
    d_ = pr["x"].shape[1]
    kernel = S.ARDKernel(number_of_dimensions=d_)
    mq_table = dev(pr["m_q"])
    approx_mean = S.CustomMean(mean_function=lambda params, x: params[: x.shape[0]])
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=pr["z"], diagonal_regularisation=lam,
                                     is_diagonal_regularisation_absolute_scale=is_abs)
    approx_gp = S.ApproximateGPRegression(mean=approx_mean, kernel=sparse)
    approx_params = approx_gp.Parameters.construct(
        mean=approx_mean.Parameters.construct(custom=mq_table),
        kernel=sparse.Parameters.construct(
            base_kernel=kernel.Parameters.construct(log_scaling=pr["ls_q"], log_lengthscales=pr["ll_q"])))
    const_mean = S.ConstantMean()
    exact_gp = S.GPRegression(mean=const_mean, kernel=kernel, x=pr["z"], y=pr["yz"])
    exact_params = exact_gp.Parameters.construct(
        log_observation_noise=0.0, mean=const_mean.Parameters.construct(constant=0.25),
        kernel=kernel.Parameters.construct(log_scaling=pr["ls_p"], log_lengthscales=pr["ll_p"]))
    gp_ = lambda a, b: O.ard_gram(pr["ls_p"], pr["ll_p"], a, b)
    dp_ = lambda a: O.ard_gram_diag(pr["ls_p"], pr["ll_p"], a)
    m_p, v_p = O.exact_gp_posterior(gp_, dp_, np.full(n, 0.25), pr["z"], pr["yz"], 0.0, pr["x"])
    risk = S.NegativeLogLikelihood(gp=approx_gp)
    regs = [(kind, getattr(S.projected, cls)) for kind, cls in ALL_REGS]
    regs += [("gwi_no_eig", S.GaussianWassersteinRegularisation)]
    regs += [("squared_difference", lambda **kw: S.GaussianSquaredDifferenceRegularisation(full_covariance=False, **kw))]
    for kind, ctor in regs:
        reg = ctor(gp=approx_gp, regulariser=exact_gp, regulariser_parameters=exact_params, mode="posterior")
        gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=risk)
        loss, grads = gvi.calculate_loss_and_gradients(approx_params, dev(pr["x"]), dev(pr["y"]))
        ref_loss, g_ls, g_ll, g_m = TT.gvi_loss_and_grads(pr["x"], pr["y"], pr["z"], pr["m_q"], pr["ls_q"], pr["ll_q"],
                                                           m_p, v_p, kind, 0.5, lam, is_abs)
        assert abs(float(loss) - ref_loss) <= TOL * abs(ref_loss), kind
        assert abs(float(grads["log_scaling"]) - g_ls) <= GRAD_TOL * max(abs(g_ls), 1e-3), kind
        scale = max(np.max(np.abs(g_ll)), 1e-3)
        assert np.max(np.abs(host(grads["log_lengthscales"]) - g_ll)) <= GRAD_TOL * scale, kind
        assert np.max(np.abs(host(grads["mean_output"]) - g_m)) <= GRAD_TOL * max(np.max(np.abs(g_m)), 1e-3), kind
        # the forward-only launch and the gradient launch agree on the loss
        assert abs(float(loss) - float(gvi.calculate_loss(approx_params, dev(pr["x"]), dev(pr["y"])))) <= 1e-13 * abs(
            ref_loss)


@pytest.mark.parametrize("n,m,d,ll_base", [(400, 40, 1, 4.5), (3000, 512, 8, float(np.log(1.0 / 8)))])
def test_gradients_ill_conditioned_inducing_set(S, n, m, d, ll_base):
    """cond(A) ~ 1e5 (40 inducing points in 1-D) and ~1e7 (config #3 in miniature: 512 points in 8-D, lengthscale
    weights 1/D, relative jitter 1e-5).  S = sum_i g_i a_i a_i^T is accumulated from the a_i = A^-1 k_i (error
    ~cond * eps); the algebraically equal A^-1 (sum_i g_i k_i k_i^T) A^-1 used before differed from the autograd
    twin by 4e-6 on the first case (cond^2 * eps)."""
    from oracle import torch_twin as TT

    pr = make_problem(n, m, d, seed=3, ll_base=ll_base)
    approx_gp, approx_params, exact_gp, exact_params, x_dev = build_models(S, pr)
    _, m_p, v_p = oracle_predictions(pr)
    risk = S.NegativeLogLikelihood(gp=approx_gp)
    reg = S.projected.ProjectedRenyiRegularisation(gp=approx_gp, regulariser=exact_gp,
                                                   regulariser_parameters=exact_params, mode="posterior", alpha=0.5)
    gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=risk)
    loss, grads = gvi.calculate_loss_and_gradients(approx_params, x_dev, dev(pr["y"]))
    ref_loss, g_ls, g_ll, g_m = TT.gvi_loss_and_grads(pr["x"], pr["y"], pr["z"], pr["m_q"], pr["ls_q"], pr["ll_q"],
                                                       m_p, v_p, "renyi", 0.5, 1e-5, False)
    assert abs(float(loss) - ref_loss) <= TOL * abs(ref_loss)
    assert abs(float(grads["log_scaling"]) - g_ls) <= GRAD_TOL * max(abs(g_ls), 1e-3)
    scale = max(np.max(np.abs(g_ll)), 1e-3)
    assert np.max(np.abs(host(grads["log_lengthscales"]) - g_ll)) <= GRAD_TOL * scale


def test_gradient_step_in_several_chunks(S):
    """The gradient step walks the points in chunks (tile launch -> S accumulation per chunk).  With the chunk capped at
    one 3552-point wave, 8000 points take three chunks, the last one ragged (896 points: a partly filled 32-point
    group); the result must equal the one-chunk result and the autograd twin."""
    from oracle import torch_twin as TT

    pr = make_problem(8000, 65, 3, seed=11, ll_base=0.5)
    approx_gp, approx_params, exact_gp, exact_params, x_dev = build_models(S, pr)
    _, m_p, v_p = oracle_predictions(pr)
    reg = S.projected.ProjectedKLRegularisation(gp=approx_gp, regulariser=exact_gp,
                                                regulariser_parameters=exact_params, mode="posterior")
    gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=S.NegativeLogLikelihood(gp=approx_gp))
    loss1, g1 = gvi.calculate_loss_and_gradients(approx_params, x_dev, dev(pr["y"]))
    with S._lib.test_library() as lib:  # the hook lives in libgvigp_test.so only; the shim calls go there in this block
        prev = lib.gvi_test_set_grad_chunk_waves(1)
        try:
            loss3, g3 = gvi.calculate_loss_and_gradients(approx_params, x_dev, dev(pr["y"]))
        finally:
            lib.gvi_test_set_grad_chunk_waves(prev)
    assert float(loss1) == float(loss3)
    assert np.array_equal(host(g1["mean_output"]), host(g3["mean_output"]))
    a, b = host(g1["log_lengthscales"]), host(g3["log_lengthscales"])
    assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(a))
    assert abs(float(g1["log_scaling"]) - float(g3["log_scaling"])) <= 1e-12 * abs(float(g1["log_scaling"]))
    ref_loss, g_ls, g_ll, g_m = TT.gvi_loss_and_grads(pr["x"], pr["y"], pr["z"], pr["m_q"], pr["ls_q"], pr["ll_q"],
                                                       m_p, v_p, "kl", 0.5, 1e-5, False)
    assert np.max(np.abs(b - g_ll)) <= GRAD_TOL * max(np.max(np.abs(g_ll)), 1e-3)
    assert abs(float(g3["log_scaling"]) - g_ls) <= GRAD_TOL * max(abs(g_ls), 1e-3)


def test_gradient_scalar_lengthscale_and_prior_mode(S):
    from oracle import torch_twin as TT

    # 12 inducing points with a short lengthscale: cond(A) ~ 10
    pr = make_problem(400, 12, 1, seed=3, ll_base=4.5)
    kernel = S.ARDKernel(number_of_dimensions=1)
    approx_mean = S.CustomMean(mean_function=lambda params, x: params[: x.shape[0]])
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=pr["z"])
    approx_gp = S.ApproximateGPRegression(mean=approx_mean, kernel=sparse)
    ll_scalar = float(pr["ll_q"][0])
    ap = approx_gp.Parameters.construct(
        mean=approx_mean.Parameters.construct(custom=dev(pr["m_q"])),
        kernel=sparse.Parameters.construct(
            base_kernel=kernel.Parameters.construct(log_scaling=pr["ls_q"], log_lengthscales=ll_scalar)))
    const_mean = S.ConstantMean()
    exact_gp = S.GPRegression(mean=const_mean, kernel=kernel, x=pr["z"], y=pr["yz"])
    ep = exact_gp.Parameters.construct(log_observation_noise=np.log(0.5),
                                       mean=const_mean.Parameters.construct(constant=-0.1),
                                       kernel=kernel.Parameters.construct(log_scaling=pr["ls_p"],
                                                                          log_lengthscales=pr["ll_p"]))
    reg = S.projected.ProjectedKLRegularisation(gp=approx_gp, regulariser=exact_gp, regulariser_parameters=ep,
                                                mode="prior")
    gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=S.NegativeLogLikelihood(gp=approx_gp))
    loss, grads = gvi.calculate_loss_and_gradients(ap, pr["x"], pr["y"])
    m_p, v_p = O.exact_gp_prior(lambda a: O.ard_gram_diag(pr["ls_p"], pr["ll_p"], a), np.full(400, -0.1), np.log(0.5),
                                pr["x"])
    ref_loss, g_ls, g_ll, g_m = TT.gvi_loss_and_grads(pr["x"], pr["y"], pr["z"], pr["m_q"], pr["ls_q"],
                                                       np.array([ll_scalar]), m_p, v_p, "kl")
    assert abs(float(loss) - ref_loss) <= TOL * abs(ref_loss)
    assert grads["log_lengthscales"].shape == ()
    assert abs(float(grads["log_lengthscales"]) - g_ll[0]) <= GRAD_TOL * max(abs(g_ll[0]), 1e-3)
    assert abs(float(grads["log_scaling"]) - g_ls) <= GRAD_TOL * max(abs(g_ls), 1e-3)
    assert np.max(np.abs(host(grads["mean_output"]) - g_m)) <= GRAD_TOL * np.max(np.abs(g_m))


def test_branch_free_exp_is_within_one_ulp(S):
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.uniform(0, 700, 200000), -rng.uniform(0, 2, 200000), -10.0 ** rng.uniform(-300, 2, 50000),
                        [0.0, -0.0, -1e-320, -708.0, -708.5, -745.0, -1e6, -np.inf, -0.34657359027997264]])
    xd, out = dev(x), torch.empty(x.size, dtype=torch.float64, device="cuda")
    S._lib.check(S._lib.load_test().gvi_test_exp_f64(xd.data_ptr(), x.size, out.data_ptr(), S._lib.current_stream()), "exp")
    got, ref = host(out), np.exp(x)
    keep = x >= -708.0
    ulps = np.abs(got[keep] - ref[keep]) / np.spacing(ref[keep])
    assert ulps.max() <= 1.0, ulps.max()
    assert np.all(got[~keep] == 0.0) and got[np.flatnonzero(x == 0.0)[0]] == 1.0


def test_nan_arguments_come_out_as_nan(S):
    """A NaN in the inputs or hyper-parameters must surface as NaN (the reference's trainer stops on a NaN loss,
    experiments/shared/trainer.py:75-82): the exp clamp is a compare-select, not fmax."""
    x = np.array([np.nan, -1.0, -np.nan])
    xd, out = dev(x), torch.empty(x.size, dtype=torch.float64, device="cuda")
    S._lib.check(S._lib.load_test().gvi_test_exp_f64(xd.data_ptr(), x.size, out.data_ptr(), S._lib.current_stream()), "exp")
    got = host(out)
    assert np.isnan(got[0]) and np.isnan(got[2]) and got[1] == np.exp(-1.0)


def test_autograd_backward_and_training_loop(S):
    """README-style training (README.md:336-354 of the reference uses jax.grad + optax): here loss.backward() runs the
    CUDA gradient step, torch autograd carries dR/dm_q through a tanh-FCN mean, and Adam decreases the objective."""
    from oracle import torch_twin as TT

    g = np.load(os.path.join(GOLDEN, "config1_readme.npz"))
    x, y = g["x"], g["y"]
    xd, yd = dev(x), dev(y)
    kernel = S.ARDKernel(number_of_dimensions=1)
    z, yz = x[g["idx"]], y[g["idx"]]
    torch.manual_seed(0)
    w1 = (0.5 * torch.randn(1, 10, dtype=torch.float64, device="cuda")).requires_grad_(True)
    b1 = torch.zeros(10, dtype=torch.float64, device="cuda", requires_grad=True)
    w2 = (0.3 * torch.randn(10, 1, dtype=torch.float64, device="cuda")).requires_grad_(True)
    ls = torch.tensor(float(g["ls_q"]), dtype=torch.float64, device="cuda", requires_grad=True)
    ll = torch.tensor(g["ll_q"], dtype=torch.float64, device="cuda", requires_grad=True)
    fcn = lambda p, xx: torch.tanh(xx @ p["w1"] + p["b1"]) @ p["w2"]
    amean = S.CustomMean(mean_function=fcn)
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=z)
    approx = S.ApproximateGPRegression(mean=amean, kernel=sparse)
    ap = approx.Parameters.construct(
        mean=amean.Parameters.construct(custom={"w1": w1, "b1": b1, "w2": w2}),
        kernel=sparse.Parameters.construct(base_kernel=kernel.Parameters.construct(log_scaling=ls,
                                                                                   log_lengthscales=ll)))
    mean0 = S.ConstantMean()
    exact = S.GPRegression(mean=mean0, kernel=kernel, x=z, y=yz)
    ep = exact.Parameters.construct(log_observation_noise=0.0, mean=mean0.Parameters.construct(constant=0.0),
                                    kernel=kernel.Parameters.construct(log_scaling=float(g["log_scaling"]),
                                                                       log_lengthscales=g["log_ls"]))
    gvi = S.GeneralisedVariationalInference(
        regularisation=S.projected.ProjectedRenyiRegularisation(gp=approx, regulariser=exact,
                                                                regulariser_parameters=ep, mode="posterior"),
        empirical_risk=S.NegativeLogLikelihood(gp=approx))
    loss = gvi.calculate_loss(ap, xd, yd)
    loss.backward()
    # reference gradients: the torch twin differentiates the same objective end to end (mean included)
    t = lambda a: torch.tensor(host(a), dtype=torch.float64, requires_grad=True)
    cw1, cb1, cw2, cls_, cll = t(w1), t(b1), t(w2), t(ls), t(ll)
    xt, yt, zt = torch.tensor(x), torch.tensor(y), torch.tensor(z)
    mq = (torch.tanh(xt @ cw1 + cb1) @ cw2).reshape(-1)
    vq = TT.sparse_posterior_diag(cls_, cll, zt, xt)
    ref = TT.gaussian_nll(yt, mq, vq) + TT.regulariser("renyi", torch.tensor(g["m_p"]), torch.tensor(g["v_p"]), mq, vq)
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
    for ours, theirs in ((w1, cw1), (b1, cb1), (w2, cw2), (ls, cls_), (ll, cll)):
        scale = max(float(theirs.grad.abs().max()), 1e-6)
        assert float((ours.grad.cpu() - theirs.grad).abs().max()) <= GRAD_TOL * scale
    # a short Adam run lowers the objective (the reference trains with optax.adabelief the same way)
    opt = torch.optim.Adam([w1, b1, w2, ls, ll], lr=1e-2)
    first = float(loss.detach())
    for _ in range(40):
        opt.zero_grad()
        loss = gvi.calculate_loss(ap, xd, yd)
        loss.backward()
        opt.step()
    assert float(loss.detach()) < first and np.isfinite(float(loss.detach()))
    # no-grad evaluation takes the forward-only launch and agrees
    with torch.no_grad():
        assert abs(float(gvi.calculate_loss(ap, xd, yd)) - float(gvi.calculate_loss(ap, xd, yd))) == 0.0


# ---------------------------------------------------------------------------------------------------------------
# section 8(f)-1: fixed sparse posterior, SVGP family, tempered kernel
# ---------------------------------------------------------------------------------------------------------------
def test_fixed_sparse_svgp_and_tempered_kernels(S):
    from gvi_gaussian_process_b200.src.kernels import TemperedKernel
    from gvi_gaussian_process_b200.src.kernels.approximate import (CholeskySVGPKernel, DiagonalSVGPKernel,
                                                                   FixedSparsePosteriorKernel)

    rng = np.random.default_rng(11)
    n, m, d = 3000, 200, 4
    x, z, xt = rng.standard_normal((n, d)), rng.standard_normal((m, d)), rng.standard_normal((500, d))
    ls_r, ll_r = 0.1, rng.normal(0.3, 0.1, d)     # regulariser (frozen) kernel
    ls_b, ll_b = -0.1, rng.normal(0.5, 0.1, d)    # trainable base kernel
    g_r = lambda a, b: O.ard_gram(ls_r, ll_r, a, b)
    g_b = lambda a, b: O.ard_gram(ls_b, ll_b, a, b)
    k = S.ARDKernel(number_of_dimensions=d)
    p_r = k.Parameters.construct(log_scaling=ls_r, log_lengthscales=ll_r)
    p_b = k.Parameters.construct(log_scaling=ls_b, log_lengthscales=ll_b)
    # ---- fixed sparse posterior -------------------------------------------------------------------------------
    fk = FixedSparsePosteriorKernel(regulariser_kernel=k, regulariser_kernel_parameters=p_r, base_kernel=k,
                                    inducing_points=z)
    fp = fk.generate_parameters({"base_kernel": p_b})
    ref_full = O.fixed_sparse_posterior_gram(g_b, g_r, z, x[:60], x[:40])
    assert rel_norm(host(fk.calculate_gram(fp, x[:60], x[:40])), ref_full) <= TOL
    ref_diag = np.diag(O.fixed_sparse_posterior_gram(g_b, g_r, z, x[:400], x[:400]))
    assert rel_norm(host(fk.calculate_gram(fp, x, x, full_covariance=False))[:400], ref_diag) <= TOL
    # identity-kernel golden of the reference (9.9999e-06) through ARD with a huge inverse lengthscale at x = Z
    zz = np.array([[5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    k3 = S.ARDKernel(number_of_dimensions=3)
    delta = k3.Parameters.construct(log_scaling=0.0, log_lengthscales=np.full(3, 20.0))
    fk3 = FixedSparsePosteriorKernel(regulariser_kernel=k3, regulariser_kernel_parameters=delta, base_kernel=k3,
                                     inducing_points=zz)
    assert np.allclose(host(fk3.calculate_gram({"base_kernel": delta}, zz, zz, full_covariance=False)), 9.9999e-06,
                       rtol=1e-5)
    # ---- SVGP, Cholesky and diagonal parameterisations -----------------------------------------------------------
    ck = CholeskySVGPKernel(regulariser_kernel=k, regulariser_kernel_parameters=p_r, log_observation_noise=np.log(0.5),
                            inducing_points=z, training_points=xt)
    cp = ck.generate_parameters()
    lower_ref, logd_ref = O.svgp_cholesky_initial_parameters(g_r, z, xt, np.log(0.5))
    assert rel_norm(host(cp.el_matrix_lower_triangle), lower_ref) <= 1e-9
    assert np.max(np.abs(host(cp.el_matrix_log_diagonal) - logd_ref)) <= 1e-9
    lower = lower_ref + 0.01 * np.tril(rng.standard_normal((m, m)), k=-1)
    logd = logd_ref + 0.1 * rng.standard_normal(m)
    cp = ck.generate_parameters({"el_matrix_lower_triangle": lower, "el_matrix_log_diagonal": logd})
    sigma = O.cholesky_sigma(lower, logd)
    ref_diag = np.diag(O.svgp_gram(g_r, z, sigma, x[:400], x[:400]))
    got = host(ck.calculate_gram(cp, x, x, full_covariance=False))
    assert got.shape == (n,) and rel_norm(got[:400], ref_diag) <= TOL
    assert rel_norm(host(ck.calculate_gram(cp, x[:50], x[:30])), O.svgp_gram(g_r, z, sigma, x[:50], x[:30])) <= TOL
    dk = DiagonalSVGPKernel(regulariser_kernel=k, regulariser_kernel_parameters=p_r, log_observation_noise=np.log(0.5),
                            inducing_points=z, training_points=xt)
    dp = dk.generate_parameters()
    assert np.max(np.abs(host(dp.log_el_matrix_diagonal) - O.svgp_diagonal_initial_parameters(g_r, z, xt, np.log(0.5)))
                  ) <= 1e-9
    dlog = host(dp.log_el_matrix_diagonal) + 0.1 * rng.standard_normal(m)
    ref_diag = np.diag(O.svgp_gram(g_r, z, np.diag(np.exp(dlog)), x[:400], x[:400]))
    got = host(dk.calculate_gram({"log_el_matrix_diagonal": dlog}, x, x, full_covariance=False))
    assert rel_norm(got[:400], ref_diag) <= TOL
    # reference goldens (identity mock kernel): 0.5000099999 and 0.50001 on the diagonal, via ARD-delta at x = Z
    xtr = np.array([[1.0, 2.0, 3.0], [5.0, 1.0, 9.0], [1.5, 2.5, 3.5]])
    ck3 = CholeskySVGPKernel(regulariser_kernel=k3, regulariser_kernel_parameters=delta, log_observation_noise=0.0,
                             inducing_points=zz, training_points=xtr)
    lo3, ld3 = ck3.initialise_el_matrix_parameters()
    assert np.allclose(host(ld3), -0.34657359027997275, rtol=1e-12) and np.all(host(lo3) == 0.0)
    g3 = host(ck3.calculate_gram({"el_matrix_lower_triangle": np.zeros((2, 2)),
                                  "el_matrix_log_diagonal": np.array([-0.34657359, -0.34657359])}, zz, zz,
                                 full_covariance=False))
    assert np.allclose(g3, 0.5000099999000007, rtol=1e-7)
    # an approximate GP on top of an SVGP kernel goes through predict_probability unchanged
    gp = S.ApproximateGPRegression(mean=S.ConstantMean(), kernel=ck)
    pred = gp.predict_probability({"mean": {"constant": 0.3}, "kernel": cp.dict()}, x)
    assert np.allclose(host(pred.mean), 0.3) and rel_norm(host(pred.covariance)[:400], np.diag(
        O.svgp_gram(g_r, z, sigma, x[:400], x[:400]))) <= TOL
    # kernelised SVGP: frozen sparse-posterior part + trainable base kernel (kernelised_svgp_kernel.py:87-99)
    from gvi_gaussian_process_b200.src.kernels.approximate import KernelisedSVGPKernel
    kk = KernelisedSVGPKernel(base_kernel=k, regulariser_kernel=k, regulariser_kernel_parameters=p_r,
                              log_observation_noise=np.log(0.5), inducing_points=z, training_points=xt)
    kparams = kk.generate_parameters({"base_kernel": p_b})
    ref_k = O.svgp_gram(g_r, z, np.zeros((m, m)), x[:300], x[:200]) + g_b(x[:300], x[:200])
    assert rel_norm(host(kk.calculate_gram(kparams, x[:300], x[:200])), ref_k) <= TOL
    ref_kd = np.diag(O.svgp_gram(g_r, z, np.zeros((m, m)), x[:300], x[:300])) + O.ard_gram_diag(ls_b, ll_b, x[:300])
    assert rel_norm(host(kk.calculate_gram(kparams, x, x, full_covariance=False))[:300], ref_kd) <= TOL
    # the fused single-launch objective accepts these kernels too (q = SVGP GP, p = exact GP posterior)
    y = np.sin(x[:, 0]) + 0.1 * rng.standard_normal(n)
    yz = np.sin(z[:, 0])
    exact = S.GPRegression(mean=S.ConstantMean(), kernel=k, x=z, y=yz)
    ep = {"log_observation_noise": 0.0, "mean": {"constant": 0.0}, "kernel": p_r}
    qparams = {"mean": {"constant": 0.3}, "kernel": cp.dict()}
    gvi = S.GeneralisedVariationalInference(
        regularisation=S.projected.ProjectedKLRegularisation(gp=gp, regulariser=exact, regulariser_parameters=ep,
                                                             mode="posterior"),
        empirical_risk=S.NegativeLogLikelihood(gp=gp))
    assert gvi._fusable()
    v_q_full = np.diag(O.svgp_gram(g_r, z, sigma, x[:600], x[:600]))
    m_p, v_p = O.exact_gp_posterior(g_r, lambda a: O.ard_gram_diag(ls_r, ll_r, a), np.zeros(600), z, yz, 0.0, x[:600])
    ref = O.gaussian_nll(y[:600], np.full(600, 0.3), v_q_full) + O.projected_regularisation(
        "kl", m_p, v_p, np.full(600, 0.3), v_q_full)
    assert abs(float(gvi.calculate_loss(qparams, x[:600], y[:600])) - ref) <= TOL * abs(ref)
    # ---- tempered kernel (reference tests/gps/test_regression.py:252-372: factor 2 doubles the variance) -----------
    sp = S.SparsePosteriorKernel(base_kernel=k, inducing_points=z)
    sp_params = sp.generate_parameters({"base_kernel": p_b})
    tk = TemperedKernel(base_kernel=sp, base_kernel_parameters=sp_params, number_output_dimensions=1)
    base_var = host(sp.calculate_gram(sp_params, x, x, full_covariance=False))
    temp_var = host(tk.calculate_gram({"log_tempering_factor": np.log(2.0)}, x, x, full_covariance=False))
    assert np.allclose(temp_var, 2.0 * base_var, rtol=1e-15)


def test_gradients_fixed_sparse_and_svgp_kernels(S):
    """loss.backward() for the section 8(f)-1 kernels against torch-autograd twins of the reference formulas:
    FixedSparsePosteriorKernel (frozen Cholesky: no d K_ZZ term), Cholesky and diagonal SVGP (dR/dE = 2 tril(E C))."""
    from gvi_gaussian_process_b200.src.kernels.approximate import (CholeskySVGPKernel, DiagonalSVGPKernel,
                                                                   FixedSparsePosteriorKernel)
    from oracle import torch_twin as TT

    rng = np.random.default_rng(21)
    n, m, d = 900, 65, 3
    x, z, xt = rng.standard_normal((n, d)), rng.standard_normal((m, d)), rng.standard_normal((200, d))
    y = np.sin(x[:, 0]) + 0.1 * rng.standard_normal(n)
    yz = np.sin(z[:, 0])
    ls_r, ll_r = 0.1, rng.normal(0.6, 0.1, d)
    k = S.ARDKernel(number_of_dimensions=d)
    p_r = k.Parameters.construct(log_scaling=ls_r, log_lengthscales=ll_r)
    exact = S.GPRegression(mean=S.ConstantMean(), kernel=k, x=z, y=yz)
    ep = {"log_observation_noise": 0.0, "mean": {"constant": 0.0}, "kernel": p_r}
    m_p, v_p = O.exact_gp_posterior(lambda a, b: O.ard_gram(ls_r, ll_r, a, b), lambda a: O.ard_gram_diag(ls_r, ll_r, a),
                                    np.zeros(n), z, yz, 0.0, x)
    tt = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    xt_, yt_, zt_, mp_, vp_ = tt(x), tt(y), tt(z), tt(m_p), tt(v_p)
    w = (0.3 * torch.randn(d, 1, dtype=torch.float64, device="cuda")).requires_grad_(True)
    amean = S.CustomMean(mean_function=lambda p, xx: torch.tanh(xx @ p))

    def check(kernel, kernel_params, leaves, twin_var):
        gp = S.ApproximateGPRegression(mean=amean, kernel=kernel)
        params = gp.Parameters.construct(mean=amean.Parameters.construct(custom=w), kernel=kernel_params)
        gvi = S.GeneralisedVariationalInference(
            regularisation=S.projected.ProjectedRenyiRegularisation(gp=gp, regulariser=exact, regulariser_parameters=ep,
                                                                    mode="posterior", alpha=0.5),
            empirical_risk=S.NegativeLogLikelihood(gp=gp))
        for t in leaves + [w]:
            t.grad = None
        loss = gvi.calculate_loss(params, dev(x), dev(y))
        loss.backward()
        c_leaves = [tt(host(t)).requires_grad_(True) for t in leaves]
        cw = tt(host(w)).requires_grad_(True)
        mq = torch.tanh(xt_ @ cw).reshape(-1)
        vq = twin_var(*c_leaves)
        ref = TT.gaussian_nll(yt_, mq, vq) + TT.regulariser("renyi", mp_, vp_, mq, vq, 0.5)
        ref.backward()
        assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
        for ours, theirs in zip(leaves + [w], c_leaves + [cw]):
            scale = max(float(theirs.grad.abs().max()), 1e-6)
            err = float((ours.grad.cpu() - theirs.grad).abs().max())
            assert err <= GRAD_TOL * scale, (type(kernel).__name__, err, scale)

    # fixed sparse posterior: trainable base kernel, Cholesky from the regulariser kernel
    ls_b = torch.tensor(-0.1, dtype=torch.float64, device="cuda", requires_grad=True)
    ll_b = torch.tensor(rng.normal(0.8, 0.1, d), dtype=torch.float64, device="cuda", requires_grad=True)
    fk = FixedSparsePosteriorKernel(regulariser_kernel=k, regulariser_kernel_parameters=p_r, base_kernel=k,
                                    inducing_points=z)
    check(fk, fk.Parameters.construct(base_kernel=k.Parameters.construct(log_scaling=ls_b, log_lengthscales=ll_b)),
          [ls_b, ll_b], lambda a, b: TT.fixed_sparse_posterior_diag(a, b, tt(ls_r), tt(ll_r), zt_, xt_))
    # Cholesky SVGP
    ck = CholeskySVGPKernel(regulariser_kernel=k, regulariser_kernel_parameters=p_r, log_observation_noise=np.log(0.5),
                            inducing_points=z, training_points=xt)
    lo0, ld0 = ck.initialise_el_matrix_parameters()
    lo = (lo0 + 0.01 * torch.tril(torch.randn(m, m, dtype=torch.float64, device="cuda"), -1)).requires_grad_(True)
    ld = (ld0 + 0.05 * torch.randn(m, dtype=torch.float64, device="cuda")).requires_grad_(True)

    def chol_var(plo, pld):
        el = torch.tril(plo, -1) + torch.diag(torch.exp(pld))
        return TT.svgp_diag(tt(ls_r), tt(ll_r), zt_, xt_, el.T @ el)

    check(ck, ck.Parameters.construct(el_matrix_lower_triangle=lo, el_matrix_log_diagonal=ld), [lo, ld], chol_var)
    # diagonal SVGP
    dk = DiagonalSVGPKernel(regulariser_kernel=k, regulariser_kernel_parameters=p_r, log_observation_noise=np.log(0.5),
                            inducing_points=z, training_points=xt)
    dl = (dk.initialise_diagonal_parameters() + 0.05 * torch.randn(m, dtype=torch.float64, device="cuda")
          ).requires_grad_(True)
    check(dk, dk.Parameters.construct(log_el_matrix_diagonal=dl), [dl],
          lambda pdl: TT.svgp_diag(tt(ls_r), tt(ll_r), zt_, xt_, torch.diag(torch.exp(pdl))))
