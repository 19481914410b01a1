"""GPU parity tests of "the other kernels.calculate_gram" (BASELINE.json north_star): InnerProductKernel, PolynomialKernel,
CustomKernel and CustomMappingKernel through the src-compatible shim, and the predictive / regulariser path over them
(K_ZZ and K(x, Z) from the kernel's own Gram routine, factorisation and triangular products in the CUDA kernels).
Values against the NumPy oracle (<= 1e-10 relative), the reference's own kernel goldens through the product path."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def S():
    import types

    from gvi_gaussian_process_b200 import _lib
    from gvi_gaussian_process_b200.src import _lib_proxy
    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
    from gvi_gaussian_process_b200.src.gps import ApproximateGPRegression, GPRegression
    from gvi_gaussian_process_b200.src.kernels import CustomKernel, CustomMappingKernel
    from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.non_stationary import InnerProductKernel, PolynomialKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean
    from gvi_gaussian_process_b200.src.regularisations import projected

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.load()
    return types.SimpleNamespace(**locals())


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


def rel_elem(a, ref):
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), 1e-300)))


def rel_norm(a, ref):
    return float(np.max(np.abs(a - ref)) / max(np.max(np.abs(ref)), 1e-300))


# ---- the reference's own goldens through the product path (tests/kernels/test_kernels.py:16-94) --------------------
@pytest.mark.parametrize("log_scaling,x1,x2,k", [(np.log(2.4), [1.0, 2.0, 3.0], [1.5, 2.5, 3.5], 40.8),
                                                 (np.log(2.2), 6.0, 6.0, 79.2), (np.log(2.0), 6.0, 6.0, 72.0)])
def test_inner_product_kernel_goldens(S, log_scaling, x1, x2, k):
    kernel = S.InnerProductKernel()
    parameters = kernel.Parameters.construct(log_scaling=log_scaling)
    out = host(kernel.calculate_gram(parameters, x1=np.array(x1), x2=np.array(x2)))
    assert out.shape == (1, 1) and np.isclose(out[0, 0], k, rtol=1e-12)


@pytest.mark.parametrize("degree,x1,x2,k,rtol", [(3, [1.0, 2.0, 3.0], [1.5, 2.5, 3.5], 73403079.4946829, 1e-12),
                                                 (1, 6.0, 6.0, 884.81980837, 1e-9),
                                                 (1.5, 6.0, 6.0, 26319.78000332, 1e-9)])
def test_polynomial_kernel_goldens(S, degree, x1, x2, k, rtol):
    kernel = S.PolynomialKernel(polynomial_degree=degree)
    parameters = kernel.generate_parameters({"log_constant": 0.5, "log_scaling": 3.2})
    out = host(kernel.calculate_gram(parameters, x1=np.array(x1), x2=np.array(x2)))
    assert out.shape == (1, 1) and np.isclose(out[0, 0], k, rtol=rtol)


def test_inner_product_kernel_rejects_other_parameters(S):
    # tests/kernels/test_kernels.py:214-237: parameters of another kernel raise an AssertionError
    ard = S.ARDKernel(number_of_dimensions=3)
    wrong = ard.Parameters.construct(log_scaling=0.5, log_lengthscales=np.zeros(3))
    with pytest.raises(AssertionError):
        S.InnerProductKernel().calculate_gram(wrong, x1=np.ones((1, 3)), x2=np.ones((2, 3)))


def test_custom_kernel_goldens(S):
    # tests/kernels/test_kernels.py:240-290: an all-ones kernel function, with and without a preprocess function
    ones = lambda parameters, x, y: torch.ones((x.shape[0], y.shape[0]), dtype=torch.float64, device=x.device)
    k1 = S.CustomKernel(kernel_function=ones)
    out = k1.calculate_gram(k1.Parameters.construct(custom=None), x1=np.array([[1.0, 2.0, 3.0]]),
                            x2=np.array([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5]]))
    assert np.array_equal(host(out), np.ones((1, 2)))
    k2 = S.CustomKernel(kernel_function=ones, preprocess_function=lambda x: x.reshape(x.shape[0], -1))
    out = k2.calculate_gram(k2.Parameters.construct(custom=None), x1=np.array([[[1.0], [2.0], [3.0]]]),
                            x2=np.array([[[1.0], [2.0], [3.0]], [[1.5], [2.5], [3.5]]]))
    assert np.array_equal(host(out), np.ones((1, 2)))


# ---- Grams against the oracle, ragged shapes -------------------------------------------------------------------------
@pytest.mark.parametrize("m1,m2,d,degree", [(1, 1, 1, 1), (33, 257, 3, 2), (300, 70, 8, 3), (65, 513, 40, 1.5),
                                            (1000, 1024, 8, 4), (17, 5, 784, 2)])
def test_polynomial_and_inner_product_gram_parity(S, m1, m2, d, degree):
    rng = np.random.default_rng(m1 * 7 + m2)
    x1, x2 = rng.standard_normal((m1, d)) / np.sqrt(d), rng.standard_normal((m2, d)) / np.sqrt(d)
    log_constant, log_scaling = 0.3, -0.2  # base = 0.82 <x,y> + 1.35 stays positive for the fractional degree
    if degree == 1.5:
        x1, x2 = np.abs(x1), np.abs(x2)
    poly = S.PolynomialKernel(polynomial_degree=degree)
    pp = poly.Parameters.construct(log_constant=log_constant, log_scaling=log_scaling)
    ref = O.polynomial_gram(log_constant, log_scaling, degree, x1, x2)
    # normwise: an inner product cancels, so single entries near a zero crossing carry eps * sum |x_d z_d|
    assert rel_norm(host(poly.calculate_gram(pp, x1, x2)), ref) <= TOL
    inner = S.InnerProductKernel()
    ip = inner.Parameters.construct(log_scaling=log_scaling)
    ref_ip = O.inner_product_gram(log_scaling, x1, x2)
    assert rel_norm(host(inner.calculate_gram(ip, x1, x2)), ref_ip) <= TOL
    # x2 = None: the Gram of x1 with itself; diagonal mode on equal inputs and on two different inputs (pairs)
    assert rel_norm(host(poly.calculate_gram(pp, x1)), O.polynomial_gram(log_constant, log_scaling, degree, x1)) <= TOL
    diag = host(poly.calculate_gram(pp, x1, full_covariance=False))
    assert diag.shape == (m1,)
    assert rel_elem(diag, np.diag(O.polynomial_gram(log_constant, log_scaling, degree, x1))) <= TOL
    n = min(m1, m2)
    pairs = host(poly.calculate_gram(pp, x1[:n], x2[:n], full_covariance=False))
    assert np.max(np.abs(pairs - np.diag(ref[:n, :n]))) <= TOL * np.max(np.abs(ref))


def test_ard_diagonal_mode_on_two_different_inputs(S):
    # src/kernels/base.py:132-147: full_covariance=False evaluates k(x1_i, x2_i)
    rng = np.random.default_rng(5)
    x1, x2 = rng.standard_normal((700, 5)), rng.standard_normal((700, 5))
    ll = 0.3 * rng.standard_normal(5)
    ard = S.ARDKernel(number_of_dimensions=5)
    p = ard.Parameters.construct(log_scaling=0.4, log_lengthscales=ll)
    out = host(ard.calculate_gram(p, x1, x2, full_covariance=False))
    assert rel_elem(out, np.diag(O.ard_gram(0.4, ll, x1, x2))) <= TOL
    with pytest.raises(AssertionError):
        ard.calculate_gram(p, x1, x2[:10], full_covariance=False)


def test_custom_mapping_kernel_matches_oracle(S):
    # feature map (a small tanh network of the host framework) followed by a polynomial kernel on the features
    rng = np.random.default_rng(9)
    x1, x2 = rng.standard_normal((120, 6)), rng.standard_normal((47, 6))
    w, b = rng.standard_normal((6, 11)) / 3.0, 0.1 * rng.standard_normal(11)
    phi = lambda params, x: torch.tanh(x @ params["w"] + params["b"])
    phi_np = lambda x: np.tanh(x @ w + b)
    base = S.PolynomialKernel(polynomial_degree=2)
    kernel = S.CustomMappingKernel(base_kernel=base, feature_mapping=phi)
    parameters = kernel.generate_parameters({"base_kernel": {"log_constant": -0.5, "log_scaling": 0.2},
                                             "feature_mapping": {"w": dev(w), "b": dev(b)}})
    ref = O.polynomial_gram(-0.5, 0.2, 2, phi_np(x1), phi_np(x2))
    assert rel_norm(host(kernel.calculate_gram(parameters, x1, x2)), ref) <= TOL
    diag = host(kernel.calculate_gram(parameters, x1, full_covariance=False))
    assert rel_elem(diag, np.diag(O.polynomial_gram(-0.5, 0.2, 2, phi_np(x1)))) <= TOL


# ---- predictives over a non-ARD base kernel: Grams from the kernel, the rest in the CUDA factor / tile kernels -------
def _poly_problem(n, m, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)) / np.sqrt(d)
    z = rng.standard_normal((m, d)) / np.sqrt(d)
    w = rng.standard_normal(d)
    y = np.sin(x @ w) + 0.1 * rng.standard_normal(n)
    yz = np.sin(z @ w) + 0.1 * rng.standard_normal(m)
    return x, z, y, yz


@pytest.mark.parametrize("n,m,d,degree", [(50, 7, 2, 2), (3000, 40, 3, 3), (5000, 300, 8, 2), (700, 1200, 8, 2)])
def test_sparse_posterior_and_exact_gp_over_polynomial_kernel(S, n, m, d, degree):
    x, z, y, yz = _poly_problem(n, m, d, seed=n + m)
    lc, ls, lam, log_noise = 0.2, -0.1, 0.1, -1.0
    gram = lambda a, b: O.polynomial_gram(lc, ls, degree, a, b)
    diag = lambda a: np.power(np.exp(ls) * np.einsum("ij,ij->i", a, a) + np.exp(lc), degree)
    kernel = S.PolynomialKernel(polynomial_degree=degree)
    kp = kernel.Parameters.construct(log_constant=lc, log_scaling=ls)
    # approximate GP: sparse posterior kernel in diagonal mode, absolute jitter (a polynomial Gram has rank << M)
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=z, diagonal_regularisation=lam,
                                     is_diagonal_regularisation_absolute_scale=True)
    sp = sparse.Parameters.construct(base_kernel=kp)
    v_q = host(sparse.calculate_gram(sp, x, full_covariance=False))
    v_q_ref = O.sparse_posterior_diag(gram, diag, z, x, lam, True)
    assert v_q.shape == (n,)
    # a polynomial Gram has low rank, so v = k(x,x) - k^T A^-1 k cancels down to the jitter level: the tolerance is taken
    # relative to the quantities that are subtracted, max k(x,x) (numerics)
    kxx = float(np.max(diag(x)))
    assert np.max(np.abs(v_q - v_q_ref)) <= TOL * kxx
    # exact GP posterior over the same kernel (regulariser GP): noise keeps A well conditioned
    const = S.ConstantMean()
    exact = S.GPRegression(mean=const, kernel=kernel, x=z, y=yz)
    ep = exact.Parameters.construct(log_observation_noise=log_noise, mean=const.Parameters.construct(constant=0.25),
                                    kernel=kp)
    p = exact.predict_probability(ep, x)
    m_ref, v_ref = O.exact_gp_posterior(gram, diag, np.full(n, 0.25), z, yz, log_noise, x)
    assert host(p.mean).shape == (1, n)
    assert rel_norm(host(p.mean).reshape(-1), m_ref) <= TOL
    assert np.max(np.abs(host(p.covariance) - v_ref)) <= TOL * kxx


def test_gvi_objective_over_polynomial_kernels(S):
    """NLL + projected KL with q = approximate GP over a sparse posterior polynomial kernel and p = exact GP over a
    polynomial kernel: the general (term-by-term) objective path over caller-supplied Grams."""
    n, m, d = 4000, 64, 4
    x, z, y, yz = _poly_problem(n, m, d, seed=21)
    lc, ls, lam, log_noise = 0.1, 0.0, 1e-2, -0.5
    gram = lambda a, b: O.polynomial_gram(lc, ls, 2, a, b)
    diag = lambda a: np.power(np.exp(ls) * np.einsum("ij,ij->i", a, a) + np.exp(lc), 2)
    rng = np.random.default_rng(3)
    m_q = 0.3 * np.tanh(x @ rng.standard_normal(d))
    kernel = S.PolynomialKernel(polynomial_degree=2)
    kp = kernel.Parameters.construct(log_constant=lc, log_scaling=ls)
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=z, diagonal_regularisation=lam,
                                     is_diagonal_regularisation_absolute_scale=True)
    amean = S.CustomMean(mean_function=lambda params, xx: params[: xx.shape[0]])
    approx = S.ApproximateGPRegression(mean=amean, kernel=sparse)
    ap = approx.Parameters.construct(mean=amean.Parameters.construct(custom=dev(m_q)),
                                     kernel=sparse.Parameters.construct(base_kernel=kp))
    const = S.ConstantMean()
    exact = S.GPRegression(mean=const, kernel=kernel, x=z, y=yz)
    ep = exact.Parameters.construct(log_observation_noise=log_noise, mean=const.Parameters.construct(constant=0.0),
                                    kernel=kp)
    reg = S.projected.ProjectedKLRegularisation(gp=approx, regulariser=exact, regulariser_parameters=ep,
                                                mode="posterior")
    gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=S.NegativeLogLikelihood(gp=approx))
    loss = float(gvi.calculate_loss(ap, x, y))
    v_q = O.sparse_posterior_diag(gram, diag, z, x, lam, True)
    m_p, v_p = O.exact_gp_posterior(gram, diag, np.zeros(n), z, yz, log_noise, x)
    ref = O.gaussian_nll(y, m_q, v_q) + O.projected_regularisation("kl", m_p, v_p, m_q, v_q)
    assert abs(loss - ref) <= 1e-8 * abs(ref)


def test_generic_gram_entry_points_equal_the_ard_path(S):
    """gvi_gp_factor_gram_f64 + gvi_gp_predict_gram_f64 fed with ARD Grams must reproduce the fused ARD path (same
    factorisation and tile kernel, the K tile read from memory instead of evaluated): M below and above one panel."""
    L = S._lib_proxy
    for n, m, d in [(2000, 96, 5), (900, 1300, 8)]:
        rng = np.random.default_rng(n)
        x, z = rng.standard_normal((n, d)), rng.standard_normal((m, d))
        ard = S.ARDKernel(number_of_dimensions=d)
        kp = ard.Parameters.construct(log_scaling=0.2, log_lengthscales=np.full(d, np.log(1.0 / d)))
        sparse = S.SparsePosteriorKernel(base_kernel=ard, inducing_points=z)
        v_fused = host(sparse.calculate_gram(sparse.Parameters.construct(base_kernel=kp), x, full_covariance=False))
        f = L.GpFactor(m, 1).factor_from_gram(ard.calculate_gram(kp, z, z).contiguous(), 1e-5, False)
        _, v_gram = L.gp_predict_any_kernel(f, ard, kp, dev(z), dev(x), want_mean=False)
        # (900 x 1300 at D = 8 takes the guarded tensor-pipe Gram of csrc/gram_mma.cu: entries within ~10 ulp of the
        # direct differences, amplified by cond(A) in v = k(x,x) - k^T A^-1 k: 1.6e-12 measured, 3.5e-14 with the
        # direct-difference Gram, tools/lab/gram_guard_dev.py)
        assert rel_norm(host(v_gram), v_fused) <= (1e-12 if n * m < (1 << 18) else 2e-11)


def test_svgp_mean_golden(S):
    # tests/test_means.py:96-128: mock mean = ones, mock kernel = all-ones Gram (ARD with exp(log_ls) = 0), weights [1, 2]
    from gvi_gaussian_process_b200.src.means import SVGPMean

    class OnesMean(S.CustomMean):
        pass

    ones_mean = OnesMean(mean_function=lambda params, x: torch.ones(x.shape[0], dtype=torch.float64, device=x.device))
    ard = S.ARDKernel(number_of_dimensions=3)
    ones_kernel_parameters = ard.Parameters.construct(log_scaling=0.0, log_lengthscales=np.full(3, -np.inf))
    svgp_mean = SVGPMean(regulariser_mean_parameters=ones_mean.Parameters.construct(custom=None),
                         regulariser_mean=ones_mean, regulariser_kernel_parameters=ones_kernel_parameters,
                         regulariser_kernel=ard, inducing_points=np.ones((2, 3)))
    out = svgp_mean.predict(parameters=svgp_mean.generate_parameters({"weights": np.array([1.0, 2.0])}),
                            x=np.array([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5]]))
    assert np.array_equal(host(out), 4.0 * np.ones(2))


def test_regression_pipeline_callers(S, tmp_path):
    """The reference's regression pipeline end to end on a toy curve (experiments/regression/trainers.py:19-85,
    experiments/shared/trainers.py:15-114, experiments/regression/metrics.py:14-28): greedy selection + exact-GP NLL
    training, approximate GP trained on the GVI objective, tempering, metrics - every stage through the CUDA path."""
    from gvi_gaussian_process_b200.experiments.regression.metrics import calculate_metrics
    from gvi_gaussian_process_b200.experiments.regression.trainers import train_regulariser_gp
    from gvi_gaussian_process_b200.experiments.shared import (Data, TrainerSettings, inducing_points_selector_resolver,
                                                              train_approximate_gp, train_tempered_gp)
    from gvi_gaussian_process_b200.src.kernels import TemperedKernel

    rng = np.random.default_rng(0)
    n, m = 600, 24
    x = np.sort(rng.uniform(-3.0, 3.0, (n, 1)), axis=0)
    y = np.sin(2.0 * x[:, 0]) + 0.15 * rng.standard_normal(n)
    data = Data(x=dev(x), y=dev(y), name="train")
    kernel = S.ARDKernel(number_of_dimensions=1)
    kp = kernel.Parameters.construct(log_scaling=0.0, log_lengthscales=np.array([0.0]))
    settings = TrainerSettings(seed=0, optimiser_schema="adam", learning_rate=0.05, number_of_epochs=15, batch_size=m,
                               batch_shuffle=True, batch_drop_last=False)
    exact, exact_params, history = train_regulariser_gp(
        data=data, empirical_risk_schema="negative_log_likelihood", trainer_settings=settings, kernel=kernel,
        kernel_parameters=kp, inducing_points_selector=inducing_points_selector_resolver("conditional_variance"),
        number_of_inducing_points=m, empirical_risk_break_condition=-1e9, save_checkpoint_frequency=0,
        checkpoint_path=str(tmp_path / "regulariser"))
    risks = [float(h["empirical-risk"]) for h in history]
    assert len(risks) == 15 and risks[-1] < risks[0] and np.all(np.isfinite(risks))
    assert exact.x.shape == (m, 1)
    # approximate GP: sparse posterior kernel over the learned kernel, tanh-network mean of the host framework
    torch.manual_seed(0)
    w1 = (0.5 * torch.randn(1, 16, dtype=torch.float64, device="cuda"))
    w2 = (0.1 * torch.randn(16, dtype=torch.float64, device="cuda"))
    amean = S.CustomMean(mean_function=lambda p, xx: torch.tanh(xx @ p["w1"]) @ p["w2"])
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=exact.x)
    approx = S.ApproximateGPRegression(mean=amean, kernel=sparse)
    ap = approx.Parameters.construct(
        mean=amean.Parameters.construct(custom={"w1": w1, "w2": w2}),
        kernel=sparse.Parameters.construct(base_kernel=kernel.Parameters.construct(**exact_params.kernel.dict())))
    settings_q = TrainerSettings(seed=1, optimiser_schema="adabelief", learning_rate=0.02, number_of_epochs=12,
                                 batch_size=200, batch_shuffle=True, batch_drop_last=False)
    ap, history_q = train_approximate_gp(
        data=data, empirical_risk_schema="negative_log_likelihood",
        regularisation_config={"regularisation_schema": "projected_renyi",
                               "regularisation_kwargs": {"mode": "posterior", "alpha": 0.5}},
        trainer_settings=settings_q, approximate_gp=approx, approximate_gp_parameters=ap, regulariser=exact,
        regulariser_parameters=exact_params, save_checkpoint_frequency=0, checkpoint_path=str(tmp_path / "approximate"))
    objective = [float(h["gvi-objective"]) for h in history_q]
    assert len(objective) == 12 and objective[-1] < objective[0] and np.all(np.isfinite(objective))
    assert abs(float(history_q[-1]["empirical-risk"]) + float(history_q[-1]["regularisation"]) - objective[-1]) <= (
        1e-10 * abs(objective[-1]))
    # tempering: only the scalar factor of the tempered kernel is trained
    tempered_kernel = TemperedKernel(base_kernel=sparse, base_kernel_parameters=ap.kernel)
    tempered = S.ApproximateGPRegression(mean=amean, kernel=tempered_kernel)
    # start clearly miscalibrated (variance x e^2), so that every optimiser step towards the optimum lowers the NLL
    tp = tempered.generate_parameters({"mean": ap.mean, "kernel": {"log_tempering_factor": 2.0}})
    before = calculate_metrics({"train": data}, tempered, tp)["train"]["negative-log-likelihood"]
    settings_t = TrainerSettings(seed=2, optimiser_schema="adam", learning_rate=0.1, number_of_epochs=10,
                                 batch_size=300, batch_shuffle=True, batch_drop_last=False)
    tp, history_t = train_tempered_gp(data=data, empirical_risk_schema="negative_log_likelihood",
                                      trainer_settings=settings_t, tempered_gp=tempered, tempered_gp_parameters=tp,
                                      save_checkpoint_frequency=0, checkpoint_path=str(tmp_path / "tempered"))
    after = calculate_metrics({"train": data, "validation": None}, tempered, tp, y_std=1.0)
    assert set(after) == {"train"}
    assert after["train"]["negative-log-likelihood"] < before
    assert float(tp.kernel.log_tempering_factor) != 2.0  # (it grows: q drops the observation noise, so it is overconfident)
    # the metric is the NLL the objective's risk term reports (+ log y_std = 0)
    assert abs(after["train"]["negative-log-likelihood"] - float(history_t[-1]["empirical-risk"])) <= 1e-10
    # mean parameters untouched by the tempering stage
    assert torch.equal(tp.mean.custom["w1"], ap.mean.custom["w1"])


def test_selector_over_polynomial_and_custom_kernels_matches_oracle(S):
    """ConditionalVarianceInducingPointsSelector over non-ARD kernels (the reference's selector only calls
    kernel.calculate_gram, src/inducing_points_selection.py:121-166): pivots bit-exact against the oracle.  Degree 3 in
    6-D has 84 monomial features, so 24 pivots stay far above the rounding floor of the residual variances."""
    from gvi_gaussian_process_b200.src.inducing_points_selection import ConditionalVarianceInducingPointsSelector
    from gvi_gaussian_process_b200.src.utils import jax_prng
    from oracle import jax_prng as OP

    rng = np.random.default_rng(17)
    n, m, d = 4000, 24, 6
    x = rng.standard_normal((n, d)) / np.sqrt(d)
    lc, ls, degree = 0.1, 0.3, 3
    _, sub = OP.split(OP.prng_key(5))
    perm = OP.permutation(sub, n)
    gram = lambda a, b: O.polynomial_gram(lc, ls, degree, a, b)
    diag = lambda a: np.power(np.exp(ls) * np.einsum("ij,ij->i", a, a) + np.exp(lc), degree)
    piv = O.conditional_variance_select(x[perm], m, gram, diag, use_argsort=False)
    poly = S.PolynomialKernel(polynomial_degree=degree)
    pp = poly.Parameters.construct(log_constant=lc, log_scaling=ls)
    z, idx = ConditionalVarianceInducingPointsSelector().compute_inducing_points(jax_prng.PRNGKey(5), x, m, poly, pp)
    assert np.array_equal(host(idx), perm[piv])
    assert np.array_equal(host(z), x[perm[piv]])
    # the same kernel wrapped as a CustomKernel (host-framework function) selects the same points
    fn = lambda params, a, b: (np.exp(ls) * (a @ b.T) + np.exp(lc)) ** degree
    custom = S.CustomKernel(kernel_function=fn)
    _, idx_c = ConditionalVarianceInducingPointsSelector().compute_inducing_points(
        jax_prng.PRNGKey(5), x, m, custom, custom.Parameters.construct(custom=None))
    assert np.array_equal(host(idx_c), perm[piv])


def test_conformal_gp_regression_matches_oracle(S):
    """src/conformal_calibration/regression: Gaussian intervals of the exact GP's predictive (CUDA path), conformally
    calibrated on a held-out split; against the NumPy restatement (no reference test exists: parity unpinned)."""
    from gvi_gaussian_process_b200.src.conformal_calibration.regression import ConformalGPRegression

    rng = np.random.default_rng(4)
    n, m, d = 900, 40, 2
    x = rng.standard_normal((n, d))
    y = np.sin(x @ np.array([1.0, -0.5])) + 0.2 * rng.standard_normal(n)
    z, yz = x[:m], y[:m]
    x_cal, y_cal, x_test = x[m:500], y[m:500], x[500:]
    ls, ll, log_noise = 0.1, np.array([0.2, -0.1]), -2.0
    kernel = S.ARDKernel(number_of_dimensions=d)
    const = S.ConstantMean()
    gp = S.GPRegression(mean=const, kernel=kernel, x=z, y=yz)
    gp_parameters = gp.Parameters.construct(
        log_observation_noise=log_noise, mean=const.Parameters.construct(constant=0.1),
        kernel=kernel.Parameters.construct(log_scaling=ls, log_lengthscales=ll))
    conformal = ConformalGPRegression(x_calibration=x_cal, y_calibration=y_cal, gp=gp, gp_parameters=gp_parameters)
    gram = lambda a, b: O.ard_gram(ls, ll, a, b)
    diag = lambda a: O.ard_gram_diag(ls, ll, a)
    m_c, v_c = O.exact_gp_posterior(gram, diag, np.full(len(x_cal), 0.1), z, yz, log_noise, x_cal)
    m_t, v_t = O.exact_gp_posterior(gram, diag, np.full(len(x_test), 0.1), z, yz, log_noise, x_test)
    coverages = [0.5, 0.9, 0.999]
    lower, upper = conformal.predict_coverage(x=x_test, coverage=np.array(coverages))
    assert lower.shape == (3, len(x_test)) and upper.shape == (3, len(x_test))
    for i, c in enumerate(coverages):
        lo_ref, up_ref = O.conformal_gp_coverage(m_c, v_c, y_cal, m_t, v_t, c)
        assert rel_norm(host(lower[i]), lo_ref) <= 1e-9 and rel_norm(host(upper[i]), up_ref) <= 1e-9
    widths = host(conformal.calculate_average_interval_width(x=x_test, coverage=np.array(coverages)))
    assert widths.shape == (3,) and widths[0] < widths[1] < widths[2]
    # empirical coverage on the calibration split itself is at least the nominal one (split-conformal guarantee there)
    lo_c, up_c = conformal.predict_coverage(x=x_cal, coverage=0.9)
    inside = (host(lo_c[0]) <= y_cal) & (y_cal <= host(up_c[0]))
    assert inside.mean() >= 0.9


# ---- gradients for non-ARD kernels: dense VJP of the caller-supplied-Gram predictive (PredictGram) + DotGram ------------
def _feature_problem(n, m, d, f, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    z = rng.standard_normal((m, d))
    wt = rng.standard_normal(d)
    y = np.sin(x @ wt) + 0.1 * rng.standard_normal(n)
    yz = np.sin(z @ wt) + 0.1 * rng.standard_normal(m)
    w = rng.standard_normal((d, f)) / np.sqrt(d)
    b = 0.1 * rng.standard_normal(f)
    m_q = 0.3 * np.tanh(x @ rng.standard_normal(d))
    return x, z, y, yz, w, b, m_q


def test_gvi_gradients_through_feature_map_kernel_match_autograd_twin(S):
    """The reference's "1-layer-fcnn-tanh-mapping-linear" kernel (CustomMappingKernel: dense + tanh features, polynomial
    kernel of degree 1 on them) as the base kernel of q's sparse posterior; NLL + projected KL against a fixed
    regulariser.  Gradients w.r.t. the network weights, log_scaling, log_constant and the mean values against a torch
    CPU autograd twin of the same objective (<= 1e-8)."""
    from oracle import torch_twin as T

    n, m, d, f = 1500, 30, 3, 10
    x, z, y, yz, w, b, m_q = _feature_problem(n, m, d, f, seed=8)
    ls, lc, reg = 0.2, -0.4, 1e-3
    rng = np.random.default_rng(1)
    m_p, v_p = 0.2 * rng.standard_normal(n), np.exp(0.3 * rng.standard_normal(n))  # fixed regulariser predictive

    leaves = {k: dev(v).requires_grad_(True) for k, v in dict(w=w, b=b, ls=np.array(ls), lc=np.array(lc), mq=m_q).items()}
    base = S.PolynomialKernel(polynomial_degree=1)
    kernel = S.CustomMappingKernel(base_kernel=base,
                                   feature_mapping=lambda p, xx: torch.tanh(xx @ p["w"] + p["b"]))
    sparse = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=z, diagonal_regularisation=reg)
    amean = S.CustomMean(mean_function=lambda p, xx: p[: xx.shape[0]])
    approx = S.ApproximateGPRegression(mean=amean, kernel=sparse)
    ap = approx.Parameters.construct(
        mean=amean.Parameters.construct(custom=leaves["mq"]),
        kernel=sparse.Parameters.construct(base_kernel=kernel.Parameters.construct(
            base_kernel=base.Parameters.construct(log_constant=leaves["lc"], log_scaling=leaves["ls"]),
            feature_mapping={"w": leaves["w"], "b": leaves["b"]})))
    # regulariser: any GP whose predictive is fixed - a prior-mode exact GP over ARD gives (m_p, v_p) through CustomMean
    pmean = S.CustomMean(mean_function=lambda p, xx: p[: xx.shape[0]])

    class FixedPredictive:  # the smallest GPBase stand-in: calculate_prior returns the fixed (m_p, v_p)
        def calculate_prior(self, parameters, x, full_covariance):
            return dev(m_p), dev(v_p)

        def _to_parameters(self, parameters):
            return parameters

    reg_term = S.projected.ProjectedKLRegularisation(gp=approx, regulariser=FixedPredictive(),
                                                     regulariser_parameters=None, mode="prior")
    gvi = S.GeneralisedVariationalInference(regularisation=reg_term,
                                            empirical_risk=S.NegativeLogLikelihood(gp=approx))
    loss = gvi.calculate_loss(ap, dev(x), dev(y))
    loss.backward()

    t = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True)
         for k, v in dict(w=w, b=b, ls=ls, lc=lc, mq=m_q).items()}
    xt, zt, yt = torch.tensor(x), torch.tensor(z), torch.tensor(y)
    phi = lambda a: torch.tanh(a @ t["w"] + t["b"])
    kern = lambda a, c: torch.exp(t["ls"]) * (phi(a) @ phi(c).T) + torch.exp(t["lc"])
    k_zz = kern(zt, zt)
    a_mat = k_zz + reg * torch.trace(k_zz) / m * torch.eye(m, dtype=torch.float64)
    k_xz = kern(xt, zt)
    k_diag = torch.exp(t["ls"]) * (phi(xt) ** 2).sum(-1) + torch.exp(t["lc"])
    v_q = k_diag - (k_xz * torch.cholesky_solve(k_xz.T, torch.linalg.cholesky(a_mat)).T).sum(-1)
    ref = T.gaussian_nll(yt, t["mq"], v_q) + T.regulariser("kl", torch.tensor(m_p), torch.tensor(v_p), t["mq"],
                                                           v_q).mean()
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-9 * abs(float(ref.detach()))
    for k in ("w", "b", "ls", "lc", "mq"):
        got, want = host(leaves[k].grad).reshape(-1), t[k].grad.numpy().reshape(-1)
        assert np.max(np.abs(got - want)) <= 1e-8 * max(np.max(np.abs(want)), 1e-3), k


def test_exact_gp_nll_gradients_over_feature_map_kernel_match_autograd_twin(S):
    """train_regulariser_gp's objective (exact-GP NLL on the inducing data) over the feature-map kernel: gradients w.r.t.
    the network weights, the polynomial kernel's scalars, log_observation_noise and the constant mean."""
    from oracle import torch_twin as T

    m, d, f = 120, 4, 10
    _, z, _, yz, w, b, _ = _feature_problem(10, m, d, f, seed=3)
    vals = dict(w=w, b=b, ls=np.array(-0.3), lc=np.array(0.1), noise=np.array(-1.2), const=np.array(0.2))
    leaves = {k: dev(v).requires_grad_(True) for k, v in vals.items()}
    base = S.PolynomialKernel(polynomial_degree=2)
    kernel = S.CustomMappingKernel(base_kernel=base,
                                   feature_mapping=lambda p, xx: torch.tanh(xx @ p["w"] + p["b"]))
    const = S.ConstantMean()
    gp = S.GPRegression(mean=const, kernel=kernel, x=z, y=yz)
    params = gp.Parameters.construct(
        log_observation_noise=leaves["noise"], mean=const.Parameters.construct(constant=leaves["const"]),
        kernel=kernel.Parameters.construct(
            base_kernel=base.Parameters.construct(log_constant=leaves["lc"], log_scaling=leaves["ls"]),
            feature_mapping={"w": leaves["w"], "b": leaves["b"]}))
    loss = S.NegativeLogLikelihood(gp=gp).calculate_empirical_risk(params, dev(z), dev(yz))
    loss.backward()

    t = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for k, v in vals.items()}
    zt, yzt = torch.tensor(z), torch.tensor(yz)
    phi = torch.tanh(zt @ t["w"] + t["b"])
    k_zz = (torch.exp(t["ls"]) * (phi @ phi.T) + torch.exp(t["lc"])) ** 2
    chol = torch.linalg.cholesky(k_zz + torch.exp(t["noise"]) * torch.eye(m, dtype=torch.float64))
    mean = k_zz @ torch.cholesky_solve(yzt.reshape(-1, 1), chol).reshape(-1) + t["const"]
    var = torch.diagonal(k_zz) - (k_zz * torch.cholesky_solve(k_zz, chol)).sum(0)
    ref = T.gaussian_nll(yzt, mean, var)
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-9 * abs(float(ref.detach()))
    for k in vals:
        got, want = host(leaves[k].grad).reshape(-1), t[k].grad.numpy().reshape(-1)
        assert np.max(np.abs(got - want)) <= 1e-8 * max(np.max(np.abs(want)), 1e-3), k


def test_fixed_sparse_and_svgp_kernels_over_polynomial_kernels(S):
    """FixedSparsePosteriorKernel and the SVGP family over a NON-ARD regulariser kernel (the reference pairs every
    approximate kernel with every regulariser kernel, experiments/regression/base_configs): values against the oracle,
    gradients (base-kernel scalars; el-matrix parameters) against torch CPU autograd twins."""
    from gvi_gaussian_process_b200.src.kernels.approximate import CholeskySVGPKernel, FixedSparsePosteriorKernel

    rng = np.random.default_rng(12)
    n, m, d, reg = 900, 20, 3, 0.05
    x, z = rng.standard_normal((n, d)) / np.sqrt(d), rng.standard_normal((m, d)) / np.sqrt(d)
    x_train = rng.standard_normal((60, d)) / np.sqrt(d)
    lc_r, ls_r, lc_b, ls_b = 0.1, -0.2, -0.3, 0.25
    reg_gram = lambda a, b: O.polynomial_gram(lc_r, ls_r, 2, a, b)
    base_gram = lambda a, b: O.polynomial_gram(lc_b, ls_b, 2, a, b)
    poly = S.PolynomialKernel(polynomial_degree=2)
    reg_params = poly.Parameters.construct(log_constant=lc_r, log_scaling=ls_r)
    g_w = rng.standard_normal(n)  # a fixed linear functional of the variances: loss = <g_w, var>

    # ---- FixedSparsePosteriorKernel: frozen Cholesky from the regulariser kernel, cross-Grams from the base kernel
    leaves = {k: dev(np.array(v)).requires_grad_(True) for k, v in dict(lc=lc_b, ls=ls_b).items()}
    fixed = FixedSparsePosteriorKernel(regulariser_kernel=poly, regulariser_kernel_parameters=reg_params,
                                       base_kernel=poly, inducing_points=z, diagonal_regularisation=reg,
                                       is_diagonal_regularisation_absolute_scale=True)
    fp = fixed.Parameters.construct(base_kernel=poly.Parameters.construct(log_constant=leaves["lc"],
                                                                          log_scaling=leaves["ls"]))
    var = fixed.calculate_gram(fp, x, full_covariance=False)
    ref = np.diag(O.fixed_sparse_posterior_gram(base_gram, reg_gram, z, x, None, reg, True))
    assert np.max(np.abs(host(var) - ref)) <= TOL * np.max(np.abs(np.diag(base_gram(x, x))))
    (dev(g_w) * var).sum().backward()
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in dict(lc=lc_b, ls=ls_b).items()}
    xt, zt = torch.tensor(x), torch.tensor(z)
    kb = lambda a, b: (torch.exp(t["ls"]) * (a @ b.T) + torch.exp(t["lc"])) ** 2
    a_inv = torch.tensor(np.linalg.inv(reg_gram(z, z) + reg * np.eye(m)))
    k_xz = kb(xt, zt)
    var_t = torch.diagonal(kb(xt, xt)) - ((k_xz @ a_inv) * k_xz).sum(-1)
    (torch.tensor(g_w) * var_t).sum().backward()
    for k in ("lc", "ls"):
        assert abs(float(leaves[k].grad) - float(t[k].grad)) <= 1e-8 * max(abs(float(t[k].grad)), 1e-3), k

    # ---- CholeskySVGPKernel over the polynomial regulariser kernel: second triangular product on the supplied K tile
    svgp = CholeskySVGPKernel(regulariser_kernel=poly, regulariser_kernel_parameters=reg_params,
                              log_observation_noise=np.log(0.5), inducing_points=z, training_points=x_train,
                              diagonal_regularisation=reg, is_diagonal_regularisation_absolute_scale=True)
    lower = np.tril(0.1 * rng.standard_normal((m, m)), k=-1)
    log_diag = -1.0 + 0.1 * rng.standard_normal(m)
    el = {"lower": dev(lower).requires_grad_(True), "log_diag": dev(log_diag).requires_grad_(True)}
    sp = svgp.Parameters.construct(el_matrix_lower_triangle=el["lower"], el_matrix_log_diagonal=el["log_diag"])
    var = svgp.calculate_gram(sp, x, full_covariance=False)
    sigma = O.cholesky_sigma(lower, log_diag)
    ref = np.diag(O.svgp_gram(reg_gram, z, sigma, x, None, reg, True))
    assert np.max(np.abs(host(var) - ref)) <= TOL * np.max(np.abs(ref))
    (dev(g_w) * var).sum().backward()
    lt = torch.tensor(lower, requires_grad=True)
    dt = torch.tensor(log_diag, requires_grad=True)
    e_t = torch.tril(lt, diagonal=-1) + torch.diag(torch.exp(dt))
    k_xz_r = torch.tensor(reg_gram(x, z))
    var_s = ((k_xz_r @ (e_t.T @ e_t)) * k_xz_r).sum(-1)  # the only part that depends on the el-matrix parameters
    (torch.tensor(g_w) * var_s).sum().backward()
    assert np.max(np.abs(host(el["lower"].grad) - lt.grad.numpy())) <= 1e-8 * np.max(np.abs(lt.grad.numpy()))
    assert np.max(np.abs(host(el["log_diag"].grad) - dt.grad.numpy())) <= 1e-8 * np.max(np.abs(dt.grad.numpy()))
    # generate_parameters() without arguments: the reference's initialisation (svgp/base.py:57-73) over this kernel.  It has
    # no jitter, so the inducing set must not exceed the kernel's rank (10 monomials of degree <= 2 in 3-D): 6 points
    small = CholeskySVGPKernel(regulariser_kernel=poly, regulariser_kernel_parameters=reg_params,
                               log_observation_noise=np.log(0.5), inducing_points=z[:6], training_points=x_train,
                               diagonal_regularisation=reg, is_diagonal_regularisation_absolute_scale=True)
    init = small.generate_parameters()
    inv_chol = O.svgp_initial_inverse_cholesky(reg_gram, z[:6], x_train, np.log(0.5))
    assert rel_norm(host(init.el_matrix_lower_triangle), np.tril(inv_chol, k=-1)) <= 1e-8
