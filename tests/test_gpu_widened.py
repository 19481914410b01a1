"""GPU parity tests of SURVEY.md section 8(f)-2 (classification) and of the CUDA vector-Jacobian products behind the
torch autograd front end.  Values against the NumPy oracle (<= 1e-10), gradients against the torch-f64 autograd twin
(<= 1e-8), the reference's own classification goldens through the real product path (all-ones Gram = ARD with
exp(log_lengthscales) = 0)."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as O
from oracle import torch_twin as T

pytestmark = pytest.mark.gpu
TOL, GTOL = 1e-10, 1e-8


@pytest.fixture(scope="module")
def S():
    import types

    from gvi_gaussian_process_b200 import _lib
    from gvi_gaussian_process_b200.src import _autograd, _lib_proxy
    from gvi_gaussian_process_b200.src.empirical_risks import CrossEntropy, NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
    from gvi_gaussian_process_b200.src.gps import (ApproximateGPClassification, ApproximateGPRegression,
                                                   GPClassification, GPRegression)
    from gvi_gaussian_process_b200.src.kernels import MultiOutputKernel, TemperedKernel
    from gvi_gaussian_process_b200.src.kernels.approximate import FixedSparsePosteriorKernel, SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean, MultiOutputMean
    from gvi_gaussian_process_b200.src.regularisations import (
        GaussianSquaredDifferenceRegularisation, GaussianWassersteinRegularisation,
        MultinomialWassersteinRegularisation)
    from gvi_gaussian_process_b200.src.regularisations import projected

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.load()
    return types.SimpleNamespace(**locals())


def dev(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
    return t.requires_grad_(True) if grad else t


def host(t):
    return t.detach().cpu().numpy()


def cpu(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
    return t.requires_grad_(True) if grad else t


def rel_norm(a, ref):
    return float(np.max(np.abs(a - ref)) / max(np.max(np.abs(ref)), 1e-300))


HERMITE = np.polynomial.hermite.hermgauss(50)


def random_gaussians(k, n, seed, spread=1.0):
    rng = np.random.default_rng(seed)
    return spread * rng.standard_normal((k, n)), np.exp(rng.uniform(-3.0, 1.0, (k, n)))


# ---- class probabilities: value and vector-Jacobian product ------------------------------------------------------
@pytest.mark.parametrize("k,n,spread", [(2, 1, 1.0), (4, 37, 1.0), (10, 5000, 1.0), (10, 300, 8.0), (33, 64, 0.5)])
def test_class_probabilities_parity(S, k, n, spread):
    mean, var = random_gaussians(k, n, seed=k * 1000 + n, spread=spread)
    roots, weights = dev(HERMITE[0]), dev(HERMITE[1])
    p = host(S._lib_proxy.class_probabilities(dev(mean), dev(var), roots, weights, 0.01, 1e-10))
    ref = O.classification_probabilities(mean, var)
    assert p.shape == (n, k)
    assert np.max(np.abs(p - ref)) <= TOL * np.max(np.abs(ref))
    # rows of the s-matrix are class probabilities of a partition up to the quadrature error of the 50-node rule
    assert np.all(p > 0) and np.all(np.abs(p.sum(axis=1) - 1.0) < 0.05)


@pytest.mark.parametrize("k,n,spread", [(2, 5, 1.0), (4, 201, 1.0), (10, 777, 1.0), (10, 64, 6.0)])
def test_class_probabilities_vjp_matches_autograd_twin(S, k, n, spread):
    mean, var = random_gaussians(k, n, seed=7 * k + n, spread=spread)
    rng = np.random.default_rng(1)
    dprob = rng.standard_normal((n, k))
    roots, weights = dev(HERMITE[0]), dev(HERMITE[1])
    dm, dv = S._lib_proxy.class_probabilities_grad(dev(mean), dev(var), roots, weights, 0.01, 1e-10, dev(dprob))
    tm, tv = cpu(mean, True), cpu(var, True)
    (T.classification_probabilities(tm, tv, HERMITE[0], HERMITE[1]) * cpu(dprob)).sum().backward()
    assert rel_norm(host(dm), tm.grad.numpy()) <= GTOL
    assert rel_norm(host(dv), tv.grad.numpy()) <= GTOL


def test_multinomial_risks_value_and_gradient(S):
    rng = np.random.default_rng(3)
    n, k = 4001, 7
    q = rng.dirichlet(np.ones(k), n)
    p = rng.dirichlet(np.ones(k), n)
    y = np.eye(k)[rng.integers(0, k, n)]
    for power in (1, 2, 3, 4):
        out4 = host(S._lib_proxy.multinomial_risk(dev(q), dev(y), dev(p), power))
        assert abs(out4[0] - O.multinomial_nll(q, y)) <= TOL * abs(O.multinomial_nll(q, y))
        assert abs(out4[1] / n - O.multinomial_wasserstein(p, q, power)) <= TOL
        assert out4[2] == n
        assert abs(out4[3] / n - O.cross_entropy_from_probabilities(q, y)) <= TOL
        w4 = np.array([0.7, -1.3, 0.0, 2.1])
        d = host(S._lib_proxy.multinomial_risk_grad(dev(q), dev(y), dev(p), power, dev(w4)))
        tq = cpu(q, True)
        (w4[0] * T.multinomial_nll(tq, cpu(y)) + w4[1] * n * T.multinomial_wasserstein(cpu(p), tq, power)
         + w4[3] * n * T.cross_entropy_from_probabilities(tq, cpu(y))).backward()
        assert rel_norm(d, tq.grad.numpy()) <= GTOL, power
    # soft labels (the reference's tests use them) and the y-only / p-only calls
    ys = rng.dirichlet(np.ones(k), n)
    out4 = host(S._lib_proxy.multinomial_risk(dev(q), dev(ys), None, 2))
    assert abs(out4[0] - O.multinomial_nll(q, ys)) <= TOL * abs(out4[0]) and out4[1] == 0.0
    assert abs(out4[3] / n - O.cross_entropy_from_probabilities(q, ys)) <= TOL
    out4 = host(S._lib_proxy.multinomial_risk(dev(q), None, dev(q), 2))
    assert out4[0] == 0.0 and out4[1] == 0.0


@pytest.mark.parametrize("kind", ["bhattacharyya", "gaussian_wasserstein", "hellinger", "kl", "renyi", "gwi_no_eig",
                                  "squared_difference"])
def test_pointwise_risk_gradient_kernel(S, kind):
    rng = np.random.default_rng(5)
    n = 3001
    y, mq, mp = rng.standard_normal((3, n))
    vq, vp = np.exp(rng.uniform(-2, 1, (2, n)))
    w2 = np.array([0.3, 1.7])
    code = S._lib_proxy.REG_KINDS[kind]
    dm, dv = S._lib_proxy.risk_grad(code, 0.5, dev(y), dev(mq), dev(vq), dev(mp), dev(vp), dev(w2))
    tm, tv = cpu(mq, True), cpu(vq, True)
    (w2[0] * n * T.gaussian_nll(cpu(y), tm, tv) + w2[1] * n * T.regulariser(kind, cpu(mp), cpu(vp), tm, tv, 0.5)
     ).backward()
    assert rel_norm(host(dm), tm.grad.numpy()) <= GTOL
    assert rel_norm(host(dv), tv.grad.numpy()) <= GTOL


# ---- the variance VJP (gvi_gp_predict_grad_f64) -------------------------------------------------------------------
@pytest.mark.parametrize("n,m,d,is_abs,scalar_ll", [(60, 12, 1, False, True), (700, 65, 3, False, False),
                                                    (2000, 200, 8, True, False),
                                                    # D > 32: the GEMM formulation of grad_wide.cu (3000 > one chunk)
                                                    (900, 70, 40, False, False), (1500, 130, 100, True, False),
                                                    (3000, 96, 48, False, True)])
def test_variance_vjp_matches_autograd_twin(S, n, m, d, is_abs, scalar_ll):
    rng = np.random.default_rng(n + m)
    x = rng.standard_normal((n, d))
    z = x[rng.permutation(n)[:m]].copy()
    ls = 0.3
    ll = np.float64(np.log(1.0 / d)) if scalar_ll else np.log(rng.uniform(0.5, 1.5, d) / d)
    c = rng.standard_normal(n)
    lam = 1e-3
    kernel = S.ARDKernel(number_of_dimensions=d)
    sp = S.SparsePosteriorKernel(base_kernel=kernel, inducing_points=z, diagonal_regularisation=lam,
                                 is_diagonal_regularisation_absolute_scale=is_abs)
    t_ls, t_ll = dev(ls, True), dev(ll, True)
    var = sp.calculate_gram({"base_kernel": {"log_scaling": t_ls, "log_lengthscales": t_ll}}, x, x,
                            full_covariance=False)
    (var * dev(c)).sum().backward()
    c_ls, c_ll = cpu(ls, True), cpu(ll, True)
    ref_var = T.sparse_posterior_diag(c_ls, c_ll, cpu(z), cpu(x), lam, is_abs)
    (ref_var * cpu(c)).sum().backward()
    assert rel_norm(host(var), ref_var.detach().numpy()) <= TOL
    assert abs(float(t_ls.grad) - float(c_ls.grad)) <= GTOL * abs(float(c_ls.grad))
    assert rel_norm(host(t_ll.grad).reshape(-1), c_ll.grad.numpy().reshape(-1)) <= GTOL


# ---- the reference's classification goldens through the product path ---------------------------------------------
X_TRAIN = np.array([[1.0, 3.0, 2.0], [1.5, 1.5, 9.5]])
X2 = np.array([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5]])
Y_CLS = np.array([[0.5, 0.1, 0.2, 0.2], [0.1, 0.2, 0.3, 0.4]])
ONES = {"log_scaling": 0.0, "log_lengthscales": np.full(3, -np.inf)}
K_CLS = 4


def ones_models(S, x_train, noise):
    kern = S.MultiOutputKernel(kernels=[S.ARDKernel(number_of_dimensions=3) for _ in range(K_CLS)])
    mean = S.ConstantMean(number_output_dimensions=K_CLS)
    exact = S.GPClassification(mean=mean, kernel=kern, x=x_train, y=Y_CLS)
    approx = S.ApproximateGPClassification(mean=mean, kernel=kern)
    params = {"log_observation_noise": noise, "mean": {"constant": np.ones(K_CLS)},
              "kernel": {"kernels": [ONES] * K_CLS}}
    return kern, mean, exact, approx, params


def test_reference_goldens_classification_prediction(S):
    noise = np.log(np.array([1.0, 0.5, 0.2, 0.9]))
    kern, mean, exact, approx, params = ones_models(S, X2, noise)
    golden = np.array([0.29579238035540445, 0.19449721942556136, 0.2159886961681648, 0.29372170534428094])
    p = host(exact.predict_probability(params, X_TRAIN).probabilities)  # tests/gps/test_classification.py:23-88
    np.testing.assert_allclose(p, np.tile(golden, (2, 1)), rtol=1e-12)
    q = host(approx.predict_probability(params, X_TRAIN).probabilities)  # :91-128
    np.testing.assert_allclose(q, 0.25, rtol=1e-12)
    # :209-265 tempered approximate GP; the noise handed in is dropped
    temper = np.log(np.array([0.5, 1.0, 3.0, 4.0]))
    tk = S.TemperedKernel(base_kernel=kern, base_kernel_parameters=params["kernel"], number_output_dimensions=K_CLS)
    tgp = S.ApproximateGPClassification(mean=mean, kernel=tk)
    tparams = {"log_observation_noise": noise, "mean": params["mean"], "kernel": {"log_tempering_factor": temper}}
    tq = host(tgp.predict_probability(tparams, X_TRAIN).probabilities)
    np.testing.assert_allclose(tq, np.tile([0.1690632, 0.20674086, 0.29816303, 0.32603688], (2, 1)), rtol=2e-7)
    # tempered exact GP (ones kernel; the reference's golden uses the identity mock, so compare with the oracle)
    tex = S.GPClassification(mean=mean, kernel=tk, x=X2, y=Y_CLS)
    tp = host(tex.predict_probability(tparams, X_TRAIN).probabilities)
    ones_gram = lambda a, b: np.ones((a.shape[0], b.shape[0]))
    gf = [(lambda a, b, t=t: np.exp(t) * ones_gram(a, b)) for t in temper]
    df = [(lambda a, t=t: np.exp(t) * np.ones(a.shape[0])) for t in temper]
    m, v = O.multi_output_exact_gp_posterior(gf, df, np.ones((K_CLS, 2)), X2, Y_CLS, noise, X_TRAIN)
    np.testing.assert_allclose(tp, O.classification_probabilities(m, v), rtol=1e-10)


def test_reference_goldens_classification_risks_and_regularisers(S):
    noise = np.log(np.array([0.1, 0.2, 0.4, 1.8]))
    kern, mean, exact, approx, params = ones_models(S, X_TRAIN, noise)
    nll = S.NegativeLogLikelihood(gp=approx)  # tests/test_empirical_risks.py:272-316
    assert np.isclose(float(nll.calculate_empirical_risk(params, X2, Y_CLS)), 2.77258869, rtol=1e-8)
    ce = S.CrossEntropy(gp=approx)  # :319-370
    assert np.isclose(float(ce.calculate_empirical_risk(params, X2, np.array([[0, 0, 0, 1], [0, 1, 0, 0]]))),
                      1.3862944, rtol=1e-7)
    cases = [("ProjectedBhattacharyyaRegularisation", 0.24135331), ("ProjectedGaussianWassersteinRegularisation",
             0.4287477), ("ProjectedHellingerRegularisation", 0.208868), ("ProjectedKLRegularisation", 0.61610398),
             ("ProjectedRenyiRegularisation", 0.49202457)]
    for cls, golden in cases:  # tests/regularisations/projected/test_projected_*.py:~316
        reg = getattr(S.projected, cls)(gp=approx, regulariser=exact, regulariser_parameters=params, mode="posterior")
        assert np.isclose(float(reg.calculate_regularisation(params, X2)), golden, rtol=2e-6), cls
        same = getattr(S.projected, cls)(gp=approx, regulariser=approx, regulariser_parameters=params,
                                         mode="posterior")
        assert abs(float(same.calculate_regularisation(params, X2))) <= 1e-12
    # the reference's golden uses full covariances; with the all-ones kernel every entry of C_p - C_q equals the
    # diagonal one, so the diagonal form (the one on the hot path) has the same mean
    gsd = S.GaussianSquaredDifferenceRegularisation(gp=approx, regulariser=exact, regulariser_parameters=params,
                                                    mode="posterior", full_covariance=False)
    assert np.isclose(float(gsd.calculate_regularisation(params, X2)), 0.71837243, rtol=1e-7)
    mw = S.MultinomialWassersteinRegularisation(gp=approx, regulariser=exact, regulariser_parameters=params,
                                                mode="posterior")  # test_multinomial_wasserstein.py:130-197
    assert np.isclose(float(mw.calculate_regularisation(params, X2)), 0.12691447, rtol=1e-7)
    zero = S.MultinomialWassersteinRegularisation(gp=approx, regulariser=approx, regulariser_parameters=params,
                                                  mode="posterior", power=4)
    assert float(zero.calculate_regularisation(params, X2)) == 0.0
    with pytest.raises(NotImplementedError):
        S.MultinomialWassersteinRegularisation(gp=approx, regulariser=approx, regulariser_parameters=params,
                                               mode="prior").calculate_regularisation(params, X2)


# ---- end to end: K-class objective with per-class ARD sparse posteriors, value and loss.backward() ---------------
def make_classification_problem(n, d, k, m, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    labels = rng.integers(0, k, n)
    y = np.eye(k)[labels]
    zs = [x[rng.permutation(n)[:m]].copy() for _ in range(k)]        # per-class inducing sets (mnist/main.py:219-231)
    zi = rng.permutation(n)[: m + 3]
    z_all, y_all = x[zi].copy(), y[zi].copy()                        # the regulariser GP's combined inducing data
    ls = rng.uniform(-0.2, 0.2, k)
    ll = [np.log(rng.uniform(0.5, 1.5, d) / d) for _ in range(k)]
    ls_p = rng.uniform(-0.2, 0.2, k)
    ll_p = [np.log(rng.uniform(0.5, 1.5, d) / d) for _ in range(k)]
    noise_p = np.log(rng.uniform(0.3, 1.0, k))
    w1 = rng.standard_normal((d, 6)) / np.sqrt(d)
    w2 = rng.standard_normal((6, k)) / np.sqrt(6)
    return dict(x=x, y=y, zs=zs, z_all=z_all, y_all=y_all, ls=ls, ll=ll, ls_p=ls_p, ll_p=ll_p, noise_p=noise_p,
                w1=w1, w2=w2)


def oracle_classification(pr, lam):
    k = pr["y"].shape[1]
    n = pr["x"].shape[0]
    nn = np.tanh(pr["x"] @ pr["w1"]) @ pr["w2"]                       # (n, k)
    m_q = nn.reshape(k, -1)                                           # the reference RESHAPES (means/base.py:97-100)
    v_q = np.stack([O.sparse_posterior_diag(lambda a, b, i=i: O.ard_gram(pr["ls"][i], pr["ll"][i], a, b),
                                            lambda a, i=i: O.ard_gram_diag(pr["ls"][i], pr["ll"][i], a),
                                            pr["zs"][i], pr["x"], lam, False) for i in range(k)])
    gf = [(lambda a, b, i=i: O.ard_gram(pr["ls_p"][i], pr["ll_p"][i], a, b)) for i in range(k)]
    df = [(lambda a, i=i: O.ard_gram_diag(pr["ls_p"][i], pr["ll_p"][i], a)) for i in range(k)]
    m_p, v_p = O.multi_output_exact_gp_posterior(gf, df, np.zeros((k, n)), pr["z_all"], pr["y_all"], pr["noise_p"],
                                                 pr["x"])
    return m_q, v_q, m_p, v_p


def build_classification_models(S, pr, lam, grad=False):
    k = pr["y"].shape[1]
    d = pr["x"].shape[1]
    kern_q = S.MultiOutputKernel(kernels=[
        S.SparsePosteriorKernel(base_kernel=S.ARDKernel(number_of_dimensions=d), inducing_points=pr["zs"][i],
                                diagonal_regularisation=lam) for i in range(k)])
    w1, w2 = dev(pr["w1"], grad), dev(pr["w2"], grad)
    mean_q = S.CustomMean(mean_function=lambda p, xx: torch.tanh(xx @ p[0]) @ p[1], number_output_dimensions=k)
    approx = S.ApproximateGPClassification(mean=mean_q, kernel=kern_q)
    leaves = [(dev(pr["ls"][i], grad), dev(pr["ll"][i], grad)) for i in range(k)]
    params = {"mean": {"custom": (w1, w2)},
              "kernel": {"kernels": [{"base_kernel": {"log_scaling": a, "log_lengthscales": b}} for a, b in leaves]}}
    kern_p = S.MultiOutputKernel(kernels=[S.ARDKernel(number_of_dimensions=d) for _ in range(k)])
    exact = S.GPClassification(mean=S.ConstantMean(number_output_dimensions=k), kernel=kern_p, x=pr["z_all"],
                               y=pr["y_all"])
    eparams = {"log_observation_noise": pr["noise_p"], "mean": {"constant": np.zeros(k)},
               "kernel": {"kernels": [{"log_scaling": pr["ls_p"][i], "log_lengthscales": pr["ll_p"][i]}
                                      for i in range(k)]}}
    return approx, params, exact, eparams, leaves, (w1, w2)


@pytest.mark.parametrize("n,d,k,m", [(500, 2, 3, 20), (3000, 8, 10, 64)])
def test_classification_objective_matches_oracle(S, n, d, k, m):
    lam = 1e-4
    pr = make_classification_problem(n, d, k, m, seed=n + k)
    approx, params, exact, eparams, _, _ = build_classification_models(S, pr, lam)
    m_q, v_q, m_p, v_p = oracle_classification(pr, lam)
    q = O.classification_probabilities(m_q, v_q)
    p = O.classification_probabilities(m_p, v_p)
    got = host(approx.predict_probability(params, pr["x"]).probabilities)
    assert np.max(np.abs(got - q)) <= TOL
    assert np.max(np.abs(host(exact.predict_probability(eparams, pr["x"]).probabilities) - p)) <= TOL
    nll = S.NegativeLogLikelihood(gp=approx)
    ref_nll = O.multinomial_nll(q, pr["y"])
    assert abs(float(nll.calculate_empirical_risk(params, pr["x"], pr["y"])) - ref_nll) <= TOL * abs(ref_nll)
    regs = {
        "kl": S.projected.ProjectedKLRegularisation,
        "renyi": S.projected.ProjectedRenyiRegularisation,
        "hellinger": S.projected.ProjectedHellingerRegularisation,
    }
    for kind, cls in regs.items():
        reg = cls(gp=approx, regulariser=exact, regulariser_parameters=eparams, mode="posterior")
        ref = O.projected_regularisation(kind, m_p, v_p, m_q, v_q)
        assert abs(float(reg.calculate_regularisation(params, pr["x"])) - ref) <= TOL * abs(ref), kind
        gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=nll)
        assert abs(float(gvi.calculate_loss(params, pr["x"], pr["y"])) - (ref_nll + ref)) <= TOL * abs(ref_nll + ref)
    gwi = S.GaussianWassersteinRegularisation(gp=approx, regulariser=exact, regulariser_parameters=eparams,
                                              mode="posterior")
    assert abs(float(gwi.calculate_regularisation(params, pr["x"])) - O.gwi_no_eig(m_p, v_p, m_q, v_q)) <= TOL * 10
    mw = S.MultinomialWassersteinRegularisation(gp=approx, regulariser=exact, regulariser_parameters=eparams,
                                                mode="posterior")
    assert abs(float(mw.calculate_regularisation(params, pr["x"])) - O.multinomial_wasserstein(p, q)) <= TOL


@pytest.mark.parametrize("reg_name", ["kl", "multinomial_wasserstein", "gwi"])
def test_classification_backward_matches_autograd_twin(S, reg_name):
    n, d, k, m, lam = 400, 3, 4, 24, 1e-3
    pr = make_classification_problem(n, d, k, m, seed=11)
    approx, params, exact, eparams, leaves, (w1, w2) = build_classification_models(S, pr, lam, grad=True)
    nll = S.NegativeLogLikelihood(gp=approx)
    if reg_name == "kl":
        reg = S.projected.ProjectedKLRegularisation(gp=approx, regulariser=exact, regulariser_parameters=eparams,
                                                    mode="posterior")
    elif reg_name == "gwi":
        reg = S.GaussianWassersteinRegularisation(gp=approx, regulariser=exact, regulariser_parameters=eparams,
                                                  mode="posterior")
    else:
        reg = S.MultinomialWassersteinRegularisation(gp=approx, regulariser=exact, regulariser_parameters=eparams,
                                                     mode="posterior")
    gvi = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=nll)
    loss = gvi.calculate_loss(params, pr["x"], pr["y"])
    loss.backward()
    # twin
    _, _, m_p, v_p = oracle_classification(pr, lam)
    tx, ty = cpu(pr["x"]), cpu(pr["y"])
    tw1, tw2 = cpu(pr["w1"], True), cpu(pr["w2"], True)
    tls = [cpu(pr["ls"][i], True) for i in range(k)]
    tll = [cpu(pr["ll"][i], True) for i in range(k)]
    m_q = (torch.tanh(tx @ tw1) @ tw2).reshape(k, -1)
    v_q = torch.stack([T.sparse_posterior_diag(tls[i], tll[i], cpu(pr["zs"][i]), tx, lam, False) for i in range(k)])
    q = T.classification_probabilities(m_q, v_q, HERMITE[0], HERMITE[1])
    ref = T.multinomial_nll(q, ty)
    if reg_name == "kl":
        ref = ref + T.regulariser("kl", cpu(m_p), cpu(v_p), m_q, v_q)
    elif reg_name == "gwi":
        ref = ref + T.regulariser("gwi_no_eig", cpu(m_p), cpu(v_p), m_q, v_q)
    else:
        ref = ref + T.multinomial_wasserstein(cpu(O.classification_probabilities(m_p, v_p)), q)
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
    assert rel_norm(host(w1.grad), tw1.grad.numpy()) <= GTOL
    assert rel_norm(host(w2.grad), tw2.grad.numpy()) <= GTOL
    for i in range(k):
        assert abs(float(leaves[i][0].grad) - float(tls[i].grad)) <= GTOL * max(abs(float(tls[i].grad)), 1e-3), i
        assert rel_norm(host(leaves[i][1].grad), tll[i].grad.numpy()) <= GTOL, i


def test_classification_wide_inputs_mnist_shaped(S):
    """config #4 in miniature: D = 784 inputs in [0, 1], K = 10 classes, per-class ARD sparse posteriors; the K x N
    predictive goes through the tensor-pipe Gram + tile kernel, the probabilities through the quadrature kernel."""
    rng = np.random.default_rng(4)
    n, d, k, m, lam = 600, 784, 10, 96, 1e-5
    x = rng.uniform(0, 1, (n, d))
    y = np.eye(k)[rng.integers(0, k, n)]
    zs = [x[rng.permutation(n)[:m]].copy() for _ in range(k)]
    ll = np.full(d, np.log(1.0 / d))
    kern = S.MultiOutputKernel(kernels=[
        S.SparsePosteriorKernel(base_kernel=S.ARDKernel(number_of_dimensions=d), inducing_points=zs[i],
                                diagonal_regularisation=lam) for i in range(k)])
    w = dev(rng.standard_normal((d, k)) / np.sqrt(d))
    mean = S.CustomMean(mean_function=lambda p, xx: xx @ p, number_output_dimensions=k)
    gp = S.ApproximateGPClassification(mean=mean, kernel=kern)
    params = {"mean": {"custom": w}, "kernel": {"kernels": [{"base_kernel": {"log_scaling": 0.1 * i,
                                                                            "log_lengthscales": ll}} for i in range(k)]}}
    got = host(gp.predict_probability(params, x).probabilities)
    m_q = (x @ host(w)).reshape(k, -1)
    v_q = np.stack([O.sparse_posterior_diag(lambda a, b, i=i: O.ard_gram(0.1 * i, ll, a, b),
                                            lambda a, i=i: O.ard_gram_diag(0.1 * i, ll, a), zs[i], x, lam, False)
                    for i in range(k)])
    ref = O.classification_probabilities(m_q, v_q)
    assert np.max(np.abs(got - ref)) <= TOL
    nll = float(S.NegativeLogLikelihood(gp=gp).calculate_empirical_risk(params, x, y))
    assert abs(nll - O.multinomial_nll(ref, y)) <= TOL * abs(nll)


# ---- section 8(f)-3: the exact-GP NLL training step on the inducing set, value and gradient -----------------------
@pytest.mark.parametrize("n,m,d,scalar_ll", [(40, 40, 1, True), (300, 200, 3, False), (5000, 700, 8, False),
                                             # D > 32: GEMM formulation (grad_wide.cu); 2500 rows > one chunk
                                             (300, 200, 40, False), (2500, 130, 64, True)])
def test_exact_gp_nll_step_backward_matches_autograd_twin(S, n, m, d, scalar_ll):
    rng = np.random.default_rng(n + d)
    z = rng.standard_normal((m, d))
    wv = rng.standard_normal(d)
    yz = np.sin(z @ wv) + 0.1 * rng.standard_normal(m)
    if n == m:  # the reference trains on the inducing data itself (trainers.py:69-80)
        x, y = z.copy(), yz.copy()
    else:
        x = rng.standard_normal((n, d))
        y = np.sin(x @ wv) + 0.1 * rng.standard_normal(n)
    ls, noise, const = 0.2, np.log(0.3), 0.15
    ll = np.float64(np.log(1.0 / d)) if scalar_ll else np.log(rng.uniform(0.5, 1.5, d) / d)
    kernel = S.ARDKernel(number_of_dimensions=d)
    mean = S.ConstantMean()
    gp = S.GPRegression(mean=mean, kernel=kernel, x=z, y=yz)
    t_ls, t_ll, t_noise, t_c = dev(ls, True), dev(ll, True), dev(noise, True), dev(const, True)
    params = {"log_observation_noise": t_noise, "mean": {"constant": t_c},
              "kernel": {"log_scaling": t_ls, "log_lengthscales": t_ll}}
    loss = S.NegativeLogLikelihood(gp=gp).calculate_empirical_risk(params, x, y)
    loss.backward()
    c_ls, c_ll, c_noise, c_c = cpu(ls, True), cpu(ll, True), cpu(noise, True), cpu(const, True)
    mean_t, var_t = T.exact_gp_posterior(c_ls, c_ll, c_c, cpu(z), cpu(yz), c_noise, cpu(x))
    ref = T.gaussian_nll(cpu(y), mean_t, var_t)
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
    for got, want in ((t_ls, c_ls), (t_noise, c_noise), (t_c, c_c)):
        assert abs(float(got.grad) - float(want.grad)) <= GTOL * max(abs(float(want.grad)), 1e-6), (got.grad, want.grad)
    assert rel_norm(host(t_ll.grad).reshape(-1), c_ll.grad.numpy().reshape(-1)) <= GTOL


# ---- SVGP family on the general path: log-SVGP (dense El) and dLoss/dSigma from the weighted-Gram kernel ----------
def svgp_problem(n, m, d, seed, sharp=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    z = x[rng.permutation(n)[:m]].copy()
    y = np.sin(x @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
    # sharp: short lengthscales -> K_ZZ close to the identity, so that the log-SVGP El = exp(P) exp(P)^T (whose
    # initialisation exponentiates log(inv(chol(.)))) stays within a few orders of magnitude
    ls, ll = 0.1, np.log(rng.uniform(0.5, 1.5, d) * (6.0 if sharp else 1.0 / d))
    return x, z, y, ls, ll


def test_log_svgp_kernel_matches_oracle(S):
    from gvi_gaussian_process_b200.src.kernels.approximate import LogSVGPKernel

    n, m, d, lam = 1500, 48, 3, 1e-5
    x, z, y, ls, ll = svgp_problem(n, m, d, seed=2, sharp=True)
    gram = lambda a, b: O.ard_gram(ls, ll, a, b)
    kern = LogSVGPKernel(regulariser_kernel=S.ARDKernel(number_of_dimensions=d),
                         regulariser_kernel_parameters={"log_scaling": ls, "log_lengthscales": ll},
                         log_observation_noise=np.log(0.5), inducing_points=z, training_points=x[:300],
                         diagonal_regularisation=lam)
    params = kern.generate_parameters()
    init = O.log_svgp_initial_parameters(gram, z, x[:300], np.log(0.5), lam)
    assert rel_norm(host(params.log_el_matrix), init) <= 1e-9
    rng = np.random.default_rng(0)
    log_el = init + 0.05 * rng.standard_normal(init.shape)
    params = kern.generate_parameters({"log_el_matrix": log_el})
    sigma = O.log_svgp_sigma(log_el)
    ref_full = O.svgp_gram(gram, z, sigma, x[:40], x[:25], lam)
    assert rel_norm(host(kern.calculate_gram(params, x[:40], x[:25])), ref_full) <= TOL
    ref_diag = np.array([O.svgp_gram(gram, z, sigma, x[i:i + 1], x[i:i + 1], lam)[0, 0] for i in range(0, n, 7)])
    got = host(kern.calculate_gram(params, x, x, full_covariance=False))[::7]
    assert rel_norm(got, ref_diag) <= TOL


@pytest.mark.parametrize("variant", ["cholesky", "diagonal", "log"])
def test_svgp_general_path_backward_matches_autograd_twin(S, variant):
    """loss = NLL + GWI regulariser evaluated term by term (not the fused launch): the SVGP variance is differentiated
    through SVGPVariance (dLoss/dSigma = weighted Gram on the tensor pipe)."""
    from gvi_gaussian_process_b200.src.kernels.approximate import (CholeskySVGPKernel, DiagonalSVGPKernel,
                                                                   LogSVGPKernel)

    n, m, d, lam = 900, 40, 2, 1e-4
    x, z, y, ls, ll = svgp_problem(n, m, d, seed=5, sharp=(variant == "log"))
    rng = np.random.default_rng(1)
    cls = {"cholesky": CholeskySVGPKernel, "diagonal": DiagonalSVGPKernel, "log": LogSVGPKernel}[variant]
    kern = cls(regulariser_kernel=S.ARDKernel(number_of_dimensions=d),
               regulariser_kernel_parameters={"log_scaling": ls, "log_lengthscales": ll},
               log_observation_noise=np.log(0.5), inducing_points=z, training_points=x[:200],
               diagonal_regularisation=lam)
    init = kern.generate_parameters()
    if variant == "cholesky":
        raw = {"el_matrix_lower_triangle": host(init.el_matrix_lower_triangle) + 0.01 * rng.standard_normal((m, m)),
               "el_matrix_log_diagonal": host(init.el_matrix_log_diagonal) + 0.05 * rng.standard_normal(m)}
    elif variant == "diagonal":
        raw = {"log_el_matrix_diagonal": host(init.log_el_matrix_diagonal) + 0.05 * rng.standard_normal(m)}
    else:
        raw = {"log_el_matrix": host(init.log_el_matrix) + 0.02 * rng.standard_normal((m, m))}
    leaves = {k: dev(v, True) for k, v in raw.items()}
    w = dev(rng.standard_normal(d) / np.sqrt(d), True)
    mean = S.CustomMean(mean_function=lambda p, xx: torch.tanh(xx @ p))
    gp = S.ApproximateGPRegression(mean=mean, kernel=kern)
    params = {"mean": {"custom": w}, "kernel": leaves}
    # standalone NLL (never fused): goes through PredictiveVariance -> PointwiseRisk
    loss = S.NegativeLogLikelihood(gp=gp).calculate_empirical_risk(params, x, y)
    loss.backward()
    # twin
    tl = {k: cpu(v, True) for k, v in raw.items()}
    tw = cpu(host(w), True)
    if variant == "cholesky":
        e = torch.tril(tl["el_matrix_lower_triangle"], diagonal=-1) + torch.diag(torch.exp(tl["el_matrix_log_diagonal"]))
        sigma = e.T @ e
    elif variant == "diagonal":
        sigma = torch.diag(torch.exp(tl["log_el_matrix_diagonal"]))
    else:
        ex = torch.exp(tl["log_el_matrix"])
        el = ex @ ex.T
        sigma = el.T @ el
    var = T.svgp_diag(cpu(ls), cpu(ll), cpu(z), cpu(x), sigma, lam, False)
    ref = T.gaussian_nll(cpu(y), torch.tanh(cpu(x) @ tw), var)
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
    assert rel_norm(host(w.grad), tw.grad.numpy()) <= GTOL
    for k in raw:
        got, want = host(leaves[k].grad), tl[k].grad.numpy()
        if k == "el_matrix_lower_triangle":
            want = np.tril(want, -1)
        assert rel_norm(got, want) <= GTOL, k


def test_classification_with_kernelised_svgp_kernels(S):
    """The reference's MNIST configuration (experiments/classification/mnist/main.py:219-246): one KernelisedSVGPKernel
    per class over that class's inducing points, trained base kernel on top of the frozen sparse-posterior part."""
    from gvi_gaussian_process_b200.src.kernels.approximate import KernelisedSVGPKernel

    n, d, k, m, lam = 700, 4, 3, 32, 1e-5
    pr = make_classification_problem(n, d, k, m, seed=21)
    kernels, kparams, leaves = [], [], []
    for i in range(k):
        reg_params = {"log_scaling": pr["ls_p"][i], "log_lengthscales": pr["ll_p"][i]}
        kernels.append(KernelisedSVGPKernel(base_kernel=S.ARDKernel(number_of_dimensions=d),
                                            regulariser_kernel=S.ARDKernel(number_of_dimensions=d),
                                            regulariser_kernel_parameters=reg_params,
                                            log_observation_noise=pr["noise_p"][i], inducing_points=pr["zs"][i],
                                            training_points=pr["x"][:100], diagonal_regularisation=lam))
        a = dev(pr["ls"][i], True)
        leaves.append(a)
        kparams.append({"base_kernel": {"log_scaling": a, "log_lengthscales": pr["ll"][i]}})
    w1, w2 = dev(pr["w1"], True), dev(pr["w2"], True)
    mean = S.CustomMean(mean_function=lambda p, xx: torch.tanh(xx @ p[0]) @ p[1], number_output_dimensions=k)
    gp = S.ApproximateGPClassification(mean=mean, kernel=S.MultiOutputKernel(kernels=kernels))
    params = {"mean": {"custom": (w1, w2)}, "kernel": {"kernels": kparams}}
    loss = S.NegativeLogLikelihood(gp=gp).calculate_empirical_risk(params, pr["x"], pr["y"])
    loss.backward()
    # twin: var_i = frozen sparse posterior of the regulariser kernel + exp(ls_base)^2
    tx = cpu(pr["x"])
    tw1, tw2 = cpu(pr["w1"], True), cpu(pr["w2"], True)
    tls = [cpu(pr["ls"][i], True) for i in range(k)]
    m_q = (torch.tanh(tx @ tw1) @ tw2).reshape(k, -1)
    v_q = torch.stack([T.sparse_posterior_diag(cpu(pr["ls_p"][i]), cpu(pr["ll_p"][i]), cpu(pr["zs"][i]), tx, lam, False)
                       + torch.exp(tls[i]) ** 2 for i in range(k)])
    ref = T.multinomial_nll(T.classification_probabilities(m_q, v_q, HERMITE[0], HERMITE[1]), cpu(pr["y"]))
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
    assert rel_norm(host(w1.grad), tw1.grad.numpy()) <= GTOL
    for i in range(k):
        assert abs(float(leaves[i].grad) - float(tls[i].grad)) <= GTOL * abs(float(tls[i].grad))


# ---- section 8(f)-4: device-resident optimiser step and the Trainer loop ------------------------------------------
@pytest.mark.parametrize("kind,code", [("adam", 0), ("adabelief", 1), ("rmsprop", 2)])
def test_optimiser_step_kernel_matches_oracle(S, kind, code):
    from gvi_gaussian_process_b200 import _lib
    from gvi_gaussian_process_b200.experiments.shared.trainer import _OPTIMISERS, OptimiserSchema

    lib = _lib.load()
    rng = np.random.default_rng(code)
    n = 10007
    p = rng.standard_normal(n)
    mu, nu = np.zeros(n), np.zeros(n)
    dp, dmu, dnu = dev(p), dev(mu), dev(nu)
    k, b1, b2, eps, eps_root = _OPTIMISERS[OptimiserSchema(kind)]
    assert k == code
    for count in range(1, 6):
        g = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 2)
        p, mu, nu = O.optimiser_step(kind, p, g, mu, nu, count, 1e-2)
        rc = lib.gvi_optimiser_step_f64(code, dp.data_ptr(), dev(g).data_ptr(), dmu.data_ptr(), dnu.data_ptr(), n,
                                        count, 1e-2, b1, b2, eps, eps_root, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    assert rel_norm(host(dp), p) <= 1e-13
    assert rel_norm(host(dnu), nu) <= 1e-13


def test_trainer_exact_gp_nll_matches_cpu_twin(S):
    """train_regulariser_gp of experiments/regression/trainers.py:19-85 in miniature: the exact GP on the inducing data,
    NLL loss, AdaBelief, shuffled mini-batches in jax.random.permutation order - against a CPU twin of the same loop
    (torch-f64 autograd gradients, NumPy optimiser, oracle PRNG)."""
    from gvi_gaussian_process_b200.experiments.shared import Data, OptimiserSchema, Trainer, TrainerSettings
    from oracle import jax_prng as OP

    rng = np.random.default_rng(8)
    m, d = 96, 2
    z = rng.standard_normal((m, d))
    yz = np.sin(z @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(m)
    ls0, ll0, noise0, c0 = 0.1, np.log(np.full(d, 0.7)), np.log(0.5), 0.0
    kernel = S.ARDKernel(number_of_dimensions=d)
    gp = S.GPRegression(mean=S.ConstantMean(), kernel=kernel, x=z, y=yz)
    params = gp.generate_parameters({"log_observation_noise": noise0, "mean": {"constant": c0},
                                     "kernel": {"log_scaling": ls0, "log_lengthscales": ll0}})
    nll = S.NegativeLogLikelihood(gp=gp)
    settings = TrainerSettings(seed=3, optimiser_schema=OptimiserSchema.adabelief, learning_rate=5e-2,
                               number_of_epochs=3, batch_size=40, batch_shuffle=True, batch_drop_last=False)
    trainer = Trainer(save_checkpoint_frequency=0, checkpoint_path="",
                      post_epoch_callback=lambda p: {"empirical-risk": float(nll.calculate_empirical_risk(p, z, yz))})
    trained, history = trainer.train(settings, params, Data(x=z, y=yz),
                                     lambda pd, x, y: nll.calculate_empirical_risk(pd, x, y))
    assert len(history) == 3 and history[-1]["empirical-risk"] < history[0]["empirical-risk"] + 1e-9
    # ---- CPU twin of the same loop: leaf order of parameters.dict() = noise, mean constant, log_scaling, lengthscales
    p = np.concatenate([[noise0], [c0], [ls0], ll0])
    mu, nu = np.zeros_like(p), np.zeros_like(p)
    key = OP.prng_key(3)
    count = 0
    for _ in range(3):
        key, sub = OP.split(key)
        perm = OP.permutation(sub, m)
        for b in range(int(np.ceil(m / 40))):
            idx = perm[b * 40:(b + 1) * 40]
            t = [cpu(p[0], True), cpu(p[1], True), cpu(p[2], True), cpu(p[3:], True)]
            mean_t, var_t = T.exact_gp_posterior(t[2], t[3], t[1], cpu(z), cpu(yz), t[0], cpu(z[idx]))
            T.gaussian_nll(cpu(yz[idx]), mean_t, var_t).backward()
            g = np.concatenate([np.atleast_1d(a.grad.numpy()) for a in t])
            count += 1
            p, mu, nu = O.optimiser_step("adabelief", p, g, mu, nu, count, 5e-2)
    got = np.concatenate([host(v).reshape(-1) for v in (trained.log_observation_noise, trained.mean["constant"],
                                                         trained.kernel["log_scaling"],
                                                         trained.kernel["log_lengthscales"])])
    assert rel_norm(got, p) <= 1e-7, (got, p)


def test_golden_fixture_classification(S):
    """tests/golden/config4_mini_classification.npz (oracle outputs, script committed next to it): the product path
    reproduces probabilities, risks and regularisers of the committed fixture."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config4_mini_classification.npz"))
    x, y, k, d = g["x"], g["y"], g["y"].shape[1], g["x"].shape[1]
    lam = float(g["lam"])
    kern_q = S.MultiOutputKernel(kernels=[
        S.SparsePosteriorKernel(base_kernel=S.ARDKernel(number_of_dimensions=d), inducing_points=g["zs"][i],
                                diagonal_regularisation=lam) for i in range(k)])
    mean_q = S.CustomMean(mean_function=lambda p, xx: xx @ p, number_output_dimensions=k)
    approx = S.ApproximateGPClassification(mean=mean_q, kernel=kern_q)
    params = {"mean": {"custom": dev(g["w"])},
              "kernel": {"kernels": [{"base_kernel": {"log_scaling": g["ls_q"][i], "log_lengthscales": g["ll_q"][i]}}
                                     for i in range(k)]}}
    exact = S.GPClassification(mean=S.ConstantMean(number_output_dimensions=k),
                               kernel=S.MultiOutputKernel(kernels=[S.ARDKernel(number_of_dimensions=d)
                                                                   for _ in range(k)]),
                               x=x[g["zi"]], y=y[g["zi"]])
    eparams = {"log_observation_noise": g["noise_p"], "mean": {"constant": np.zeros(k)},
               "kernel": {"kernels": [{"log_scaling": g["ls_p"][i], "log_lengthscales": g["ll_p"][i]}
                                      for i in range(k)]}}
    assert np.max(np.abs(host(approx.predict_probability(params, x).probabilities) - g["prob_q"])) <= TOL
    assert np.max(np.abs(host(exact.predict_probability(eparams, x).probabilities) - g["prob_p"])) <= TOL
    close = lambda a, b: abs(float(a) - float(b)) <= TOL * max(abs(float(b)), 1e-3)
    assert close(S.NegativeLogLikelihood(gp=approx).calculate_empirical_risk(params, x, y), g["nll"])
    assert close(S.CrossEntropy(gp=approx).calculate_empirical_risk(params, x, y), g["cross_entropy"])
    kw = dict(gp=approx, regulariser=exact, regulariser_parameters=eparams, mode="posterior")
    assert close(S.MultinomialWassersteinRegularisation(**kw).calculate_regularisation(params, x),
                 g["reg_multinomial_wasserstein"])
    assert close(S.GaussianWassersteinRegularisation(**kw).calculate_regularisation(params, x), g["reg_gwi_no_eig"])
    for kind, cls in [("bhattacharyya", "ProjectedBhattacharyyaRegularisation"), ("kl", "ProjectedKLRegularisation"),
                      ("gaussian_wasserstein", "ProjectedGaussianWassersteinRegularisation"),
                      ("hellinger", "ProjectedHellingerRegularisation"), ("renyi", "ProjectedRenyiRegularisation")]:
        assert close(getattr(S.projected, cls)(**kw).calculate_regularisation(params, x), g["reg_" + kind]), kind


def test_classification_backward_wide_inputs(S):
    """D > 32 (config #4's regime): the per-class variance VJP takes the GEMM formulation (grad_wide.cu)."""
    n, d, k, m, lam = 500, 48, 3, 20, 1e-3
    pr = make_classification_problem(n, d, k, m, seed=31)
    approx, params, exact, eparams, leaves, (w1, w2) = build_classification_models(S, pr, lam, grad=True)
    nll = S.NegativeLogLikelihood(gp=approx)
    reg = S.GaussianWassersteinRegularisation(gp=approx, regulariser=exact, regulariser_parameters=eparams,
                                              mode="posterior")
    loss = S.GeneralisedVariationalInference(regularisation=reg, empirical_risk=nll).calculate_loss(
        params, pr["x"], pr["y"])
    loss.backward()
    _, _, m_p, v_p = oracle_classification(pr, lam)
    tx, ty = cpu(pr["x"]), cpu(pr["y"])
    tw1, tw2 = cpu(pr["w1"], True), cpu(pr["w2"], True)
    tls = [cpu(pr["ls"][i], True) for i in range(k)]
    tll = [cpu(pr["ll"][i], True) for i in range(k)]
    m_q = (torch.tanh(tx @ tw1) @ tw2).reshape(k, -1)
    v_q = torch.stack([T.sparse_posterior_diag(tls[i], tll[i], cpu(pr["zs"][i]), tx, lam, False) for i in range(k)])
    ref = (T.multinomial_nll(T.classification_probabilities(m_q, v_q, HERMITE[0], HERMITE[1]), ty)
           + T.regulariser("gwi_no_eig", cpu(m_p), cpu(v_p), m_q, v_q))
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
    assert rel_norm(host(w2.grad), tw2.grad.numpy()) <= GTOL
    for i in range(k):
        assert abs(float(leaves[i][0].grad) - float(tls[i].grad)) <= GTOL * max(abs(float(tls[i].grad)), 1e-3), i
        assert rel_norm(host(leaves[i][1].grad), tll[i].grad.numpy()) <= GTOL, i


def test_exact_gp_classifier_nll_backward_matches_autograd_twin(S):
    """train_regulariser_gp for classification (experiments/classification/trainers.py:20-100): multinomial NLL of the
    exact GP classifier on the inducing data, differentiated w.r.t. every output's kernel hyper-parameters, its
    observation noise and the constant means."""
    rng = np.random.default_rng(17)
    m, d, k = 61, 3, 3  # m not a multiple of k: the tiled-and-reshaped constant mean differs between outputs
    z = rng.standard_normal((m, d))
    yz = np.eye(k)[rng.integers(0, k, m)]
    ls = rng.uniform(-0.2, 0.2, k)
    ll = [np.log(rng.uniform(0.5, 1.5, d) / d) for _ in range(k)]
    noise = np.log(rng.uniform(0.3, 1.0, k))
    const = rng.uniform(-0.1, 0.1, k)
    gp = S.GPClassification(mean=S.ConstantMean(number_output_dimensions=k),
                            kernel=S.MultiOutputKernel(kernels=[S.ARDKernel(number_of_dimensions=d) for _ in range(k)]),
                            x=z, y=yz)
    t_noise, t_c = dev(noise, True), dev(const, True)
    leaves = [(dev(ls[i], True), dev(ll[i], True)) for i in range(k)]
    params = {"log_observation_noise": t_noise, "mean": {"constant": t_c},
              "kernel": {"kernels": [{"log_scaling": a, "log_lengthscales": b} for a, b in leaves]}}
    loss = S.NegativeLogLikelihood(gp=gp).calculate_empirical_risk(params, z, yz)
    loss.backward()
    c_noise, c_c = cpu(noise, True), cpu(const, True)
    cl = [(cpu(ls[i], True), cpu(ll[i], True)) for i in range(k)]
    # the reference's K-output ConstantMean: jnp.tile(constant, n) RESHAPED to [K, n] (constant_mean.py:66,
    # means/base.py:97-100) - every row cycles through all K constants; reproduced, not "fixed"
    prior = c_c.repeat(m).reshape(k, m)
    means, variances = [], []
    for i in range(k):
        mi, vi = T.exact_gp_posterior(cl[i][0], cl[i][1], prior[i], cpu(z), cpu(yz[:, i]), c_noise[i], cpu(z))
        means.append(mi)
        variances.append(vi)
    q = T.classification_probabilities(torch.stack(means), torch.stack(variances), HERMITE[0], HERMITE[1])
    ref = T.multinomial_nll(q, cpu(yz))
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
    assert rel_norm(host(t_noise.grad), c_noise.grad.numpy()) <= GTOL
    assert rel_norm(host(t_c.grad), c_c.grad.numpy()) <= GTOL
    for i in range(k):
        assert abs(float(leaves[i][0].grad) - float(cl[i][0].grad)) <= GTOL * max(abs(float(cl[i][0].grad)), 1e-3)
        assert rel_norm(host(leaves[i][1].grad), cl[i][1].grad.numpy()) <= GTOL
