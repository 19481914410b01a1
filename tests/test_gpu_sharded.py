"""GPU tests of the multi-rank paths that run on ONE device:
  - the selector's step API (gvi_select_init / resolve / update) driven per slice for R emulated ranks, the gather done in
    torch, and
  - the sharded selector proper (gvi_cond_var_select_sharded_f64's kernels and mailbox exchange) over R virtual shards,
both bit-identical to the single-shard call and to the oracle, including ties that straddle a shard boundary and empty
shards; the device permutation against the oracle's jax.random restatement; the communicator at world size 1.
With two or more GPUs visible the torchrun script tests/run_multigpu.py runs as well."""
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O
from oracle import jax_prng as OP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import types

    from gvi_gaussian_process_b200 import _lib, distributed
    from gvi_gaussian_process_b200.src import _lib_proxy as L
    from gvi_gaussian_process_b200.src.inducing_points_selection import ConditionalVarianceInducingPointsSelector
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.utils import jax_prng

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.load()
    return types.SimpleNamespace(**locals())


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


def _bounds(n, r, kind):
    if kind == "even":
        return [n * k // r for k in range(r + 1)]
    if kind == "empty_first_and_middle":  # shards 0 and r // 2 own no point
        live = [k for k in range(r) if k not in (0, r // 2)]
        cuts, filled = [0], 0
        for k in range(r):
            if k in live:
                filled += 1
                cuts.append(n * filled // len(live))
            else:
                cuts.append(cuts[-1])
        return cuts
    raise ValueError(kind)


def _problem(seed, n, d, ties):
    rng = np.random.default_rng(seed)
    if ties:
        # every point appears 3 times, the copies far apart: each pivot is an exact 3-way tie decided by the
        # highest-index rule, and the copies of one point fall into DIFFERENT shards
        base = rng.standard_normal((n // 3, d))
        x = np.concatenate([base, base, base])
    else:
        x = rng.standard_normal((n, d))
    ls, ll = 0.2, rng.normal(1.0 if ties else -0.5, 0.3, d)  # short lengthscales keep the tied pivots well-posed
    return x, ls, ll


def _oracle_pivots(x, m, ls, ll):
    return O.conditional_variance_select(x, m, lambda a, b: O.ard_gram(ls, ll, a, b),
                                         lambda a: O.ard_gram_diag(ls, ll, a), use_argsort=(x.shape[0] <= 5000))


def _step_api_emulated(S, x, m_sel, ls, ll, bounds, jitter=1e-12):
    """R ranks emulated in one process: init / update per slice, the all-gather is a torch.stack, resolve on the device."""
    lib, _lib = S._lib.load(), S._lib
    n, d = x.shape
    r = len(bounds) - 1
    rl = lib.gvi_select_record_len(d, m_sel)
    lsd, lld = dev([ls]), dev(ll)
    st = _lib.current_stream
    slices, wss, recs = [], [], []
    for k in range(r):
        lo, hi = bounds[k], bounds[k + 1]
        assert hi > lo, "the step API takes non-empty slices"
        slices.append(dev(x[lo:hi]))
        wss.append(torch.empty(lib.gvi_select_workspace_bytes(hi - lo, m_sel, d), dtype=torch.uint8, device="cuda"))
        recs.append(torch.zeros(rl, dtype=torch.float64, device="cuda"))
        _lib.check(lib.gvi_select_init_f64(slices[k].data_ptr(), hi - lo, lo, d, m_sel, lsd.data_ptr(), jitter,
                                           recs[k].data_ptr(), wss[k].data_ptr(), st()), "init")
    winner = torch.zeros(rl, dtype=torch.float64, device="cuda")
    idx = torch.zeros(m_sel, dtype=torch.int64, device="cuda")
    tr = torch.zeros((r, m_sel - 1), dtype=torch.float64, device="cuda")
    for m in range(m_sel):
        gathered = torch.stack(recs).contiguous()
        _lib.check(lib.gvi_select_resolve_f64(gathered.data_ptr(), r, d, m_sel, int(m == 0), winner.data_ptr(),
                                              idx[m:].data_ptr(), st()), "resolve")
        if m == m_sel - 1:
            break
        for k in range(r):
            lo, hi = bounds[k], bounds[k + 1]
            _lib.check(lib.gvi_select_update_f64(slices[k].data_ptr(), hi - lo, lo, d, m_sel, m, winner.data_ptr(),
                                                 lsd.data_ptr(), lld.data_ptr(), lld.numel(), jitter,
                                                 recs[k].data_ptr(), tr[k, m:].data_ptr(), wss[k].data_ptr(), st()),
                       "update")
    return idx, tr.sum(0)


@pytest.mark.parametrize("r", [2, 3, 8])
@pytest.mark.parametrize("ties", [False, True])
def test_selector_step_api_emulated_ranks_bit_exact(S, r, ties):
    n, d, m = (3 * 1700, 3, 60) if not ties else (3 * 400, 2, 48)
    x, ls, ll = _problem(10 + r, n, d, ties)
    ref_idx, ref_tr = S.L.cond_var_select(dev(x), m, dev([ls]), dev(ll), 1e-12)
    idx, tr = _step_api_emulated(S, x, m, ls, ll, _bounds(n, r, "even"))
    assert torch.equal(idx, ref_idx)
    assert np.array_equal(host(idx), _oracle_pivots(x, m, ls, ll))
    assert float((tr - ref_tr).abs().max()) <= 1e-12 * float(ref_tr.abs().max())
    if ties:  # the three copies of a point sit in different shards: the tie was resolved ACROSS shards
        third = n // 3
        assert len(set((host(idx) % third).tolist())) == m and np.all(host(idx)[1:] >= 2 * third)


@pytest.mark.parametrize("r,kind", [(1, "even"), (2, "even"), (3, "even"), (8, "even"), (4, "empty_first_and_middle"),
                                    (8, "empty_first_and_middle")])
@pytest.mark.parametrize("ties", [False, True])
def test_sharded_selector_kernels_on_virtual_shards_bit_exact(S, r, kind, ties):
    """The kernels and the mailbox exchange of gvi_cond_var_select_sharded_f64, R shards in one launch."""
    n, d, m = (3 * 2300, 4, 70) if not ties else (3 * 500, 2, 64)
    x, ls, ll = _problem(20 + r, n, d, ties)
    bounds = _bounds(n, r, kind)
    assert bounds[0] == 0 and bounds[-1] == n and len(bounds) == r + 1
    ref_idx, ref_tr = S.L.cond_var_select(dev(x), m, dev([ls]), dev(ll), 1e-12)
    idx, tr = S.L.cond_var_select_virtual_shards(dev(x), m, dev([ls]), dev(ll), 1e-12, bounds)
    assert idx.shape == (r, m)
    for k in range(r):  # every shard (the empty ones too) saw the same winners
        assert torch.equal(idx[k], ref_idx), (k, host(idx[k])[:8], host(ref_idx)[:8])
        assert float((tr[k] - ref_tr).abs().max()) <= 1e-12 * float(ref_tr.abs().max())
        assert torch.equal(tr[k], tr[0])  # summed in rank order: the same bits everywhere
    assert np.array_equal(host(idx[0]), _oracle_pivots(x, m, ls, ll))


def test_sharded_selector_large_and_scalar_lengthscale(S):
    """N = 3e5 over 8 virtual shards, M = 128 (several waves of CTAs per shard), scalar lengthscale broadcast over D."""
    rng = np.random.default_rng(5)
    n, d, m = 300_007, 8, 128
    x = rng.standard_normal((n, d))
    xd, lsd, lld = dev(x), dev([0.0]), dev([np.log(1.0 / d)])
    ref_idx, _ = S.L.cond_var_select(xd, m, lsd, lld, 1e-12)
    idx, _ = S.L.cond_var_select_virtual_shards(xd, m, lsd, lld, 1e-12, _bounds(n, 8, "even"))
    assert torch.equal(idx[0], ref_idx) and torch.equal(idx[7], ref_idx)


def test_communicator_world_size_one(S):
    """gvi_comm at world size 1 (no NCCL involved): the sharded entry point equals the single-shard one, the all-reduce is
    the identity; the selector shim with comm= takes the same path as without."""
    comm = S.distributed.Comm()
    assert comm.world == 1 and comm.rank == 0 and comm.has_peer_window
    rng = np.random.default_rng(6)
    n, d, m = 4000, 3, 50
    x = rng.standard_normal((n, d))
    lsd, lld = dev([0.1]), dev(rng.normal(-0.5, 0.2, d))
    ref_idx, ref_tr = S.L.cond_var_select(dev(x), m, lsd, lld, 1e-12)
    for _ in range(2):  # twice: tags keep growing over the communicator's lifetime
        idx, tr = S.distributed.ShardedSelector(comm).select(dev(x), 0, m, lsd, lld)
        assert torch.equal(idx, ref_idx)
        assert float((tr - ref_tr).abs().max()) <= 1e-12 * float(ref_tr.abs().max())
    v = dev([1.5, -2.0, 3.0])
    assert torch.equal(comm.allreduce_sum(v.clone()), v)
    assert torch.equal(comm.allgather(v), v[None])
    comm.destroy()


@pytest.mark.parametrize("n", [1, 2, 5, 1625, 1626, 100_000, 1_000_000, 3_000_000])
def test_device_permutation_is_bit_exact(S, n):
    """jax.random.permutation restated: threefry hash on the device + stable key sort, against the NumPy oracle
    (1, 2 and 3 shuffle rounds: n <= 1625, <= ~2.6e6, above)."""
    _, sub = OP.split(OP.prng_key(7))
    got = host(S.jax_prng.permutation_device(np.asarray(sub), n, torch.device("cuda")))
    assert np.array_equal(got, OP.permutation(sub, n))


def test_threefry_random_bits_match_the_oracle(S):
    for n in (1, 2, 3, 1000, 1001):
        key = OP.split(OP.prng_key(n))[1]
        got = host(S.L.threefry_random_bits(key, n, torch.device("cuda")))
        assert np.array_equal(got, OP.random_bits32(key, n).astype(np.int64))


def test_multi_gpu_when_two_gpus_are_visible(S):
    """tests/run_multigpu.py under torchrun on 2 GPUs (skipped on a 1-GPU box)."""
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "run_multigpu.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "[multigpu] OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_preprocess_function_is_applied_like_the_reference():
    """the reference applies `preprocess_function` inside every calculate_gram call
    (src/kernels/base.py:45-63, sparse_posterior_kernel.py:63-90).  A kernel that carries one must not take the raw-input
    ARD fast paths: predictive, exact posterior and the (then unfused) GVI loss equal the oracle on the preprocessed
    inputs - for a preprocess_function on the ARD base kernel and for one on the sparse wrapper kernel."""
    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
    from gvi_gaussian_process_b200.src.gps import ApproximateGPRegression, GPRegression
    from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean
    from gvi_gaussian_process_b200.src.regularisations.projected import ProjectedKLRegularisation

    rng = np.random.default_rng(3)
    n, m, d = 600, 40, 3
    x, z = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    y, yz = rng.standard_normal(n), rng.standard_normal(m)
    ls, ll = 0.1, rng.normal(-0.3, 0.1, d)
    f = lambda a: 2.0 * a + 0.5
    gram = lambda a, b: O.ard_gram(ls, ll, a, b)
    diag = lambda a: O.ard_gram_diag(ls, ll, a)

    def reference_loss(fx, fz):
        v_q = O.sparse_posterior_diag(gram, diag, fz, fx, 1e-5, False)
        m_p, v_p = O.exact_gp_posterior(gram, diag, np.zeros(n), fz, yz, 0.0, fx)
        return v_q, O.gaussian_nll(y, np.zeros(n), v_q) + O.projected_regularisation("kl", m_p, v_p, np.zeros(n), v_q, 0.5)

    for where in ("base", "wrapper"):
        base = ARDKernel(number_of_dimensions=d, preprocess_function=f if where == "base" else None)
        kp = base.Parameters.construct(log_scaling=ls, log_lengthscales=ll)
        sparse = SparsePosteriorKernel(base_kernel=base, inducing_points=z,
                                       preprocess_function=f if where == "wrapper" else None)
        mean = ConstantMean()
        approx = ApproximateGPRegression(mean=mean, kernel=sparse)
        ap = approx.Parameters.construct(mean=mean.Parameters.construct(constant=0.0),
                                         kernel=sparse.Parameters.construct(base_kernel=kp))
        exact = GPRegression(mean=mean, kernel=base, x=z, y=yz)
        ep = exact.Parameters.construct(log_observation_noise=0.0, mean=mean.Parameters.construct(constant=0.0), kernel=kp)
        gvi = GeneralisedVariationalInference(
            regularisation=ProjectedKLRegularisation(gp=approx, regulariser=exact, regulariser_parameters=ep,
                                                     mode="posterior"),
            empirical_risk=NegativeLogLikelihood(gp=approx))
        assert not gvi._fusable()
        if where == "base":   # the base kernel preprocesses everything it is handed: x and Z (and the exact GP's x)
            v_ref, loss_ref = reference_loss(f(x), f(z))
        else:                 # the wrapper preprocesses x only; Z and the exact GP see raw inputs
            v_q = O.sparse_posterior_diag(gram, diag, z, f(x), 1e-5, False)
            m_p, v_p = O.exact_gp_posterior(gram, diag, np.zeros(n), z, yz, 0.0, x)
            v_ref = v_q
            loss_ref = O.gaussian_nll(y, np.zeros(n), v_q) + O.projected_regularisation("kl", m_p, v_p, np.zeros(n), v_q, 0.5)
        var = host(approx.predict_probability(ap, dev(x)).covariance)
        assert np.max(np.abs(var - v_ref)) <= 1e-10 * np.max(np.abs(v_ref)), where
        loss = float(gvi.calculate_loss(ap, dev(x), dev(y)))
        assert abs(loss - loss_ref) <= 1e-10 * abs(loss_ref), (where, loss, loss_ref)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (37, 129, 70), (256, 128, 64), (300, 300, 513), (1000, 257, 19)])
def test_in_tree_dgemm_matches_a_float64_matmul(S, ta, tb, m, n, k):
    """csrc/dgemm.cu (the tensor-pipe GEMM that replaced the library GEMMs of the exact-GP / wide-input / SVGP gradient
    routes) against torch.matmul in float64: all four operand orientations, ragged sizes, beta accumulation."""
    lib = S._lib.load_test()
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    a = dev(rng.standard_normal((k, m) if ta else (m, k)))
    b = dev(rng.standard_normal((n, k) if tb else (k, n)))
    c0 = dev(rng.standard_normal((m, n)))
    ref = (a.T if ta else a) @ (b.T if tb else b)
    for beta in (0.0, 1.0, -0.5):
        c = c0.clone()
        S._lib.check(lib.gvi_test_dgemm_f64(ta, tb, m, n, k, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), beta,
                                            c.data_ptr(), c.stride(0), 0, 0, S._lib.current_stream()), "dgemm")
        want = ref + beta * c0
        assert float((c - want).abs().max()) <= 1e-12 * max(1.0, float(want.abs().max())), (beta, ta, tb)


def test_in_tree_dgemm_triangular_and_symmetric_modes(S):
    """The two structure short-cuts: a lower-triangular operand (k range cut at the tile's columns) and a symmetric result
    (only tiles on or below the diagonal multiplied, the rest mirrored) give the full-GEMM answer."""
    lib = S._lib.load_test()
    rng = np.random.default_rng(0)
    nr, mp = 500, 392
    kmat = dev(rng.standard_normal((nr, mp)))
    linv = torch.tril(dev(rng.standard_normal((mp, mp))))
    t = torch.empty(nr, mp, dtype=torch.float64, device="cuda")
    S._lib.check(lib.gvi_test_dgemm_f64(0, 1, nr, mp, mp, kmat.data_ptr(), mp, linv.data_ptr(), mp, 0.0, t.data_ptr(), mp,
                                        1, 0, S._lib.current_stream()), "K L^-T")
    assert float((t - kmat @ linv.T).abs().max()) <= 1e-12 * float((kmat @ linv.T).abs().max())
    aa = torch.empty_like(t)
    S._lib.check(lib.gvi_test_dgemm_f64(0, 0, nr, mp, mp, t.data_ptr(), mp, linv.data_ptr(), mp, 0.0, aa.data_ptr(), mp,
                                        2, 0, S._lib.current_stream()), "T L^-1")
    assert float((aa - t @ linv).abs().max()) <= 1e-12 * float((t @ linv).abs().max())
    g = dev(rng.standard_normal(nr))
    bmat = (aa * g[:, None]).contiguous()
    s0 = dev(rng.standard_normal((mp, mp)))
    s0 = (s0 + s0.T).contiguous()
    sm = s0.clone()
    S._lib.check(lib.gvi_test_dgemm_f64(1, 0, mp, mp, nr, aa.data_ptr(), mp, bmat.data_ptr(), mp, 1.0, sm.data_ptr(), mp,
                                        0, 1, S._lib.current_stream()), "Aa^T B")
    want = s0 + aa.T @ bmat
    assert float((sm - want).abs().max()) <= 1e-12 * float(want.abs().max())
    # tiles below the diagonal are mirrored bit for bit; inside a diagonal tile (i, j) and (j, i) are two dot products
    # that round differently, as in any GEMM
    assert torch.equal(sm[128:256, :128], sm[:128, 128:256].T)
    x = dev(rng.standard_normal(nr))
    y0 = dev(rng.standard_normal(mp))
    y = y0.clone()
    S._lib.check(lib.gvi_test_dgemv_f64(1, mp, nr, aa.data_ptr(), mp, x.data_ptr(), 1.0, y.data_ptr(),
                                        S._lib.current_stream()), "gemv T")
    assert float((y - (y0 + aa.T @ x)).abs().max()) <= 1e-12 * float((aa.T @ x).abs().max())
    y2 = torch.empty(nr, dtype=torch.float64, device="cuda")
    S._lib.check(lib.gvi_test_dgemv_f64(0, nr, mp, aa.data_ptr(), mp, y0.data_ptr(), 0.0, y2.data_ptr(),
                                        S._lib.current_stream()), "gemv N")
    assert float((y2 - aa @ y0).abs().max()) <= 1e-12 * float((aa @ y0).abs().max())


def test_training_gradients_beyond_3456_inducing_points():
    """r01 limit lifted: with M > 3456 the b tile of the gradient tile kernel no longer fits shared memory; the objective
    is then differentiated term by term and the variance VJP takes the GEMM route (csrc/grad_wide.cu on csrc/dgemm.cu),
    which has no size limit - BASELINE configs[4] reaches M = 4096.  loss.backward() against the torch-f64 twin (1e-8)."""
    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
    from gvi_gaussian_process_b200.src.gps import ApproximateGPRegression, GPRegression
    from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean
    from gvi_gaussian_process_b200.src.regularisations.projected import ProjectedKLRegularisation
    from oracle import torch_twin as TT

    rng = np.random.default_rng(12)
    n, m, d = 300, 3600, 4
    x, z = rng.standard_normal((n, d)), 2.0 * rng.standard_normal((m, d))
    y, yz = rng.standard_normal(n), rng.standard_normal(m)
    ls0, ll0 = 0.1, rng.normal(1.0, 0.1, d)  # short lengthscales: a well-conditioned 3600 x 3600 Gram
    kernel = ARDKernel(number_of_dimensions=d)
    ls = torch.tensor(ls0, dtype=torch.float64, device="cuda", requires_grad=True)
    ll = torch.tensor(ll0, dtype=torch.float64, device="cuda", requires_grad=True)
    mq = (0.3 * torch.randn(n, dtype=torch.float64, device="cuda")).requires_grad_(True)
    amean = CustomMean(mean_function=lambda p, xx: p[: xx.shape[0]])
    sparse = SparsePosteriorKernel(base_kernel=kernel, inducing_points=z, diagonal_regularisation=1e-5)
    approx = ApproximateGPRegression(mean=amean, kernel=sparse)
    ap = approx.Parameters.construct(mean=amean.Parameters.construct(custom=mq),
                                     kernel=sparse.Parameters.construct(
                                         base_kernel=kernel.Parameters.construct(log_scaling=ls, log_lengthscales=ll)))
    mean0 = ConstantMean()
    exact = GPRegression(mean=mean0, kernel=kernel, x=z, y=yz)
    ep = exact.Parameters.construct(log_observation_noise=0.0, mean=mean0.Parameters.construct(constant=0.0),
                                    kernel=kernel.Parameters.construct(log_scaling=-0.2, log_lengthscales=ll0 + 0.05))
    gvi = GeneralisedVariationalInference(
        regularisation=ProjectedKLRegularisation(gp=approx, regulariser=exact, regulariser_parameters=ep, mode="posterior"),
        empirical_risk=NegativeLogLikelihood(gp=approx))
    assert gvi._fusable()
    loss = gvi.calculate_loss(ap, dev(x), dev(y))
    loss.backward()
    gp_ = lambda a, b: O.ard_gram(-0.2, ll0 + 0.05, a, b)
    dp_ = lambda a: O.ard_gram_diag(-0.2, ll0 + 0.05, a)
    m_p, v_p = O.exact_gp_posterior(gp_, dp_, np.zeros(n), z, yz, 0.0, x)
    ref_loss, g_ls, g_ll, g_m = TT.gvi_loss_and_grads(x, y, z, host(mq), ls0, ll0, m_p, v_p, "kl", 0.5, 1e-5, False)
    assert abs(float(loss.detach()) - ref_loss) <= 1e-10 * abs(ref_loss)
    assert abs(float(ls.grad) - g_ls) <= 1e-8 * max(abs(g_ls), 1e-3)
    assert np.max(np.abs(host(ll.grad) - g_ll)) <= 1e-8 * max(np.max(np.abs(g_ll)), 1e-3)
    assert np.max(np.abs(host(mq.grad) - g_m)) <= 1e-8 * max(np.max(np.abs(g_m)), 1e-3)
