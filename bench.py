#!/usr/bin/env python
"""bench.py - BASELINE.json's metric on its config #3 (synthetic UCI-shaped regression: N=1e6 points per GPU,
D=8, M=1024 inducing points, ARD kernel, Gaussian NLL + projected Renyi(0.5) regulariser against the exact GP
posterior): fp64 points/s through the fused "Gram + predict + pGVI risk" step.

One step = everything the reference recomputes per objective evaluation (src/generalised_variational_inference.py
:73-94 under jit): the tanh-FCN mean on the host framework, the q and p factorisations (K_ZZ, Cholesky, L^-1),
and one fused launch over all points of the shard (both K(x,Z) tiles, both triangular products on the FP64 tensor
pipe, NLL + D0 reduction); N>1 adds one 3-double all-reduce.  Weak scaling: every rank owns 1e6 points.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun ... bench.py --gpus N ...           (one rank per GPU, NCCL)

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference (the NumPy/SciPy
oracle - the JAX reference itself is not installable here) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core (set before the BLAS libraries load)
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1_000_000
D = 8
M = 1024
LAMBDA = 1e-5
RENYI_ALPHA = 0.5
CPU_SAMPLE_ROWS = 50_000  # the reference's own batch size (experiments/regression/base_configs/train_approximate/base.yaml)
METRIC = "fp64 Gram+predict+pGVI risk points/sec"
UNIT = "points/s"
# measured on this pool's B200 (profiles/fp64_peak_r01.txt): DMMA m8n8k4 issue rate 37.1 TFLOP/s, cuBLAS DGEMM
# 8192^3 35.4 TFLOP/s.  MEASURED_PEAKS.json carries no FP64 entry, so the harder (DMMA) number is the denominator.
FP64_PEAK_TFLOPS = 37.1


def algorithmic_flops_per_point(m: int, d: int) -> float:
    """per point, q and p together: 2*[M^2 (triangular product) + 2*M*D (Gram)] + 2*M."""
    return 2.0 * (m * m + 2.0 * m * d) + 2.0 * m


def make_inputs(rank: int, n: int, seed: int):
    """Seeded synthetic shard: X ~ N(0, I_8), y = sin(X w) + 0.1 eps."""
    rng = np.random.default_rng(1000 * seed + rank)
    x = rng.standard_normal((n, D))
    w = np.random.default_rng(12345).standard_normal(D)
    y = np.sin(x @ w) + 0.1 * rng.standard_normal(n)
    return x, y


def make_replicated(seed: int = 7):
    """Z (first M rows of a seeded draw), y_Z, hyper-parameters and the tanh-FCN mean weights: same on all ranks."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((M, D))
    w = np.random.default_rng(12345).standard_normal(D)
    yz = np.sin(z @ w) + 0.1 * rng.standard_normal(M)
    fcn = dict(w1=rng.standard_normal((D, 10)) / np.sqrt(D), b1=0.1 * rng.standard_normal(10),
               w2=rng.standard_normal((10, 1)) / np.sqrt(10), b2=np.zeros(1))
    theta = dict(ls_q=0.0, ll_q=np.full(D, np.log(1.0 / D)), ls_p=0.0, ll_p=np.full(D, np.log(1.0 / D)), log_noise=0.0)
    return z, yz, fcn, theta


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference arithmetic) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_step(x, y, z, yz, fcn, theta):
    """One objective evaluation of the reference (src/generalised_variational_inference.py:73-94) in NumPy/SciPy f64 on
    all host cores: the oracle's functions with its BLAS-form Gram (a jit-compiled reference fuses the per-dimension
    loop the faithful `ard_gram` spells out in Python; see oracle/gp_oracle.py:ard_gram_blas)."""
    from oracle import gp_oracle as O

    th = os.cpu_count() or 1
    gq = lambda a, b: O.ard_gram_blas(theta["ls_q"], theta["ll_q"], a, b, threads=th)
    dq = lambda a: O.ard_gram_diag(theta["ls_q"], theta["ll_q"], a)
    gp = lambda a, b: O.ard_gram_blas(theta["ls_p"], theta["ll_p"], a, b, threads=th)
    dp = lambda a: O.ard_gram_diag(theta["ls_p"], theta["ll_p"], a)
    m_q = (np.tanh(x @ fcn["w1"] + fcn["b1"]) @ fcn["w2"] + fcn["b2"]).reshape(-1)
    v_q = O.sparse_posterior_diag(gq, dq, z, x, LAMBDA, False)
    m_p, v_p = O.exact_gp_posterior(gp, dp, np.zeros(x.shape[0]), z, yz, theta["log_noise"], x)
    return O.gaussian_nll(y, m_q, v_q) + O.projected_regularisation("renyi", m_p, v_p, m_q, v_q, RENYI_ALPHA)


def time_cpu(steps: int, warmup: int):
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits

        limiter = threadpool_limits(limits=cores)
    except Exception:  # pragma: no cover
        limiter = None
    z, yz, fcn, theta = make_replicated()
    x, y = make_inputs(0, CPU_SAMPLE_ROWS, seed=1)
    for _ in range(warmup):
        cpu_step(x, y, z, yz, fcn, theta)
    times = []
    loss = None
    for _ in range(steps):
        t = time.perf_counter()
        loss = cpu_step(x, y, z, yz, fcn, theta)
        times.append(time.perf_counter() - t)
    del limiter
    sec = float(np.mean(times))
    return dict(value=CPU_SAMPLE_ROWS / sec, unit=UNIT, cores=cores, kind="port",
                sample=f"{CPU_SAMPLE_ROWS} rows x M={M} (the reference's batch size), mean of {steps} step(s), "
                       f"NumPy/SciPy restatement of the reference (not JAX; Gram through BLAS, all cores), loss={loss:.6f}"), sec


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=2)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def hbm_peak_gbs() -> float:
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        return float(json.load(open(mp)).get("hbm_gbs", 6558.1))
    return 6650.0  # fallback stated in B200_PROFILING.md


def timed_ms(torch, fn, reps=1, sync_all=None):
    """CUDA-event time of reps calls of fn on the current stream (ms per call)."""
    if sync_all:
        sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from gvi_gaussian_process_b200 import _lib, distributed
    from gvi_gaussian_process_b200.src import _lib_proxy as L
    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
    from gvi_gaussian_process_b200.src.gps import ApproximateGPRegression, GPRegression
    from gvi_gaussian_process_b200.src.inducing_points_selection import ConditionalVarianceInducingPointsSelector
    from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean
    from gvi_gaussian_process_b200.src.regularisations.projected import ProjectedRenyiRegularisation
    from gvi_gaussian_process_b200.src.utils import jax_prng

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    comm = distributed.Comm()  # gvi_comm: NCCL + peer windows; the data plane of every multi-rank number below

    z, yz, fcn, theta = make_replicated()
    # two rotating input batches: 2 x (64 + 8) MB > 126 MB of L2 together with the factor workspaces
    host_batches = [make_inputs(rank, N_PER_GPU, seed=s) for s in (1, 2)]
    pinned = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()) for x, y in host_batches]
    dev_batches = [(x.to(dev), y.to(dev)) for x, y in pinned]
    fcn_d = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in fcn.items()}

    def fcn_mean(params, x):  # 1-hidden-layer tanh FCN of width 10 (experiments/*/base_configs/train_approximate/mean)
        return torch.addmm(params["b2"], torch.tanh(torch.addmm(params["b1"], x, params["w1"])), params["w2"])

    kernel = ARDKernel(number_of_dimensions=D)
    approx_mean = CustomMean(mean_function=fcn_mean)
    sparse = SparsePosteriorKernel(base_kernel=kernel, inducing_points=z, diagonal_regularisation=LAMBDA)
    approx_gp = ApproximateGPRegression(mean=approx_mean, kernel=sparse)
    to_d = lambda v: torch.as_tensor(np.atleast_1d(v), dtype=torch.float64, device=dev)
    approx_params = approx_gp.Parameters.construct(
        mean=approx_mean.Parameters.construct(custom=fcn_d),
        kernel=sparse.Parameters.construct(base_kernel=kernel.Parameters.construct(
            log_scaling=to_d(theta["ls_q"]), log_lengthscales=to_d(theta["ll_q"]))))
    const_mean = ConstantMean()
    exact_gp = GPRegression(mean=const_mean, kernel=kernel, x=z, y=yz)
    exact_params = exact_gp.Parameters.construct(
        log_observation_noise=to_d(theta["log_noise"]), mean=const_mean.Parameters.construct(constant=to_d(0.0)),
        kernel=kernel.Parameters.construct(log_scaling=to_d(theta["ls_p"]), log_lengthscales=to_d(theta["ll_p"])))
    risk = NegativeLogLikelihood(gp=approx_gp)
    reg = ProjectedRenyiRegularisation(gp=approx_gp, regulariser=exact_gp, regulariser_parameters=exact_params,
                                       mode="posterior", alpha=RENYI_ALPHA)
    gvi = GeneralisedVariationalInference(regularisation=reg, empirical_risk=risk)
    dgvi = distributed.DistributedGVI(gvi, comm)

    def step(x, y):
        """The reference's public call (GeneralisedVariationalInference.calculate_loss, src/...inference.py:96-119); with
        several ranks its distributed wrapper: every rank passes its slice, the sums cross NVLink once."""
        with torch.no_grad():
            return gvi.calculate_loss(approx_params, x, y) if world == 1 else dgvi.calculate_loss(approx_params, x, y)

    def grad_step(x, y):
        """value and gradient (what jax.grad of calculate_loss gives the reference's trainer): loss,
        d/dlog_scaling, d/dlog_lengthscales[D] all-reduced as one packed (D+4)-vector; d/dm_q back-propagated
        through the tanh-FCN mean by the host framework (torch autograd), its parameter grads all-reduced."""
        out, dm, m_q = gvi.loss_and_grad_sums(approx_params, x, y)
        comm.allreduce_sum(out)
        m_q.backward(dm / out[2])
        flat = torch.cat([p_.grad.reshape(-1) for p_ in fcn_d.values()])
        comm.allreduce_sum(flat)
        g = torch.cat([out[3:] / out[2], flat])
        for p_ in fcn_d.values():
            p_.grad = None
        return (out[0] + out[1]) / out[2], g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- FP64 tensor-pipe peak of THIS device (the roofline denominator), ~50 ms --------------------------------
    scratch = torch.empty(lib.gvi_fp64_tensor_peak_scratch_bytes(), dtype=torch.uint8, device=dev)
    import ctypes

    peak_here = ctypes.c_double(0.0)
    _lib.check(lib.gvi_fp64_tensor_peak_tflops(50.0, ctypes.byref(peak_here), scratch.data_ptr(), _lib.current_stream()),
               "gvi_fp64_tensor_peak_tflops")
    peak_here = float(peak_here.value)
    fp64_peak = max(peak_here, FP64_PEAK_TFLOPS)  # the harder denominator of the two

    # ---- device-resident throughput ("value") ---------------------------------------------------------------------
    for i in range(args.warmup):
        loss = step(*dev_batches[i % 2])
    barrier()
    launches0 = lib.gvi_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            loss = step(*dev_batches[i % 2])
        e1.record()
        barrier()
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = lib.gvi_launch_count() - launches0
    loss_value = float(loss.item())

    # ---- forward + gradient step (reported beside the forward number) -----------------------------------------
    for p_ in fcn_d.values():
        p_.requires_grad_(True)
    for i in range(3):
        grad_step(*dev_batches[i % 2])
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    grad_steps = max(2, min(args.steps, 5))
    g0.record()
    for i in range(grad_steps):
        gl, gvec = grad_step(*dev_batches[i % 2])
    g1.record()
    barrier()
    grad_ms = max_over_ranks(g0.elapsed_time(g1)) / grad_steps
    grad_norm = float(gvec.norm().item())
    for p_ in fcn_d.values():
        p_.requires_grad_(False)

    # ---- the dominant kernel alone (roofline): events around the fused launch only, factors prepared before ------
    f_q = sparse.factorise(approx_params.kernel)
    f_p = exact_gp.factorise(exact_params)
    kt = []
    for i in range(max(3, min(args.steps, 10))):
        x, y = dev_batches[i % 2]
        m_q = approx_mean.predict(approx_params.mean, x).contiguous()
        pm = const_mean.predict(exact_params.mean, x).contiguous()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        L.fused_step(f_q, f_p, x, y, m_q, pm, None, None, L.REG_KINDS["renyi"], RENYI_ALPHA)
        k1.record()
        torch.cuda.synchronize()
        kt.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kt[1:]))
    flops = algorithmic_flops_per_point(M, D) * N_PER_GPU
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    hbm_peak = hbm_peak_gbs()

    # ---- multi-GPU parity, checked in the run the numbers come from ------------------------------------------------
    parity = None
    if world > 1:
        parity = multi_gpu_parity(torch, dist, comm, L, distributed, gvi, dgvi, approx_params, dev_batches[0], to_d, theta,
                                  rank, world)

    # ---- greedy conditional-variance selection: N_PER_GPU points per rank, sharded over all ranks ----------------------
    sel_m = 512
    try:
        sel = bench_selector(torch, comm, L, distributed, dev_batches[0][0], to_d, theta, rank, world, sel_m, barrier,
                             max_over_ranks, hbm_peak)
    except Exception as e:  # pragma: no cover - reported, never hidden
        sel = {"error": f"{type(e).__name__}: {e}"}
    # the public call, end to end: PRNG permutation (device), gather, selection (N_PER_GPU points in total)
    try:
        selector = ConditionalVarianceInducingPointsSelector(comm=comm if world > 1 else None)
        kp = kernel.Parameters.construct(log_scaling=to_d(theta["ls_q"]), log_lengthscales=to_d(theta["ll_q"]))
        x_full = dev_batches[1][0]  # every rank passes the same full inputs in the multi-GPU form of the public call
        if world > 1:
            dist.broadcast(x_full, 0)
        selector.compute_inducing_points(jax_prng.PRNGKey(0), x_full, sel_m, kernel, kp)
        barrier()
        t0 = time.perf_counter()
        _, pub_idx = selector.compute_inducing_points(jax_prng.PRNGKey(0), x_full, sel_m, kernel, kp)
        torch.cuda.synchronize()
        sel["compute_inducing_points_public_api"] = {
            "n": N_PER_GPU, "m": sel_m, "ms": max_over_ranks((time.perf_counter() - t0) * 1e3),
            "note": "wall clock of ConditionalVarianceInducingPointsSelector.compute_inducing_points incl. the "
                    "jax.random.permutation restatement on the device (threefry kernel + stable sort), the gather and "
                    "the selection; the same N in total at every rank count (strong scaling)",
            "unique_indices": int(torch.unique(pub_idx).numel())}
    except Exception as e:  # pragma: no cover
        sel["compute_inducing_points_public_api"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- configs[4]: scaling sweep, inducing selection + predictive variance over the global N sharded by rank ------------
    try:
        config5 = bench_config5(torch, comm, L, distributed, dev, to_d, theta, rank, world, barrier, max_over_ranks,
                                hbm_peak, fp64_peak, quick=args.quick or args.short)
    except Exception as e:  # pragma: no cover
        config5 = {"error": f"{type(e).__name__}: {e}"}

    # ---- HBM-bound rows of the path on their own: the materialised Gram and the risk reduction ---------------------------
    hbm_rows = {}
    if rank == 0:
        try:
            hbm_rows = bench_hbm_rows(torch, L, kernel, dev, to_d, theta, hbm_peak)
        except Exception as e:  # pragma: no cover
            hbm_rows = {"error": f"{type(e).__name__}: {e}"}
    barrier()

    # ---- widened rows, reported beside the headline: never part of `value` ------------------
    widened = {}
    if rank == 0 and not args.quick:
        try:
            widened["classification_config4"] = bench_classification(dev)
        except Exception as e:  # pragma: no cover - reported, never hidden
            widened["classification_config4"] = {"error": f"{type(e).__name__}: {e}"}
        try:
            widened["exact_gp_nll_step"] = bench_exact_nll_step(dev, z, yz, theta)
        except Exception as e:  # pragma: no cover
            widened["exact_gp_nll_step"] = {"error": f"{type(e).__name__}: {e}"}
    barrier()

    # ---- end to end through the public API with HOST buffers ---------------------------------------------------
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    for i in range(2):
        float(step(*pinned[i % 2]).item())
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        loss_e2e = float(step(*pinned[i % 2]).item())  # H2D of x, y inside; D2H of the loss
    torch.cuda.synchronize()
    e2e_sec = max_over_ranks(time.perf_counter() - t0) / e2e_steps

    result = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = time_cpu(steps=4, warmup=1)
            cpu["cpu_model"] = cpu_model()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("gp_step_kernel_dram_bytes_per_launch")
        grad_flops = algorithmic_flops_per_point(M, D) + 2.0 * M * M + 2.0 * M * D
        result = {
            "metric": METRIC, "value": N_PER_GPU * world / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[2]: synthetic UCI-shaped regression N={N_PER_GPU} per GPU, D={D}, "
                                   f"M={M}, ARD, Gaussian NLL + projected Renyi(0.5) vs exact-GP posterior, "
                                   f"forward step incl. both factorisations",
                       "n_per_gpu": N_PER_GPU, "d": D, "m": M, "jitter": LAMBDA,
                       "l2_policy": "2 rotating input batches (2 x 72 MB) plus factor state > 126 MB L2",
                       "parallelism": f"points sharded over {world} rank(s), Z and theta replicated; sums all-reduced by "
                                      f"gvi_comm (NCCL), selector exchange over "
                                      f"{'CUDA-IPC peer windows' if comm.has_peer_window else 'ncclAllGather'}",
                       "api": "GeneralisedVariationalInference.calculate_loss" + (" via distributed.DistributedGVI"
                                                                                 if world > 1 else "")},
            "loss": loss_value, "gpu_launches": int(launches),
            "forward_plus_gradient": {"value": N_PER_GPU * world / (grad_ms * 1e-3), "unit": UNIT,
                                      "ms_per_step": grad_ms, "steps": grad_steps, "grad_norm": grad_norm,
                                      "algorithmic_flops_per_point": grad_flops,
                                      "roofline_frac": grad_flops * N_PER_GPU / (grad_ms * 1e-3) / (fp64_peak * 1e12),
                                      "note": "value and gradient w.r.t. log_scaling, log_lengthscales[D] and the "
                                              "tanh-FCN mean parameters (autograd through the mean on the host "
                                              "framework), factorisations included"},
            "e2e": {"value": N_PER_GPU * world / e2e_sec, "unit": UNIT,
                    "h2d_bytes_per_step": int(N_PER_GPU * (D + 1) * 8), "d2h_bytes_per_step": 8,
                    "steps": e2e_steps, "loss": loss_e2e},
            "roofline": {"bound": "tensor", "kernel": "gp_step_kernel<3,8> (FP64 DMMA m8n8k4)",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": traffic, "kernel_ms": kernel_ms,
                         "algorithmic_flops_per_point": algorithmic_flops_per_point(M, D),
                         "peak_measured_here": peak_here, "peak_constant_r01": FP64_PEAK_TFLOPS,
                         "peak_source": "max(FP64 DMMA m8n8k4 issue rate measured in this run by "
                                        "gvi_fp64_tensor_peak_tflops, the r01 measurement 37.1 in "
                                        "profiles/fp64_peak_r01.txt); cuBLAS DGEMM 8192^3 = 35.4; MEASURED_PEAKS.json has "
                                        "no FP64 entry"},
            "selector": sel,
            "config5": config5,
            "hbm_rows": hbm_rows,
            "widened": widened,
            "clocks": clocks.summary(),
        }
        if parity is not None:
            result["multi_gpu_parity_rel"] = parity["loss_rel"]
            result["multi_gpu_parity"] = parity
        if cpu is not None:
            result["cpu_baseline"] = cpu
    comm.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


def multi_gpu_parity(torch, dist, comm, L, distributed, gvi, dgvi, approx_params, batch, to_d, theta, rank, world):
    """A 2e5-point subsample evaluated twice: sharded over the ranks (the path every number of this run takes) and once
    more on rank 0 alone.  Loss to <= 1e-12 relative; selector pivots bit-identical."""
    sub = 200_000 // world
    xs, ys = batch[0][:sub].contiguous(), batch[1][:sub].contiguous()
    with torch.no_grad():
        loss_sharded = dgvi.calculate_loss(approx_params, xs, ys)
    x_all = comm.allgather(xs).reshape(world * sub, -1)
    y_all = comm.allgather(ys).reshape(-1)
    sel_m = 96
    ls, ll = to_d(theta["ls_q"]), to_d(theta["ll_q"])
    idx_sharded, _ = distributed.ShardedSelector(comm).select(xs, rank * sub, sel_m, ls, ll)
    out = {"points": world * sub, "selector_m": sel_m}
    flag = torch.zeros(2, dtype=torch.float64, device=xs.device)
    if rank == 0:
        with torch.no_grad():
            loss_single = gvi.calculate_loss(approx_params, x_all, y_all)
        idx_single, _ = L.cond_var_select(x_all, sel_m, ls, ll, 1e-12)
        flag[0] = (loss_sharded - loss_single).abs() / loss_single.abs()
        flag[1] = 1.0 if torch.equal(idx_sharded, idx_single) else 0.0
    dist.broadcast(flag, 0)
    out["loss_rel"] = float(flag[0].item())
    out["loss_ok_1e-12"] = bool(out["loss_rel"] <= 1e-12)
    out["selector_pivots_bit_identical"] = bool(flag[1].item() == 1.0)
    return out


def bench_selector(torch, comm, L, distributed, x_local, to_d, theta, rank, world, sel_m, barrier, max_over_ranks,
                   hbm_peak):
    """BASELINE.json's second metric (selector points*M/s): every rank owns N_PER_GPU permuted points (weak scaling); with
    several ranks the one-call sharded selector (exchange inside the kernels), at one rank the single-shard call."""
    ls, ll = to_d(theta["ls_q"]), to_d(theta["ll_q"])
    n_local = x_local.shape[0]
    if world > 1:
        sharded = distributed.ShardedSelector(comm)
        run = lambda: sharded.select(x_local, rank * n_local, sel_m, ls, ll)
    else:
        run = lambda: L.cond_var_select(x_local, sel_m, ls, ll, 1e-12)
    run()  # warm-up at full size: the timed call must not include the first cudaMalloc of the 4 GB C matrix
    ms = None
    for _ in range(2):  # best of two (a pivot chain of 511 dependent steps picks up every hiccup of every rank)
        t_ms, (sel_idx, _) = timed_ms(torch, run, sync_all=barrier)
        t_ms = max_over_ranks(t_ms)
        ms = t_ms if ms is None else min(ms, t_ms)
    n_total = n_local * world
    sel_bytes = 4.0 * n_total * sel_m * sel_m + 32.0 * n_total * sel_m
    gbs = sel_bytes / (ms * 1e-3) / 1e9 / world
    return {"n": n_total, "n_per_gpu": n_local, "m": sel_m, "ms": ms, "points_times_m_per_s": n_total * sel_m / (ms * 1e-3),
            "sharded_over_ranks": world,
            "exchange": ("none (single shard)" if world == 1 else
                         "in-kernel stores into CUDA-IPC peer windows" if comm.has_peer_window else
                         "ncclAllGather enqueued from C"),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s per GPU",
                         "frac": gbs / hbm_peak, "algorithmic_bytes": sel_bytes},
            "unique_pivots": int(torch.unique(sel_idx).numel()), "timed_out": bool(int(sel_idx.min()) < 0)}


def bench_config5(torch, comm, L, distributed, dev, to_d, theta, rank, world, barrier, max_over_ranks, hbm_peak,
                  fp64_peak, quick=False):
    """BASELINE.json configs[4]: N = 1e5 - 1e7 (GLOBAL, sharded contiguously over the ranks), D = 8, M = 512 - 4096:
    greedy selection (C = 8 N M bytes sharded over the ranks: N = 1e7 with M >= 2048 only fits on several GPUs) followed by
    the predictive variance of the sparse posterior over the selected points."""
    ls, ll = to_d(theta["ls_q"]), to_d(theta["ll_q"])
    total_mem = torch.cuda.get_device_properties(dev).total_memory
    rows = []
    sizes = [(100_000, 512), (1_000_000, 1024)] if quick else [
        (n, m) for n in (100_000, 1_000_000, 10_000_000) for m in (512, 1024, 2048, 4096)]
    for n, m in sizes:
        lo, hi = distributed.shard_bounds(n, rank, world)
        row = {"n": n, "m": m, "n_per_gpu": hi - lo}
        c_bytes = 8.0 * (hi - lo) * (m - 1)
        rng = np.random.default_rng(10_000 + rank)
        x = torch.as_tensor(rng.standard_normal((hi - lo, D)), dtype=torch.float64, device=dev)
        z = None
        if c_bytes > 0.8 * total_mem:
            row["selection"] = (f"skipped: C = {8.0 * n * m / 1e9:.0f} GB does not fit {world} GPU(s) "
                                f"({c_bytes / 1e9:.0f} GB per rank > 0.8 x {total_mem / 1e9:.0f} GB)")
        else:
            if world > 1:
                sharded = distributed.ShardedSelector(comm)
                run = lambda: sharded.select(x, lo, m, ls, ll)
            else:
                run = lambda: L.cond_var_select(x, m, ls, ll, 1e-12)
            if n <= 1_000_000 and m <= 1024:
                run()  # small cases: take the allocation out of the timing; the big ones are dominated by HBM streaming
            ms, (idx, _) = timed_ms(torch, run, sync_all=barrier)
            ms = max_over_ranks(ms)
            by = 4.0 * n * m * m + 32.0 * n * m
            row["selection"] = {"ms": ms, "points_times_m_per_s": n * m / (ms * 1e-3),
                                "hbm_gbs_per_gpu": by / (ms * 1e-3) / 1e9 / world,
                                "hbm_frac": by / (ms * 1e-3) / 1e9 / world / hbm_peak,
                                "unique_pivots": int(torch.unique(idx).numel())}
            # the selected points, replicated: the owner of each pivot contributes its row, everyone else zeros
            mine = (idx >= lo) & (idx < hi)
            z = torch.zeros((m, D), dtype=torch.float64, device=dev)
            z[mine] = x[idx[mine] - lo]
            comm.allreduce_sum(z)
            del idx
        torch.cuda.empty_cache()
        if z is None:  # selection did not fit: predictive variance over a random inducing set of the same size
            z = torch.as_tensor(np.random.default_rng(77).standard_normal((m, D)), dtype=torch.float64, device=dev)
            row["inducing_points"] = "random (selection skipped)"
        factor = L.GpFactor(m, D)

        def predictive():
            f = factor.factor(z, ls, ll, LAMBDA, False)
            return L.gp_predict(f, x, want_mean=False)[1]

        predictive()
        ms, var = timed_ms(torch, predictive, sync_all=barrier)
        ms = max_over_ranks(ms)
        fl = (float(m) * m + 4.0 * m * D) * n
        row["predictive_variance"] = {"ms": ms, "points_per_s": n / (ms * 1e-3),
                                      "tflops_per_gpu": fl / (ms * 1e-3) / 1e12 / world,
                                      "fp64_tensor_frac": fl / (ms * 1e-3) / 1e12 / world / fp64_peak,
                                      "min_variance": float(var.min().item())}
        rows.append(row)
        del x, z, var, factor
        torch.cuda.empty_cache()
    return {"workload": "configs[4]: N global, sharded over the ranks; selection then predictive variance (factorisation "
                        "included); times are max over ranks", "n_gpus": world, "rows": rows}


def bench_hbm_rows(torch, L, kernel, dev, to_d, theta, hbm_peak):
    """north_star bullets 1 and 3 on their own (single GPU): the materialised ARD Gram (store-bound: 8 N M bytes written)
    and the risk + regulariser reduction (reads 40 bytes per point), both against the measured HBM copy bandwidth."""
    out = {}
    rng = np.random.default_rng(5)
    kp = kernel.Parameters.construct(log_scaling=to_d(theta["ls_q"]), log_lengthscales=to_d(theta["ll_q"]))
    n, m = 400_000, 1024
    x = torch.as_tensor(rng.standard_normal((n, D)), dtype=torch.float64, device=dev)
    zz = torch.as_tensor(rng.standard_normal((m, D)), dtype=torch.float64, device=dev)
    warm = [kernel.calculate_gram(kp, x, zz) for _ in range(2)]  # two result buffers in the caching allocator: the timed
    del warm                                                      # calls then allocate without a cudaMalloc
    ms, g = timed_ms(torch, lambda: kernel.calculate_gram(kp, x, zz), reps=4)
    out["materialised_gram"] = {"n": n, "m": m, "d": D, "ms": ms, "gb_written": 8.0 * n * m / 1e9,
                                "store_gbs": 8.0 * n * m / (ms * 1e-3) / 1e9,
                                "hbm_frac": 8.0 * n * m / (ms * 1e-3) / 1e9 / hbm_peak,
                                "api": "ARDKernel.calculate_gram (output allocation included)"}
    del g
    torch.cuda.empty_cache()
    for n in (1_000_000, 10_000_000, 100_000_000):
        v = [torch.rand(n, dtype=torch.float64, device=dev) + 0.5 for _ in range(5)]
        y, mq, vq, mp, vp = v
        L.risk(L.REG_KINDS["renyi"], RENYI_ALPHA, y, mq, vq, mp, vp)
        ms, _ = timed_ms(torch, lambda: L.risk(L.REG_KINDS["renyi"], RENYI_ALPHA, y, mq, vq, mp, vp), reps=5)
        out[f"risk_reduction_n{n:.0e}".replace("+0", "")] = {
            "n": n, "ms": ms, "read_gbs": 40.0 * n / (ms * 1e-3) / 1e9, "hbm_frac": 40.0 * n / (ms * 1e-3) / 1e9 / hbm_peak}
        del v, y, mq, vq, mp, vp
        torch.cuda.empty_cache()
    return out


def bench_classification(dev):
    """BASELINE.json configs[3] (synthetic MNIST-shaped classification): N=6e4, D=784, K=10 classes, per-class ARD
    sparse posteriors over M=2048 inducing points each, multinomial NLL + GWI (no eigendecomposition) against the
    exact GP classifier on a combined inducing set of 2048 points - one forward objective = 20 predictive passes
    (tensor-pipe Gram + tile kernel) + the Gauss-Hermite class-probability kernel + the reductions."""
    import torch

    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
    from gvi_gaussian_process_b200.src.gps import ApproximateGPClassification, GPClassification
    from gvi_gaussian_process_b200.src.kernels import MultiOutputKernel
    from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean
    from gvi_gaussian_process_b200.src.regularisations import GaussianWassersteinRegularisation
    from gvi_gaussian_process_b200.src import _lib_proxy as L

    n, d, k, m = 60_000, 784, 10, 2048
    rng = np.random.default_rng(4)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    x = t(rng.uniform(0.0, 1.0, (n, d)))
    labels = rng.integers(0, k, n)
    y = t(np.eye(k)[labels])
    ll = t(np.full(d, np.log(1.0 / d)))
    kern_q = MultiOutputKernel(kernels=[
        SparsePosteriorKernel(base_kernel=ARDKernel(number_of_dimensions=d),
                              inducing_points=x[torch.as_tensor(rng.permutation(n)[:m], device=dev)].contiguous(),
                              diagonal_regularisation=LAMBDA) for _ in range(k)])
    w = t(rng.standard_normal((d, k)) / np.sqrt(d))
    mean_q = CustomMean(mean_function=lambda p, xx: xx @ p, number_output_dimensions=k)
    approx = ApproximateGPClassification(mean=mean_q, kernel=kern_q)
    params = {"mean": {"custom": w},
              "kernel": {"kernels": [{"base_kernel": {"log_scaling": t(0.0), "log_lengthscales": ll}}] * k}}
    zi = torch.as_tensor(rng.permutation(n)[:m], device=dev)
    exact = GPClassification(mean=ConstantMean(number_output_dimensions=k),
                             kernel=MultiOutputKernel(kernels=[ARDKernel(number_of_dimensions=d) for _ in range(k)]),
                             x=x[zi].contiguous(), y=y[zi].contiguous())
    eparams = {"log_observation_noise": t(np.zeros(k)), "mean": {"constant": t(np.zeros(k))},
               "kernel": {"kernels": [{"log_scaling": t(0.0), "log_lengthscales": ll}] * k}}
    gvi = GeneralisedVariationalInference(
        regularisation=GaussianWassersteinRegularisation(gp=approx, regulariser=exact, regulariser_parameters=eparams,
                                                         mode="posterior"),
        empirical_risk=NegativeLogLikelihood(gp=approx))
    with torch.no_grad():
        loss = gvi.calculate_loss(params, x, y)  # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 2
        e0.record()
        for _ in range(steps):
            loss = gvi.calculate_loss(params, x, y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # the quadrature kernel alone, on the predictive of q
        g = approx.calculate_prediction_gaussian(params, x, full_covariance=False)
        mq, vq = g.mean.reshape(k, -1).contiguous(), g.covariance.reshape(k, -1).contiguous()
        roots, weights = approx._hermite_device()
        L.class_probabilities(mq, vq, roots, weights, approx.epsilon, approx.cdf_lower_bound)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            L.class_probabilities(mq, vq, roots, weights, approx.epsilon, approx.cdf_lower_bound)
        c1.record()
        torch.cuda.synchronize()
        cms = c0.elapsed_time(c1) / 5
    # value + gradient of the same objective w.r.t. the mean weights, K log-scalings and K x 784 log-lengthscales
    w.requires_grad_(True)
    leaves = [(t(0.0).requires_grad_(True), ll.clone().requires_grad_(True)) for _ in range(k)]
    gparams = {"mean": {"custom": w},
               "kernel": {"kernels": [{"base_kernel": {"log_scaling": a, "log_lengthscales": b}} for a, b in leaves]}}
    gvi.calculate_loss(gparams, x, y).backward()  # warm-up
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    gloss = gvi.calculate_loss(gparams, x, y)
    gloss.backward()
    g1.record()
    torch.cuda.synchronize()
    grad_ms = g0.elapsed_time(g1)
    grad_norm = float(torch.cat([b.grad for _, b in leaves]).norm())
    cdf_evals = float(n) * 50 * k * (k - 1)
    flops = 2.0 * k * n * (2.0 * m * d + m * m) + 2.0 * k * n * m  # q and p: Gram contraction + triangular product
    return {"n": n, "d": d, "classes": k, "m": m, "ms_per_objective": ms, "points_per_s": n / (ms * 1e-3),
            "loss": float(loss), "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
            "value_and_gradient_ms": grad_ms, "lengthscale_grad_norm": grad_norm,
            "class_probability_kernel": {"ms": cms, "normal_cdf_evaluations_per_s": cdf_evals / (cms * 1e-3),
                                         "points_per_s": n / (cms * 1e-3)},
            "note": "forward objective NLL + GWI(no eig): 10 q-passes + 10 p-passes (factorisation, tensor-pipe "
                    "Gram for D=784, tile kernel) + Gauss-Hermite class probabilities + reductions"}


def bench_exact_nll_step(dev, z, yz, theta):
    """one exact-GP NLL training step on the inducing set (value and gradient w.r.t.
    log_scaling, log_lengthscales[D], log_observation_noise and the prior-mean constant), M = N = 1024."""
    import torch

    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.gps import GPRegression
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean

    t = lambda a: torch.as_tensor(np.atleast_1d(a), dtype=torch.float64, device=dev).requires_grad_(True)
    gp = GPRegression(mean=ConstantMean(), kernel=ARDKernel(number_of_dimensions=D), x=z, y=yz)
    leaves = dict(ls=t(theta["ls_p"]), ll=t(theta["ll_p"]), noise=t(theta["log_noise"]), c=t(0.0))
    params = {"log_observation_noise": leaves["noise"], "mean": {"constant": leaves["c"]},
              "kernel": {"log_scaling": leaves["ls"], "log_lengthscales": leaves["ll"]}}
    nll = NegativeLogLikelihood(gp=gp)
    xz = torch.as_tensor(z, dtype=torch.float64, device=dev)
    yy = torch.as_tensor(yz, dtype=torch.float64, device=dev)

    def one():
        for v in leaves.values():
            v.grad = None
        loss = nll.calculate_empirical_risk(params, xz, yy)
        loss.backward()
        return loss

    one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return {"m": int(z.shape[0]), "n": int(z.shape[0]), "ms_per_step": ms, "loss": float(loss.detach()),
            "grad_log_scaling": float(leaves["ls"].grad), "grad_log_noise": float(leaves["noise"].grad)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cpu, sec = time_cpu(steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    cpu["cpu_model"] = cpu_model()
    return {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"configs[2]: synthetic UCI-shaped regression, D={D}, M={M}, ARD, Gaussian NLL + "
                               f"projected Renyi(0.5) vs exact-GP posterior; each step = a {CPU_SAMPLE_ROWS}-row "
                               f"sample of the workload on the host cores",
                   "d": D, "m": M, "jitter": LAMBDA},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference (NumPy/SciPy float64 oracle), not JAX: jax==0.4.13 is not "
                "installable in this image (SURVEY.md F4)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="profiling runs: skip the widened rows, shorten the config5 sweep")
    ap.add_argument("--short", action="store_true", help="launch-list runs: shorten the config5 sweep, keep the widened rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    result = run_reference(args) if args.impl == "reference" else run_gpu(args)
    if result is not None:
        print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
