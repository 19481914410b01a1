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
    from oracle import gp_oracle as O

    gq = lambda a, b: O.ard_gram(theta["ls_q"], theta["ll_q"], a, b)
    dq = lambda a: O.ard_gram_diag(theta["ls_q"], theta["ll_q"], a)
    gp = lambda a, b: O.ard_gram(theta["ls_p"], theta["ll_p"], a, b)
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
                       f"NumPy/SciPy restatement of the reference (not JAX), loss={loss:.6f}"), sec


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
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from gvi_gaussian_process_b200 import _lib, distributed
    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
    from gvi_gaussian_process_b200.src.gps import ApproximateGPRegression, GPRegression
    from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean
    from gvi_gaussian_process_b200.src.regularisations.projected import ProjectedRenyiRegularisation

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

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

    def step(x, y):
        out3 = gvi.loss_sums(approx_params, x, y)
        distributed.allreduce_loss_sums(out3)
        return distributed.loss_from_sums(out3)

    def grad_step(x, y):
        """value and gradient (what jax.grad of calculate_loss gives the reference's trainer): loss,
        d/dlog_scaling, d/dlog_lengthscales[D] all-reduced as one packed (D+4)-vector; d/dm_q back-propagated
        through the tanh-FCN mean by the host framework (torch autograd), its parameter grads all-reduced."""
        out, dm, m_q = gvi.loss_and_grad_sums(approx_params, x, y)
        distributed.allreduce_loss_sums(out)
        m_q.backward(dm / out[2])
        if world > 1:
            for p_ in fcn_d.values():
                dist.all_reduce(p_.grad)
        g = torch.cat([out[3:] / out[2]] + [p_.grad.reshape(-1) for p_ in fcn_d.values()])
        for p_ in fcn_d.values():
            p_.grad = None
        return (out[0] + out[1]) / out[2], g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    ms_total = e0.elapsed_time(e1)
    launches = lib.gvi_launch_count() - launches0
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
    tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    grad_ms = float(tg.item()) / grad_steps
    grad_norm = float(gvec.norm().item())
    for p_ in fcn_d.values():
        p_.requires_grad_(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    loss_value = float(loss.item())

    # ---- the dominant kernel alone (roofline): events around the fused launch only, factors prepared before ------
    from gvi_gaussian_process_b200.src import _lib_proxy as L

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

    # ---- greedy conditional-variance selection (the other metric of BASELINE.json: selector points*M/s) ----------
    sel_m = 512
    sel = None
    try:
        xs_sel = dev_batches[0][0]
        # warm-up at full size: the timed call must not include the first cudaMalloc of the 4 GB C matrix
        L.cond_var_select(xs_sel, sel_m, to_d(theta["ls_q"]), to_d(theta["ll_q"]), 1e-12)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        sel_idx, _ = L.cond_var_select(xs_sel, sel_m, to_d(theta["ls_q"]), to_d(theta["ll_q"]), 1e-12)
        s1.record()
        torch.cuda.synchronize()
        sel_ms = s0.elapsed_time(s1)
        sel_bytes = 4.0 * N_PER_GPU * sel_m * sel_m + 32.0 * N_PER_GPU * sel_m
        hbm_peak = 6558.1
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            hbm_peak = json.load(open(mp)).get("hbm_gbs", hbm_peak)
        sel = {"n": N_PER_GPU, "m": sel_m, "ms": sel_ms, "points_times_m_per_s": N_PER_GPU * sel_m / (sel_ms * 1e-3),
               "roofline": {"bound": "hbm", "achieved": sel_bytes / (sel_ms * 1e-3) / 1e9, "peak": hbm_peak,
                            "unit": "GB/s", "frac": sel_bytes / (sel_ms * 1e-3) / 1e9 / hbm_peak,
                            "algorithmic_bytes": sel_bytes},
               "unique_pivots": int(torch.unique(sel_idx).numel())}
    except Exception as e:  # pragma: no cover - reported, never hidden
        sel = {"error": f"{type(e).__name__}: {e}"}

    # ---- widened rows, reported beside the headline: never part of `value` ------------------
    widened = {}
    if rank == 0:
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
    e2e_sec = time.perf_counter() - t0
    t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sec = float(t.item()) / e2e_steps

    result = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = time_cpu(steps=2, warmup=1)
            cpu["cpu_model"] = cpu_model()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("gp_step_kernel_dram_bytes_per_launch")
        result = {
            "metric": METRIC, "value": N_PER_GPU * world / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[2]: synthetic UCI-shaped regression N={N_PER_GPU} per GPU, D={D}, "
                                   f"M={M}, ARD, Gaussian NLL + projected Renyi(0.5) vs exact-GP posterior, "
                                   f"forward step incl. both factorisations",
                       "n_per_gpu": N_PER_GPU, "d": D, "m": M, "jitter": LAMBDA,
                       "l2_policy": "2 rotating input batches (2 x 72 MB) plus factor state > 126 MB L2",
                       "parallelism": f"points sharded over {world} rank(s), Z and theta replicated"},
            "loss": loss_value, "gpu_launches": int(launches),
            "forward_plus_gradient": {"value": N_PER_GPU * world / (grad_ms * 1e-3), "unit": UNIT,
                                      "ms_per_step": grad_ms, "steps": grad_steps, "grad_norm": grad_norm,
                                      "algorithmic_flops_per_point": algorithmic_flops_per_point(M, D)
                                      + 2.0 * M * M + 2.0 * M * D,
                                      "roofline_frac": (algorithmic_flops_per_point(M, D) + 2.0 * M * M + 2.0 * M * D)
                                      * N_PER_GPU / (grad_ms * 1e-3) / (FP64_PEAK_TFLOPS * 1e12),
                                      "note": "value and gradient w.r.t. log_scaling, log_lengthscales[D] and the "
                                              "tanh-FCN mean parameters (autograd through the mean on the host "
                                              "framework), factorisations included"},
            "e2e": {"value": N_PER_GPU * world / e2e_sec, "unit": UNIT,
                    "h2d_bytes_per_step": int(N_PER_GPU * (D + 1) * 8), "d2h_bytes_per_step": 8,
                    "steps": e2e_steps, "loss": loss_e2e},
            "roofline": {"bound": "tensor", "kernel": "gp_step_kernel<3,8> (FP64 DMMA m8n8k4)",
                         "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": achieved / FP64_PEAK_TFLOPS, "traffic": traffic, "kernel_ms": kernel_ms,
                         "algorithmic_flops_per_point": algorithmic_flops_per_point(M, D),
                         "peak_source": "measured FP64 DMMA issue rate on this pool's B200 (profiles/fp64_peak_r01.txt; "
                                        "cuBLAS DGEMM 8192^3 = 35.4); MEASURED_PEAKS.json has no FP64 entry"},
            "selector": sel,
            "widened": widened,
            "clocks": clocks.summary(),
        }
        if cpu is not None:
            result["cpu_baseline"] = cpu
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    result = run_reference(args) if args.impl == "reference" else run_gpu(args)
    if result is not None:
        print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
