"""BASELINE.json configs[4] on one B200: N in {1e5, 1e6, 1e7}, D = 8, M in {512, 1024, 2048, 4096} - predictive variance
(Gram + triangular product, no N x M intermediate in HBM) and greedy conditional-variance selection, each with its
roofline fraction.  Selection needs the (M-1) x N matrix C resident: skipped where 8 N M > 0.8 x 180 GB.
Prints a markdown table (profiles/sweep_config5_r01.md)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gvi_gaussian_process_b200.src import _lib_proxy as L  # noqa: E402

T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
FP64_PEAK, HBM_PEAK = 37.1, 6558.1
if os.path.exists("MEASURED_PEAKS.json"):
    HBM_PEAK = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", HBM_PEAK)
D = 8
rng = np.random.default_rng(0)
ls, ll = T([0.0]), T(np.full(D, np.log(1.0 / D)))
print("| N | M | predictive variance ms | points/s | TFLOP/s (M^2 + 4MD per point) | % of DMMA peak | selector ms | "
      "points*M/s | GB/s (4NM^2 + 32NM) | % of measured HBM |")
print("|---|---|---|---|---|---|---|---|---|---|")
for n in (100_000, 1_000_000, 10_000_000):
    x = torch.randn(n, D, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    for m in (512, 1024, 2048, 4096):
        z = x[torch.randperm(n, device="cuda", generator=torch.Generator("cuda").manual_seed(2))[:m]].contiguous()
        f = L.GpFactor(m, D).factor(z, ls, ll, 1e-5, False)
        L.gp_predict(f, x, want_mean=False)
        torch.cuda.synchronize()
        reps = 3 if n * m * m < 5e13 else 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _, v = L.gp_predict(f, x, want_mean=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = n * (m * m + 4.0 * m * D) / (ms * 1e-3) / 1e12
        sel = "| - | - | - | - |"
        if 8.0 * n * m <= 0.8 * 180e9 and 4.0 * n * m * m <= 1.5e14:
            if n <= 1_000_000:
                L.cond_var_select(x, m, ls, ll, 1e-12)  # warm-up (first cudaMalloc of C)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            idx, _ = L.cond_var_select(x, m, ls, ll, 1e-12)
            s1.record()
            torch.cuda.synchronize()
            sms = s0.elapsed_time(s1)
            gbs = (4.0 * n * m * m + 32.0 * n * m) / (sms * 1e-3) / 1e9
            assert int(torch.unique(idx).numel()) == m
            sel = f"| {sms:.1f} | {n * m / (sms * 1e-3):.3e} | {gbs:.0f} | {100 * gbs / HBM_PEAK:.1f} |"
            del idx
            torch.cuda.empty_cache()
        print(f"| {n:.0e} | {m} | {ms:.2f} | {n / (ms * 1e-3):.3e} | {tf:.2f} | {100 * tf / FP64_PEAK:.1f} {sel}",
              flush=True)
        assert bool(torch.isfinite(v).all()) and float(v.min()) > -1e-8
    del x
    torch.cuda.empty_cache()
