"""Latency-bound shapes: exact-GP NLL value-and-gradient step at N = M (bench.bench_exact_nll_step) and small-batch
predictive variance (N = 256 .. 4096 at M = 1024): python tools/small_time.py"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from gvi_gaussian_process_b200.src import _lib_proxy as L
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
z, yz, fcn, theta = bench.make_replicated()
for _ in range(2):
    r = bench.bench_exact_nll_step(dev, z, yz, theta)
print("exact_gp_nll_step", r, flush=True)
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
rng = np.random.default_rng(0)
M, D = 1024, 8
Z = T(rng.standard_normal((M, D))); ls = T([0.0]); ll = T(np.full(D, np.log(1 / D)))
f = L.GpFactor(M, D).factor(Z, ls, ll, 1e-5, False)
for N in (256, 1024, 2048, 3000, 3552, 4096):
    X = T(rng.standard_normal((N, D)))
    for _ in range(3): L.gp_predict(f, X)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m, v = L.gp_predict(f, X)
    e1.record(); torch.cuda.synchronize()
    print(f"predict N={N} M={M}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us  var sum {float(v.sum()):.12f}", flush=True)
