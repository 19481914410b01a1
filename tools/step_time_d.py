"""Fused forward step timing over D (N = 1e6, M = 1024): python tools/step_time_d.py [D ...]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200.src import _lib_proxy as L
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')
rng = np.random.default_rng(0)
N, M = 1_000_000, 1024
for D in ([int(a) for a in sys.argv[1:]] or [2, 4, 8, 12, 16, 24, 32]):
    X, Z = T(rng.standard_normal((N, D))), T(rng.standard_normal((M, D)))
    y, mq, yz = T(rng.standard_normal(N)), T(0.3 * rng.standard_normal(N)), T(rng.standard_normal(M))
    ls, ll, ln, pm = T([0.0]), T(np.full(D, np.log(1 / D))), T([0.0]), T(np.zeros(N))
    fq = L.GpFactor(M, D).factor(Z, ls, ll, 1e-5, False)
    fp = L.GpFactor(M, D).factor(Z, ls, ll, 0.0, True, log_noise=ln, y_z=yz)
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = L.fused_step(fq, fp, X, y, mq, pm, None, None, 5, 0.5); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.mean(ts[1:])); fl = (2.0 * (M * M + 2.0 * M * D) + 2.0 * M) * N
    print(f"fused step D={D} M={M}: {ms:.2f} ms  {fl / ms * 1e-9:.2f} TFLOP/s = {fl / ms * 1e-9 / 37.1:.3f} of 37.1  loss sums {out[:2].tolist()}", flush=True)
