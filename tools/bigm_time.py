import sys, numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200.src import _lib_proxy as L
dev='cuda'
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
for N, M, D in [(1_000_000, 512, 8), (500_000, 2048, 8), (200_000, 4096, 8)]:
    X = T(rng.standard_normal((N, D))); Z = T(rng.standard_normal((M, D))); ls = T([0.0]); ll = T(np.full(D, np.log(1.0)))
    fq = L.GpFactor(M, D).factor(Z, ls, ll, 1e-5, False)
    torch.cuda.synchronize()
    for _ in range(2): L.gp_predict(fq, X, want_mean=False)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(3): m, v = L.gp_predict(fq, X, want_mean=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/3
    print(f"predict N={N} M={M}: {ms:.2f} ms {N/ms*1e3:.3e} pts/s {N*(M*M+4*M*D)/ms*1e-9:.2f} TFLOP/s  info={int(fq.info.item())} vmin={float(v.min()):.2e}")
