"""Factorisation chain timing (gvi_gp_factor_f64: Gram, jitter, Cholesky, L^-1, packing, alpha): python tools/factor_only.py
Also checks L L^T = A and L^-1 L = I against torch on the same A (float64)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200.src import _lib_proxy as L
dev = 'cuda'
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
D = int(os.environ.get('D', 8))
rng = np.random.default_rng(0)
for M in [int(m) for m in os.environ.get('MS', '2,40,64,100,512,1024,1056,2048,4096').split(',')]:
    Z = rng.standard_normal((M, D)); lsd = T([0.0]); lld = T(np.full(D, np.log(1 / D))); Zd = T(Z); yz = T(rng.standard_normal(M))
    f = L.GpFactor(M, D)
    for _ in range(2): f.factor(Zd, lsd, lld, 1e-5, False, y_z=yz)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f.factor(Zd, lsd, lld, 1e-5, False, y_z=yz)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    K = torch.empty(M, M, dtype=torch.float64, device=dev); L.ard_gram(Zd, Zd, lsd, lld, K)
    A = K + 1e-5 * torch.trace(K) / M * torch.eye(M, dtype=torch.float64, device=dev)
    Lc = torch.tril(f.cholesky()); Li = torch.tril(f.inverse_cholesky())
    e_chol = float((Lc @ Lc.T - A).abs().max() / A.abs().max())
    e_inv = float((Li @ Lc - torch.eye(M, dtype=torch.float64, device=dev)).abs().max())
    ref = torch.linalg.cholesky(A)
    e_ref = float((Lc - ref).abs().max() / ref.abs().max())
    print(f"M={M}: factor chain {ms:.3f} ms  |LL^T-A|/|A| {e_chol:.1e}  |L^-1 L - I| {e_inv:.1e}  |L - potrf|/|L| {e_ref:.1e}  info {int(f.info.item())}", flush=True)
