import sys, numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200.src import _lib_proxy as L
dev='cuda'
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
for N, M, D in [(60000, 2048, 784), (200000, 1024, 64), (1000000, 1024, 8)]:
    X = T(rng.uniform(0,1,(N, D))); Z = T(rng.uniform(0,1,(M, D))); ls = T([0.0]); ll = T(np.full(D, np.log(1.0/D)))
    out = torch.empty(N, M, dtype=torch.float64, device=dev)
    for tp in (True, False):
        for _ in range(2): L.ard_gram(X, Z, ls, ll, out, tensor_pipe=tp)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): L.ard_gram(X, Z, ls, ll, out, tensor_pipe=tp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/3
        print(f"gram N={N} M={M} D={D} tensor_pipe={tp}: {ms:.2f} ms  store {8*N*M/ms*1e-6:.0f} GB/s  contraction {2*N*M*D/ms*1e-9:.2f} TFLOP/s")
