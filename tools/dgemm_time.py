"""In-tree FP64 tensor-pipe GEMM timing (csrc/dgemm.cu through the test-hook library) beside torch.matmul (cuBLAS):
python tools/dgemm_time.py"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200 import _lib
lib = _lib.load_test()
dev = 'cuda'
rng = np.random.default_rng(0)
def T(a): return torch.as_tensor(a, dtype=torch.float64, device=dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
st = _lib.current_stream
for (name, ta, tb, m, n, k, kmode, sym) in [
        ("K L^-T   (N,T) triangular", 0, 1, 2048, 2048, 2048, 1, 0), ("T L^-1   (N,N) triangular", 0, 0, 2048, 2048, 2048, 2, 0),
        ("Aa^T B   (T,N) symmetric ", 1, 0, 2048, 2048, 2048, 0, 1), ("W^T [X|1] (T,N)          ", 1, 0, 2048, 785, 2048, 0, 0),
        ("V Z^     (N,N)           ", 0, 0, 2048, 785, 2048, 0, 0), ("K L^-T   M=1024 nr=4096  ", 0, 1, 4096, 1024, 1024, 1, 0),
        ("full     (N,N) 4096^3    ", 0, 0, 4096, 4096, 4096, 0, 0)]:
    a = T(rng.standard_normal((k, m) if ta else (m, k))); b = T(rng.standard_normal((n, k) if tb else (k, n)))
    if kmode: b = torch.tril(b).contiguous()
    c = torch.empty(m, n, dtype=torch.float64, device=dev)
    ours = timed(lambda: lib.gvi_test_dgemm_f64(ta, tb, m, n, k, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), 0.0,
                                                c.data_ptr(), n, kmode, sym, st()))
    full = timed(lambda: lib.gvi_test_dgemm_f64(ta, tb, m, n, k, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), 0.0,
                                                c.data_ptr(), n, 0, 0, st()))
    ref = timed(lambda: torch.matmul(a.T if ta else a, b.T if tb else b, out=c))
    fl = 2.0 * m * n * k
    print(f"{name} m={m} n={n} k={k}: structured {ours:.3f} ms | same kernel, full range {full:.3f} ms = {fl / full * 1e-9:.1f} TFLOP/s | "
          f"cuBLAS {ref:.3f} ms = {fl / ref * 1e-9:.1f} TFLOP/s | speed-up vs cuBLAS {ref / ours:.2f}x", flush=True)
