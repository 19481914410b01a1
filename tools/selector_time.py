"""Greedy selector timing (CUDA events): python tools/selector_time.py  -> ms and algorithmic GB/s for a few (N, M)."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200.src import _lib_proxy as L
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')
rng = np.random.default_rng(0)
for n, m in [(100_000, 512), (1_000_000, 512), (1_000_000, 128)]:
    x = T(rng.standard_normal((n, 8))); ls = T([0.0]); ll = T(np.full(8, np.log(1 / 8)))
    idx, _ = L.cond_var_select(x, m, ls, ll, 1e-12); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); idx, _ = L.cond_var_select(x, m, ls, ll, 1e-12); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); by = 4.0 * n * m * m + 32.0 * n * m
    print(f"N={n} M={m}: {ms:.2f} ms, {by / ms / 1e6:.0f} GB/s, unique pivots {len(set(idx.tolist()))}")
