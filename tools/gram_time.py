"""Materialised ARD Gram timing (CUDA events): python tools/gram_time.py [quick]
Store roofline: 8 N M bytes written once; the fraction is against MEASURED_PEAKS.json's hbm_gbs."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200.src import _lib_proxy as L
dev = 'cuda'
peak = 6558.1
if os.path.exists('MEASURED_PEAKS.json'):
    peak = json.load(open('MEASURED_PEAKS.json')).get('hbm_gbs', peak)
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
quick = len(sys.argv) > 1
shapes = [(400000, 1024, 8)] if quick else [(400000, 1024, 8), (1000000, 512, 8), (400000, 1024, 1), (400000, 1024, 4),
                                            (400000, 1024, 16), (200000, 1024, 20), (200000, 1024, 64), (60000, 2048, 784)]
for N, M, D in shapes:
    X = T(rng.standard_normal((N, D))); Z = T(rng.standard_normal((M, D))); ls = T([0.0]); ll = T(np.full(D, np.log(1.0/D)))
    out = torch.empty(N, M, dtype=torch.float64, device=dev)
    for tp in ((True, False) if D >= int(os.environ.get("GRAM_TP_MIN_D", 5)) else (False,)):
        for _ in range(2): L.ard_gram(X, Z, ls, ll, out, tensor_pipe=tp)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): L.ard_gram(X, Z, ls, ll, out, tensor_pipe=tp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/5
        gbs = 8*N*M/ms*1e-6
        print(f"gram N={N} M={M} D={D} tensor_pipe={tp}: {ms:.3f} ms  store {gbs:.0f} GB/s = {gbs/peak:.3f} of measured HBM "
              f"({peak:.0f})  contraction {2*N*M*D/ms*1e-9:.2f} TFLOP/s", flush=True)
