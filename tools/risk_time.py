"""Risk + regulariser reduction alone (gvi_risk_f64, 40 B / point read): python tools/risk_time.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200.src import _lib_proxy as L
peak = 6558.1
if os.path.exists('MEASURED_PEAKS.json'):
    peak = json.load(open('MEASURED_PEAKS.json')).get('hbm_gbs', peak)
for n in (1_000_000, 10_000_000, 100_000_000):
    y, mq, vq, mp, vp = [torch.rand(n, dtype=torch.float64, device='cuda') + 0.5 for _ in range(5)]
    for kind in ("renyi", "kl", "bhattacharyya", "hellinger", "gaussian_wasserstein", "squared_difference"):
        for _ in range(2): out = L.risk(L.REG_KINDS[kind], 0.5, y, mq, vq, mp, vp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): out = L.risk(L.REG_KINDS[kind], 0.5, y, mq, vq, mp, vp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"risk n={n:.0e} NLL + {kind}: {ms * 1e3:.1f} us  {40.0 * n / ms * 1e-6:.0f} GB/s = {40.0 * n / ms * 1e-6 / peak:.3f} of HBM  "
              f"sums {out[0].item():.12e} {out[1].item():.12e}", flush=True)
    del y, mq, vq, mp, vp
    torch.cuda.empty_cache()
