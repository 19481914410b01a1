"""Times the fused forward launch of config #3 alone (events around gvi_fused_step_f64).  With the profiling switch
GVI_DEBUG_SKIP=1 (no phase 1) / 2 (no phase 2) in the environment the difference gives the per-phase budget."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gvi_gaussian_process_b200.src import _lib_proxy as L  # noqa: E402

N, M, D = 1_000_000, int(os.environ.get("M", 1024)), 8
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
rng = np.random.default_rng(0)
X, Z = T(rng.standard_normal((N, D))), T(rng.standard_normal((M, D)))
y, mq, yz, pm = T(rng.standard_normal(N)), T(0.3 * rng.standard_normal(N)), T(rng.standard_normal(M)), T(np.zeros(N))
ls, ll, ln = T([0.0]), T(np.full(D, np.log(1 / D))), T([0.0])
fq = L.GpFactor(M, D).factor(Z, ls, ll, 1e-5, False)
fp = L.GpFactor(M, D).factor(Z, ls, ll, 0.0, True, log_noise=ln, y_z=yz)
ts = []
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = L.fused_step(fq, fp, X, y, mq, pm, None, None, 5, 0.5)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"skip={os.environ.get('GVI_DEBUG_SKIP', '0')} M={M} ms={np.mean(ts[2:]):.3f} loss_sum={float(out[0] + out[1]):.6f}")
