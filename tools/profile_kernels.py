"""Small driver for ncu evidence of the non-headline kernels (run from the repo root under gpurun):
selector update, weighted-Gram (SYRK) + gradient tile kernel, tensor-pipe Gram, risk reduction, factorisation."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gvi_gaussian_process_b200.src import _lib_proxy as L  # noqa: E402

dev = "cuda"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
N, M, D = 1_000_000, 1024, 8
X, Z = T(rng.standard_normal((N, D))), T(rng.standard_normal((M, D)))
y, mq, yz = T(rng.standard_normal(N)), T(0.3 * rng.standard_normal(N)), T(rng.standard_normal(M))
ls, ll, ln, pm = T([0.0]), T(np.full(D, np.log(1 / D))), T([0.0]), T(np.zeros(N))
fq = L.GpFactor(M, D).factor(Z, ls, ll, 1e-5, False)
fp = L.GpFactor(M, D).factor(Z, ls, ll, 0.0, True, log_noise=ln, y_z=yz)
out, dm = L.fused_step_grad(fq, fp, X, y, mq, pm, None, None, 5, 0.5)          # gp_step<grad>, kgen, syrk, ...
_, vq = L.gp_predict(fq, X, want_mean=False)
mp, vp = L.gp_predict(fp, X, prior_mean=pm)
L.risk(5, 0.5, y, mq, vq, mp, vp)                                              # risk_kernel
L.cond_var_select(X, 160, ls, ll, 1e-12)                                       # select_update_kernel
Xw, Zw = T(rng.uniform(0, 1, (60000, 784))), T(rng.uniform(0, 1, (2048, 784)))
gout = torch.empty(60000, 2048, dtype=torch.float64, device=dev)
L.ard_gram(Xw, Zw, ls, T(np.full(784, np.log(1 / 784))), gout)                  # gram_dmma_kernel
torch.cuda.synchronize()
print("done", out[:4].cpu().numpy())
