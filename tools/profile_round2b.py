"""Driver for ncu evidence of the kernels of the last session of round 2 (run from the repo root under gpurun): the guarded
tensor-pipe Gram for small D (gram_guard_kernel + ard_gram_mma_kernel), the risk reduction with the table-driven log at
N = 1e8, the quarter-tile GEMM and the 8-point forward tiles of the exact-GP NLL step at M = N = 1024, and one chunk of the
gradient step at config #3 (gradient-mode tile kernel + syrk_kernel)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gvi_gaussian_process_b200.src import _lib_proxy as L  # noqa: E402

dev = "cuda"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
M, D = 1024, 8
Z, yz = T(rng.standard_normal((M, D))), T(rng.standard_normal(M))
ls, ll, ln = T([0.0]), T(np.full(D, np.log(1 / D))), T([0.0])
X = T(rng.standard_normal((400_000, D)))
out = torch.empty(400_000, M, dtype=torch.float64, device=dev)
for _ in range(2):
    L.ard_gram(X, Z, ls, ll, out)                                               # gram_guard_kernel x2, ard_gram_mma_kernel<1,1,4>
del out
n = 100_000_000
y, mq, vq, mp, vp = [torch.rand(n, dtype=torch.float64, device=dev) + 0.5 for _ in range(5)]
for _ in range(2):
    L.risk(L.REG_KINDS["renyi"], 0.5, y, mq, vq, mp, vp)                        # risk_kernel (N = 1e8)
del y, mq, vq, mp, vp
torch.cuda.empty_cache()
f = L.GpFactor(M, D)
for _ in range(2):
    f.factor(Z, ls, ll, 0.0, True, log_noise=ln, y_z=yz)
xs = X[:1024].contiguous()
g, h = T(rng.standard_normal(1024)), T(rng.standard_normal(1024))
for _ in range(2):
    L.gp_predict(f, xs)                                                         # gp_step_kernel<1,8,0,16>: 8-point tiles
    L.exact_gp_predict_grad(f, xs, g, h, ln)                                    # dgemm_kernel<16,..> x3: 64 x 64 tiles
N = 1_000_000
Xl = T(rng.standard_normal((N, D)))
yl, ml, pm = T(rng.standard_normal(N)), T(0.3 * rng.standard_normal(N)), T(np.zeros(N))
fq = L.GpFactor(M, D).factor(Z, ls, ll, 1e-5, False)
fp = L.GpFactor(M, D).factor(Z, ls, ll, 0.0, True, log_noise=ln, y_z=yz)
for _ in range(2):
    o, dm = L.fused_step_grad(fq, fp, Xl, yl, ml, pm, None, None, 5, 0.5)       # gp_step_kernel<3,8,1,16>, syrk_kernel
torch.cuda.synchronize()
print("done", o[:4].cpu().numpy())
