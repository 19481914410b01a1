"""Small driver for ncu evidence of the round-2 kernels (run from the repo root under gpurun): the cooperative
factorisation, the in-tree GEMM on the gradient routes' shapes, the materialised Gram tile kernel, the sharded selector's
kernels (R = 2 virtual shards) and the exact-GP NLL step."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gvi_gaussian_process_b200 import _lib  # noqa: E402
from gvi_gaussian_process_b200.src import _lib_proxy as L  # noqa: E402

dev = "cuda"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
M, D = 1024, 8
Z, yz = T(rng.standard_normal((M, D))), T(rng.standard_normal(M))
ls, ll, ln = T([0.0]), T(np.full(D, np.log(1 / D))), T([0.0])
f = L.GpFactor(M, D)
for _ in range(2):
    f.factor(Z, ls, ll, 0.0, True, log_noise=ln, y_z=yz)                      # chol_inverse_coop_kernel, packs, matvecs
X = T(rng.standard_normal((400_000, D)))
out = torch.empty(400_000, M, dtype=torch.float64, device=dev)
for _ in range(2):
    L.ard_gram(X, Z, ls, ll, out, tensor_pipe=False)                         # ard_gram_tile_kernel<8,2>
del out
g = T(rng.standard_normal(4096))
h = T(rng.standard_normal(4096))
for _ in range(2):
    L.exact_gp_predict_grad(f, X[:4096].contiguous(), g, h, ln)               # dgemm_kernel x3 (triangular, symmetric), dgemv
xs = X[:200_000].contiguous()
L.cond_var_select_virtual_shards(xs, 96, ls, ll, 1e-12, [0, 100_000, 200_000])  # select_*_sharded_kernel
torch.cuda.synchronize()
print("done")
