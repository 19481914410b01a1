"""Driver for ncu evidence of the kernels of SURVEY.md section 8(f)-2/3 (run from the repo root under gpurun):
class-probability quadrature and its VJP, multinomial risk, variance VJP, exact-GP dense VJP."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gvi_gaussian_process_b200.src import _lib_proxy as L  # noqa: E402

dev = "cuda"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
N, K = 60_000, 10
mean, var = T(rng.standard_normal((K, N))), T(np.exp(rng.uniform(-3, 1, (K, N))))
roots, weights = (T(a) for a in np.polynomial.hermite.hermgauss(50))
for _ in range(2):
    prob = L.class_probabilities(mean, var, roots, weights, 0.01, 1e-10)            # class_prob_kernel
y = T(np.eye(K)[rng.integers(0, K, N)])
out4 = L.multinomial_risk(prob, y, prob.flip(0).contiguous(), 2)                     # multinomial_risk_kernel
dprob = L.multinomial_risk_grad(prob, y, prob.flip(0).contiguous(), 2, T([1.0, 1.0, 0.0, 0.0]))
for _ in range(2):
    dm, dv = L.class_probabilities_grad(mean, var, roots, weights, 0.01, 1e-10, dprob)  # class_prob_grad_kernel
n, m, d = 1_000_000, 1024, 8
X, Z = T(rng.standard_normal((n, d))), T(rng.standard_normal((m, d)))
ls, ll, ln, yz = T([0.0]), T(np.full(d, np.log(1 / d))), T([0.0]), T(rng.standard_normal(m))
fq = L.GpFactor(m, d).factor(Z, ls, ll, 1e-5, False)
g = T(rng.standard_normal(n))
L.gp_predict_grad(fq, X, g)                                                          # gp_step<grad> with external g
fp = L.GpFactor(m, d).factor(Z, ls, ll, 0.0, True, log_noise=ln, y_z=yz)
L.exact_gp_predict_grad(fp, Z, T(rng.standard_normal(m)), T(rng.standard_normal(m)), ln)  # exact.cu kernels
torch.cuda.synchronize()
print("done", float(out4[0]), float(dm.abs().max()))
