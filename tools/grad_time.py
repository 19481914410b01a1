import sys, numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200 import _lib
from gvi_gaussian_process_b200.src import _lib_proxy as L
dev='cuda'
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
N, M, D = 1_000_000, 1024, 8
rng = np.random.default_rng(0)
X = T(rng.standard_normal((N, D))); Z = T(rng.standard_normal((M, D))); y = T(rng.standard_normal(N)); mq = T(0.3*rng.standard_normal(N)); yz = T(rng.standard_normal(M))
ls = T([0.0]); ll = T(np.full(D, np.log(1/D))); pm = T(np.zeros(N)); ln = T([0.0])
fq = L.GpFactor(M, D).factor(Z, ls, ll, 1e-5, False)
fp = L.GpFactor(M, D).factor(Z, ls, ll, 0.0, True, log_noise=ln, y_z=yz)
for name, fn in [("fwd", lambda: L.fused_step(fq, fp, X, y, mq, pm, None, None, 5, 0.5)), ("fwd+grad", lambda: L.fused_step_grad(fq, fp, X, y, mq, pm, None, None, 5, 0.5))]:
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): r = fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1)/3, "ms", (r[0] if isinstance(r, tuple) else r)[:6].cpu().numpy())
