import sys, numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200 import _lib
L = _lib.load(); P = _lib.ptr; chk = _lib.check
dev='cuda'
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
import os
M, D = int(os.environ.get('M', 1024)), int(os.environ.get('D', 8))
rng = np.random.default_rng(0); Z = rng.standard_normal((M, D))
F = torch.empty(L.gvi_gp_factor_bytes(M, D)//8, dtype=torch.float64, device=dev)
ws = torch.empty(L.gvi_gp_factor_workspace_bytes(M)//8, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
lsd = T([0.0]); lld = T(np.full(D, np.log(1/D))); Zd = T(Z); yz = T(rng.standard_normal(M))
for i in range(3):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    chk(L.gvi_gp_factor_f64(P(Zd), M, D, P(lsd), P(lld), D, 1e-5, 0, None, P(yz), P(F), P(ws), P(info), _lib.current_stream()), 'factor')
    e1.record(); torch.cuda.synchronize()
    print("factor ms", e0.elapsed_time(e1), "info", info.item())
