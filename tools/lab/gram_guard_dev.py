import sys, numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200.src import _lib_proxy as L
from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
def dev(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')
for n, m, d, lsc in [(2000, 96, 5, 1.0), (900, 1300, 8, 1.0), (900, 1300, 8, 0.1), (3000, 1024, 8, 0.03)]:
    rng = np.random.default_rng(n)
    x, z = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    ard = ARDKernel(number_of_dimensions=d)
    ll = np.full(d, np.log(1.0 / (d * lsc)))
    kp = ard.Parameters.construct(log_scaling=0.2, log_lengthscales=ll)
    sparse = SparsePosteriorKernel(base_kernel=ard, inducing_points=z)
    v_fused = sparse.calculate_gram(sparse.Parameters.construct(base_kernel=kp), x, full_covariance=False).cpu().numpy()
    for tp in (True, False):
        kzz = torch.empty(m, m, dtype=torch.float64, device='cuda'); L.ard_gram(dev(z), dev(z), dev([0.2]), dev(ll), kzz, tensor_pipe=tp)
        kxz = torch.empty(n, m, dtype=torch.float64, device='cuda'); L.ard_gram(dev(x), dev(z), dev([0.2]), dev(ll), kxz, tensor_pipe=tp)
        f = L.GpFactor(m, 1).factor_from_gram(kzz, 1e-5, False)
        kd = torch.full((n,), float(np.exp(0.4)), dtype=torch.float64, device='cuda')
        var = torch.empty(n, dtype=torch.float64, device='cuda')
        L.gp_predict_gram(f, kxz, kd, None, None, None, var)
        v = var.cpu().numpy()
        print(n, m, d, lsc, 'tensor_pipe' if tp else 'direct', 'var dev vs fused (normwise):', np.max(np.abs(v - v_fused)) / np.max(np.abs(v_fused)))
    a = torch.empty(n, m, dtype=torch.float64, device='cuda'); L.ard_gram(dev(x), dev(z), dev([0.2]), dev(ll), a, tensor_pipe=True)
    b = torch.empty(n, m, dtype=torch.float64, device='cuda'); L.ard_gram(dev(x), dev(z), dev([0.2]), dev(ll), b, tensor_pipe=False)
    print('   gram max rel dev', float(((a - b).abs() / b.abs().clamp_min(1e-300)).max()), 'identical' if torch.equal(a, b) else '')
