import torch, time
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn_like(a)
for _ in range(2): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("cuBLAS DGEMM 8192^3 burst: %.2f ms %.2f TFLOP/s" % (best, 2 * 8192**3 / best * 1e-9))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): c = a @ b
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("cuBLAS DGEMM 8192^3 sustained(20): %.2f ms %.2f TFLOP/s" % (ms, 2 * 8192**3 / ms * 1e-9))
a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda"); a = a @ a.T + 4096 * torch.eye(4096, dtype=torch.float64, device="cuda")
torch.linalg.cholesky(a); torch.cuda.synchronize()
t = time.time(); L = torch.linalg.cholesky(a); torch.cuda.synchronize(); print("cusolver potrf 4096: %.2f ms" % ((time.time() - t) * 1e3))
