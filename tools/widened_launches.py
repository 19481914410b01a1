"""The widened rows of bench.py on their own (config #4 classification forward + value-and-gradient, exact-GP NLL step), for
an ncu launch list: no library GEMM may appear apart from the host framework's mean network (x @ w)."""
import sys
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
z, yz, fcn, theta = bench.make_replicated()
print("exact_gp_nll_step", bench.bench_exact_nll_step(dev, z, yz, theta))
r = bench.bench_classification(dev)
print("classification_config4", {k: r[k] for k in ("ms_per_objective", "value_and_gradient_ms", "loss")})
