"""Sharded selector timing under torchrun (one rank per GPU): N_PER_GPU points per rank, M pivots."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from gvi_gaussian_process_b200 import distributed  # noqa: E402
from gvi_gaussian_process_b200.src import _lib_proxy as L  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_per, d, m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 8, int(sys.argv[2]) if len(sys.argv) > 2 else 512
rng = np.random.default_rng(rank)
x = torch.as_tensor(rng.standard_normal((n_per, d)), device=dev)
ls = torch.zeros(1, dtype=torch.float64, device=dev)
ll = torch.full((d,), float(np.log(1.0 / d)), dtype=torch.float64, device=dev)
for it in range(2):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if world > 1:
        idx, tr = distributed.ShardedSelector().select(x, rank * n_per, m, ls, ll)
    else:
        idx, tr = L.cond_var_select(x, m, ls, ll, 1e-12)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0 and it == 1:
        ms = float(t.item())
        print(f"selector world={world} N={n_per * world} M={m}: {ms:.1f} ms  {n_per * world * m / ms * 1e3:.3e} points*M/s  "
              f"{(4.0 * n_per * world * m * m + 32.0 * n_per * world * m) / ms * 1e-6 / world:.0f} GB/s per GPU")
if world > 1:
    dist.destroy_process_group()
