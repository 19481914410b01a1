"""Sharded selector timing (one rank per GPU under torchrun, or a single process): N_PER_GPU points per rank, M pivots.
    python tools/selector_scaling.py [n_per_gpu] [m]
Times, per transport, the one-call sharded selector (gvi_cond_var_select_sharded_f64: exchange inside the kernels over
the peer windows / over ncclAllGather enqueued from C) and, at world size 1, the single-shard call beside it."""
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
comms = [("peer-window", distributed.Comm())]
if world > 1:
    comms.append(("nccl-allgather", distributed.Comm(peer_window=False)))


def timed(fn):
    best = None
    for it in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            best = float(t.item()) if best is None else min(best, float(t.item()))
    return best, out


def report(name, ms):
    if rank == 0:
        by = 4.0 * n_per * world * m * m + 32.0 * n_per * world * m
        print(f"selector[{name}] world={world} N={n_per * world} M={m}: {ms:.2f} ms  {n_per * world * m / ms * 1e3:.3e} "
              f"points*M/s  {by / ms * 1e-6 / world:.0f} GB/s per GPU", flush=True)


ref = None
if world == 1:
    ms, (ref, _) = timed(lambda: L.cond_var_select(x, m, ls, ll, 1e-12))
    report("single-shard call", ms)
for name, c in comms:
    ms, (idx, _) = timed(lambda: distributed.ShardedSelector(c).select(x, rank * n_per, m, ls, ll))
    report(f"sharded, {name}{'' if c.has_peer_window or 'nccl' in name else ' (NOT mapped: fell back to nccl)'}", ms)
    if ref is not None:
        assert torch.equal(idx, ref)
    assert int(idx.min()) >= 0 and len(set(idx.tolist())) == m
for _, c in comms:
    c.destroy()
if world > 1:
    dist.destroy_process_group()
