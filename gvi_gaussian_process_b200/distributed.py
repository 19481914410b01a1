"""Multi-GPU plumbing: one process per GPU, N sharded contiguously by rank, Z / theta replicated
(SURVEY.md section 8e).  torch.distributed (NCCL over NVLink / NVSwitch; gloo in the CPU tests) carries only
  - the packed loss sums [sum NLL, sum D0, N] of the fused step (one 3-double all-reduce per step), and
  - one fixed-size candidate record per rank per pivot step of the greedy selector (all-gather).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of n points: the first n % world_size ranks get one extra point."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_loss_sums(out3: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the shard-local [sum NLL, sum D0, N] over ranks (in place) and return it."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out3, op=dist.ReduceOp.SUM, group=group)
    return out3


def loss_from_sums(out3: torch.Tensor) -> torch.Tensor:
    """reference: mean NLL + mean D0 over ALL points (jnp.mean over the global batch)."""
    return (out3[0] + out3[1]) / out3[2]


def resolve_records(records: torch.Tensor, first: bool) -> int:
    """Host/torch restatement of the selector's winner rule over gathered candidate records [R, record_len]:
    max d; ties -> lowest global index for the first pivot (argmax), highest otherwise (reversed stable argsort).
    Used by the gloo CPU tests; the GPU path resolves on the device (gvi_select_resolve_f64)."""
    d, idx = records[:, 0], records[:, 1]
    best = 0
    for r in range(1, records.shape[0]):
        if d[r] > d[best] or (d[r] == d[best] and ((idx[r] < idx[best]) if first else (idx[r] > idx[best]))):
            best = r
    return best


class ShardedSelector:
    """Greedy conditional-variance selection over points sharded across ranks (one exchange per pivot).

    Every rank calls select() with its contiguous slice of the PERMUTED inputs; all ranks return the same pivot
    positions (global, into the permuted order)."""

    def __init__(self, group=None):
        self.group = group

    def select(self, x_local: torch.Tensor, offset: int, m_sel: int, log_scaling, log_ls, jitter: float = 1e-12):
        from . import _lib

        lib = _lib.load()
        world = dist.get_world_size(self.group)
        n_local, d = x_local.shape
        rl = lib.gvi_select_record_len(d, m_sel)
        dev = x_local.device
        ws = torch.empty(lib.gvi_select_workspace_bytes(n_local, m_sel, d), dtype=torch.uint8, device=dev)
        record = torch.zeros(rl, dtype=torch.float64, device=dev)
        gathered = torch.zeros(world * rl, dtype=torch.float64, device=dev)  # flat: accepted by nccl and gloo
        winner = torch.zeros(rl, dtype=torch.float64, device=dev)
        idx = torch.zeros(m_sel, dtype=torch.int64, device=dev)
        trace_local = torch.zeros(m_sel - 1, dtype=torch.float64, device=dev)
        st = _lib.current_stream
        _lib.check(lib.gvi_select_init_f64(_lib.ptr(x_local), n_local, offset, d, m_sel, _lib.ptr(log_scaling),
                                           float(jitter), _lib.ptr(record), _lib.ptr(ws), st()), "select_init")
        for m in range(m_sel):
            dist.all_gather_into_tensor(gathered, record, group=self.group)
            _lib.check(lib.gvi_select_resolve_f64(_lib.ptr(gathered), world, d, m_sel, int(m == 0), _lib.ptr(winner),
                                                  idx[m:].data_ptr(), st()), "select_resolve")
            if m == m_sel - 1:
                break
            _lib.check(lib.gvi_select_update_f64(_lib.ptr(x_local), n_local, offset, d, m_sel, m, _lib.ptr(winner),
                                                 _lib.ptr(log_scaling), _lib.ptr(log_ls), log_ls.numel(),
                                                 float(jitter), _lib.ptr(record), trace_local[m:].data_ptr(),
                                                 _lib.ptr(ws), st()), "select_update")
        dist.all_reduce(trace_local, group=self.group)
        return idx, trace_local
