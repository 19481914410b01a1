"""Multi-GPU plumbing: one process per GPU, N sharded contiguously by rank, Z / theta replicated
(SURVEY.md section 8e).  What crosses NVLink / NVSwitch:
  - the packed loss sums [sum NLL, sum D0, N] of the fused step and the 4 + D gradient vector (one small all-reduce per
    step, summed in rank order: every rank holds the same bits),
  - one candidate record per rank per pivot of the greedy selector, written by the selector's own kernels into the
    peers' windows (gvi_cond_var_select_sharded_f64).
The data plane is the library's communicator (gvi_comm_*: NCCL + a CUDA-IPC peer window, include/gvigp.h);
torch.distributed only carries the 128-byte bootstrap id (and is what the gloo CPU tests of the host logic run on).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of n points: the first n % world_size ranks get one extra point."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Comm:
    """gvi_comm handle: rank / world of a torch.distributed group, bootstrapped through it.

    `peer_window=False` forces the NCCL transport for the selector's exchange (every rank must pass the same value)."""

    def __init__(self, group=None, peer_window: bool = True):
        from . import _lib

        self._lib = _lib
        lib = _lib.load()
        initialised = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if initialised else 0
        self.world = dist.get_world_size(group) if initialised else 1
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                _lib.check(lib.gvi_comm_unique_id_h(uid.data_ptr()), "gvi_comm_unique_id_h")
            if dist.get_backend(group) == "nccl":
                dev = uid.cuda()
                dist.broadcast(dev, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                uid = dev.cpu()
            else:
                dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        handle = ctypes.c_void_p()
        _lib.check(lib.gvi_comm_init(ctypes.byref(handle), self.rank, self.world, uid.data_ptr(),
                                     0 if peer_window else 1), "gvi_comm_init")
        self.handle = handle
        self.has_peer_window = bool(lib.gvi_comm_has_peer_window(handle))

    def allreduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks of a contiguous float64 device tensor, on the current stream."""
        assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()
        self._lib.check(self._lib.load().gvi_comm_allreduce_sum_f64(self.handle, t.data_ptr(), t.numel(),
                                                                    self._lib.current_stream()),
                        "gvi_comm_allreduce_sum_f64")
        return t

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        """[world, *t.shape] tensor holding every rank's t (contiguous float64 device tensor)."""
        assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()
        out = torch.empty((self.world, *t.shape), dtype=torch.float64, device=t.device)
        self._lib.check(self._lib.load().gvi_comm_allgather_f64(self.handle, t.data_ptr(), out.data_ptr(), t.numel(),
                                                                self._lib.current_stream()),
                        "gvi_comm_allgather_f64")
        return out

    def destroy(self):
        if self.handle is not None and self.handle.value:
            self._lib.load().gvi_comm_destroy(self.handle)
        self.handle = None

    def __del__(self):  # best effort; explicit destroy() is the orderly way (all ranks, before the process group goes)
        try:
            self.destroy()
        except Exception:
            pass


def allreduce_loss_sums(out: torch.Tensor, comm: Optional[Comm] = None, group=None) -> torch.Tensor:
    """Sum the shard-local [sum NLL, sum D0, N, ...] over ranks (in place) and return it.  With a Comm the reduction runs
    in the library (rank-order sum); without one it falls back to torch.distributed (the gloo CPU tests)."""
    if comm is not None:
        return comm.allreduce_sum(out)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def loss_from_sums(out3: torch.Tensor) -> torch.Tensor:
    """reference: mean NLL + mean D0 over ALL points (jnp.mean over the global batch)."""
    return (out3[0] + out3[1]) / out3[2]


class DistributedGVI:
    """`GeneralisedVariationalInference.calculate_loss` over points sharded across the communicator's ranks: every rank
    passes ITS slice (x_local, y_local) and receives the loss of the union, as the reference computes it for the whole
    batch (src/generalised_variational_inference.py:73-119: mean over all points)."""

    def __init__(self, gvi, comm: Comm):
        self.gvi = gvi
        self.comm = comm

    def calculate_loss(self, parameters, x_local, y_local):
        return loss_from_sums(self.comm.allreduce_sum(self.gvi.loss_sums(parameters, x_local, y_local)))

    def calculate_loss_and_gradients(self, parameters, x_local, y_local):
        return self.gvi.calculate_loss_and_gradients(parameters, x_local, y_local, allreduce=self.comm.allreduce_sum)


def calculate_gram_row_sharded(kernel, parameters, x1_local, x2, comm: Optional[Comm] = None, gather: bool = False):
    """`kernel.calculate_gram(parameters, x1, x2)` with the rows of x1 sharded by rank (reference:
    src/kernels/standard/base.py:95-103 builds the whole [m1, m2] array on one device).  Returns this rank's row block
    [m1_local, m2]; with gather=True (equal slice sizes on every rank) the replicated [m1, m2] Gram."""
    block = kernel.calculate_gram(parameters, x1_local, x2)
    if not gather:
        return block
    if comm is not None:
        return block if comm.world == 1 else comm.allgather(block.contiguous()).reshape(-1, block.shape[-1])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:  # host-logic tests over gloo
        block = block.contiguous()
        out = torch.empty((dist.get_world_size() * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(out, block)
        return out
    return block


def resolve_records(records: torch.Tensor, first: bool) -> int:
    """Host/torch restatement of the selector's winner rule over gathered candidate records [R, record_len]:
    max d; ties -> lowest global index for the first pivot (argmax), highest otherwise (reversed stable argsort).
    Used by the gloo CPU tests; the GPU path resolves on the device."""
    d, idx = records[:, 0], records[:, 1]
    best = 0
    for r in range(1, records.shape[0]):
        if d[r] > d[best] or (d[r] == d[best] and ((idx[r] < idx[best]) if first else (idx[r] > idx[best]))):
            best = r
    return best


class ShardedSelector:
    """Greedy conditional-variance selection over points sharded across ranks (reference loop:
    src/inducing_points_selection.py:137-172).  Every rank calls select() with its contiguous slice of the PERMUTED
    inputs; all ranks return the same pivot positions (global, into the permuted order) and the same tr(K - Q) trace.
    The whole M-step loop is ONE library call: no Python and no collective call between pivots."""

    def __init__(self, comm: Comm):
        self.comm = comm

    def select(self, x_local: torch.Tensor, offset: int, m_sel: int, log_scaling, log_ls, jitter: float = 1e-12):
        from . import _lib

        lib = _lib.load()
        n_local, d = x_local.shape
        dev = x_local.device
        ws = torch.empty(lib.gvi_select_sharded_workspace_bytes(n_local, m_sel, d), dtype=torch.uint8, device=dev)
        idx = torch.empty(m_sel, dtype=torch.int64, device=dev)
        trace = torch.empty(m_sel - 1, dtype=torch.float64, device=dev)
        _lib.check(lib.gvi_cond_var_select_sharded_f64(_lib.ptr(x_local) if n_local else None, n_local, offset, d, m_sel,
                                                       _lib.ptr(log_scaling), _lib.ptr(log_ls), log_ls.numel(),
                                                       float(jitter), _lib.ptr(idx), _lib.ptr(trace), _lib.ptr(ws),
                                                       self.comm.handle, _lib.current_stream()),
                   "gvi_cond_var_select_sharded_f64")
        return idx, trace
