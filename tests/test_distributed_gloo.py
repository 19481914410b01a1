"""CPU-only, world_size = 2 over gloo: the N>1 host-side logic of SURVEY.md section 8(e) - contiguous sharding,
the packed loss/gradient all-reduce, and the selector's one-record-per-rank-per-step exchange with the reference
tie rules.  The per-rank numerics here come from the CPU oracle (the GPU path is covered by -m gpu tests); what is
under test is the protocol in gvi_gaussian_process_b200/distributed.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gvi_gaussian_process_b200 import distributed
from oracle import gp_oracle as O

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rng = np.random.default_rng(0)
    n, m, d = 1001, 24, 3
    x = rng.standard_normal((n, d))
    z = rng.standard_normal((m, d))
    y = np.sin(x @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
    yz = rng.standard_normal(m)
    m_q = 0.2 * rng.standard_normal(n)
    ls, ll = 0.1, np.array([0.2, -0.1, 0.3])
    return x, z, y, yz, m_q, ls, ll


def _loss_sums(x, z, y, yz, m_q, ls, ll):
    g = lambda a, b: O.ard_gram(ls, ll, a, b)
    dg = lambda a: O.ard_gram_diag(ls, ll, a)
    v_q = O.sparse_posterior_diag(g, dg, z, x)
    m_p, v_p = O.exact_gp_posterior(g, dg, np.zeros(x.shape[0]), z, yz, 0.0, x)
    n = x.shape[0]
    return np.array([O.gaussian_nll(y, m_q, v_q) * n, O.projected_regularisation("kl", m_p, v_p, m_q, v_q) * n, n])


def _worker_loss(rank, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    x, z, y, yz, m_q, ls, ll = _problem()
    lo, hi = distributed.shard_bounds(x.shape[0], rank, WORLD)
    out3 = torch.tensor(_loss_sums(x[lo:hi], z, y[lo:hi], yz, m_q[lo:hi], ls, ll))
    distributed.allreduce_loss_sums(out3)
    ret[rank] = float(distributed.loss_from_sums(out3))
    dist.destroy_process_group()


def test_sharded_loss_allreduce_matches_full_batch():
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_loss, args=(port, ret), nprocs=WORLD, join=True)
    x, z, y, yz, m_q, ls, ll = _problem()
    full = _loss_sums(x, z, y, yz, m_q, ls, ll)
    ref = (full[0] + full[1]) / full[2]
    assert ret[0] == ret[1]
    assert abs(ret[0] - ref) <= 1e-12 * abs(ref)


def _worker_select(rank, port, ret, n, m_sel, ones_kernel):
    """Runs the sharded selector protocol with NumPy standing in for the per-rank CUDA kernels: local update
    (select_update), local candidate record, all-gather, winner resolution (distributed.resolve_records)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    rng = np.random.default_rng(5)
    d = 2
    x = rng.standard_normal((n, d))
    ls, ll = 0.0, (np.full(d, -np.inf) if ones_kernel else np.array([0.3, -0.2]))
    jitter = 1e-12
    lo, hi = distributed.shard_bounds(n, rank, WORLD)
    xl = x[lo:hi]
    nl = hi - lo
    rl = 2 + d + (m_sel - 1)
    dloc = O.ard_gram_diag(ls, ll, xl) + jitter
    cloc = np.zeros((m_sel - 1, nl))
    chosen = np.zeros(nl, dtype=bool)

    def local_record(first, mrows):
        rec = np.zeros(rl)
        cand = np.where(chosen, -1.0, dloc)
        if nl == 0 or cand.max() < 0:
            rec[0], rec[1] = -1.0, -1.0
            return rec
        mx = cand.max()
        hits = np.flatnonzero(cand == mx)
        i = hits[0] if first else hits[-1]
        rec[0], rec[1] = mx, lo + i
        rec[2:2 + d] = xl[i]
        rec[2 + d:2 + d + mrows] = cloc[:mrows, i]
        return rec

    idx = []
    rec = local_record(True, 0)
    for m in range(m_sel):
        flat = torch.zeros(WORLD * rl, dtype=torch.float64)
        dist.all_gather_into_tensor(flat, torch.from_numpy(rec))
        gathered = flat.view(WORLD, rl)
        win = gathered[distributed.resolve_records(gathered, first=(m == 0))].numpy()
        idx.append(int(win[1]))
        if m == m_sel - 1:
            break
        j, dj, xj, cj = int(win[1]), np.sqrt(win[0]), win[2:2 + d], win[2 + d:2 + d + m]
        g = np.round(O.ard_gram(ls, ll, xl, xj[None, :]).reshape(-1), 20)
        if lo <= j < hi:
            g[j - lo] += jitter
            chosen[j - lo] = True
        e = (g - cj @ cloc[:m]) / dj
        cloc[m] = e
        dloc = np.clip(dloc - e * e, 0, None)
        rec = local_record(False, m + 1)
    ret[rank] = idx
    dist.destroy_process_group()


@pytest.mark.parametrize("n,m_sel,ones_kernel", [(257, 16, False), (9, 4, True), (3, 3, False)])
def test_sharded_selector_protocol_matches_single_process(n, m_sel, ones_kernel):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_select, args=(port, ret, n, m_sel, ones_kernel), nprocs=WORLD, join=True)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((n, 2))
    ls, ll = 0.0, (np.full(2, -np.inf) if ones_kernel else np.array([0.3, -0.2]))
    ref = O.conditional_variance_select(x, m_sel, lambda a, b: O.ard_gram(ls, ll, a, b),
                                        lambda a: O.ard_gram_diag(ls, ll, a))
    assert ret[0] == ret[1] == ref.tolist()


class _OracleArdKernel:
    """Stands in for the CUDA ARDKernel in the CPU test: same `calculate_gram(parameters, x1, x2)` signature."""

    def calculate_gram(self, parameters, x1, x2):
        return torch.tensor(O.ard_gram(parameters["log_scaling"], parameters["log_lengthscales"], np.asarray(x1), np.asarray(x2)))


def _worker_gram(rank, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    x, z, _, _, _, ls, ll = _problem()
    x = x[:1000]  # equal slices on both ranks (the gathered form needs them)
    lo, hi = distributed.shard_bounds(x.shape[0], rank, WORLD)
    params = {"log_scaling": ls, "log_lengthscales": ll}
    block = distributed.calculate_gram_row_sharded(_OracleArdKernel(), params, x[lo:hi], z)
    full = distributed.calculate_gram_row_sharded(_OracleArdKernel(), params, x[lo:hi], z, gather=True)
    ret[rank] = (block.numpy(), full.numpy())
    dist.destroy_process_group()


def test_row_sharded_gram_blocks_tile_the_full_gram():
    """`calculate_gram` with the rows of x1 sharded by rank - every rank's block is its rows of the
    full Gram, and the gathered form is the full Gram on every rank."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_gram, args=(port, ret), nprocs=WORLD, join=True)
    x, z, _, _, _, ls, ll = _problem()
    ref = O.ard_gram(ls, ll, x[:1000], z)
    for rank in range(WORLD):
        lo, hi = distributed.shard_bounds(1000, rank, WORLD)
        block, full = ret[rank]
        assert np.array_equal(block, ref[lo:hi]) and np.array_equal(full, ref)
