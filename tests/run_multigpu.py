"""Multi-GPU check, launched by torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multigpu.py
Every rank owns a contiguous shard of the points; the sharded results must equal the single-GPU results that
rank 0 computes on the full data: selector pivots bit-exact, loss and packed gradient to 1e-12."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvi_gaussian_process_b200 import distributed  # noqa: E402
from gvi_gaussian_process_b200.src import _lib_proxy as L  # noqa: E402


def check_trainer(rank, world, dev) -> bool:
    """Approximate-GP classification objective (multinomial NLL, a SUM over points -> gradient_reduction="sum") trained
    for a few AdaBelief steps on `world` ranks (each evaluating its slice of every batch) and, on rank 0, once more on a
    sub-group of size 1: the parameters must agree."""
    from gvi_gaussian_process_b200.experiments.shared import Data, OptimiserSchema, Trainer, TrainerSettings
    from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
    from gvi_gaussian_process_b200.src.gps import ApproximateGPClassification
    from gvi_gaussian_process_b200.src.kernels import MultiOutputKernel
    from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.means import CustomMean

    rng = np.random.default_rng(5)
    n, d, k, m = 1203, 3, 3, 16
    x = rng.standard_normal((n, d))
    y = np.eye(k)[rng.integers(0, k, n)]
    kern = MultiOutputKernel(kernels=[
        SparsePosteriorKernel(base_kernel=ARDKernel(number_of_dimensions=d), inducing_points=x[i * m:(i + 1) * m],
                              diagonal_regularisation=1e-4) for i in range(k)])
    # the mean returns [k, n]: MeanBase.predict RESHAPES its output to (k, -1) (reference means/base.py:97-100), so an
    # (n, k) output would mix points and classes in a batch-size dependent way and the objective would not be a sum
    # over points (true of the reference as well) - with [k, n] the reshape is the identity
    mean = CustomMean(mean_function=lambda p, xx: (torch.tanh(xx @ p[0]) @ p[1]).T, number_output_dimensions=k)
    gp = ApproximateGPClassification(mean=mean, kernel=kern)
    nll = NegativeLogLikelihood(gp=gp)
    w1, w2 = rng.standard_normal((d, 5)) / np.sqrt(d), rng.standard_normal((5, k)) / np.sqrt(5)

    def fresh_parameters():
        return gp.Parameters.construct(
            mean={"custom": (w1.copy(), w2.copy())},
            kernel={"kernels": [{"base_kernel": {"log_scaling": 0.1 * i, "log_lengthscales": np.full(d, -1.0)}}
                                for i in range(k)]})

    settings = TrainerSettings(seed=1, optimiser_schema=OptimiserSchema.adabelief, learning_rate=1e-2,
                               number_of_epochs=2, batch_size=400, batch_shuffle=True, batch_drop_last=False)
    loss_fn = lambda pd, xb, yb: nll.calculate_empirical_risk(pd, xb, yb)
    flatten = lambda p: torch.cat([torch.as_tensor(v).reshape(-1) for v in (
        list(p.mean["custom"]) + [kk["base_kernel"][name] for kk in p.kernel["kernels"]
                                  for name in ("log_scaling", "log_lengthscales")])])
    sharded, _ = Trainer(0, "", lambda p: {}, gradient_reduction="sum").train(settings, fresh_parameters(),
                                                                               Data(x=x, y=y), loss_fn)
    solo_group = dist.new_group([0])  # collective: every rank creates it, only rank 0 uses it
    ok = True
    if rank == 0:
        single, _ = Trainer(0, "", lambda p: {}, gradient_reduction="sum", group=solo_group).train(
            settings, fresh_parameters(), Data(x=x, y=y), loss_fn)
        a, b = flatten(sharded), flatten(single)
        err = float((a - b).abs().max() / b.abs().max())
        moved = float((b - flatten(fresh_parameters()).to(b.device)).abs().max())
        print(f"[multigpu] trainer: {world}-rank vs 1-rank parameters rel err {err:.2e} (moved {moved:.2e})")
        ok = err < 1e-9 and moved > 1e-3
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    rng = np.random.default_rng(0)
    n, d, m = 200_003, 8, 96
    x = rng.standard_normal((n, d))
    y = np.sin(x @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
    mq = 0.3 * rng.standard_normal(n)
    ls, ll = t([0.0]), t(np.full(d, np.log(1.0 / d)))
    lo, hi = distributed.shard_bounds(n, rank, world)
    # ---- selector: one library call per rank, the exchange inside the kernels (peer windows, then the NCCL transport) ------
    comm = distributed.Comm()
    comm_nccl = distributed.Comm(peer_window=False)
    ok = True
    if rank == 0:
        print(f"[multigpu] world={world} peer window mapped: {comm.has_peer_window}; forced NCCL transport: "
              f"{not comm_nccl.has_peer_window}")
        ref_idx, ref_trace = L.cond_var_select(t(x), m, ls, ll, 1e-12)
        ok &= not comm_nccl.has_peer_window
    for name, c in (("peer-window", comm), ("nccl", comm_nccl), ("peer-window again", comm)):
        idx, trace = distributed.ShardedSelector(c).select(t(x[lo:hi]), lo, m, ls, ll)
        gathered = c.allgather(idx.to(torch.float64))
        if rank == 0:
            same = bool(torch.equal(idx, ref_idx)) and all(bool(torch.equal(gathered[r].long(), ref_idx)) for r in range(world))
            tr_err = float((trace - ref_trace).abs().max() / ref_trace.abs().max())
            print(f"[multigpu] selector ({name}) pivots identical on every rank: {same}; trace rel err {tr_err:.2e}")
            ok &= same and tr_err < 1e-12
    # the public API: every rank passes the full inputs, works on its slice of the permuted points
    from gvi_gaussian_process_b200.src.inducing_points_selection import ConditionalVarianceInducingPointsSelector
    from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
    from gvi_gaussian_process_b200.src.utils import jax_prng

    kern = ARDKernel(number_of_dimensions=d)
    kp = kern.Parameters.construct(log_scaling=0.0, log_lengthscales=np.full(d, np.log(1.0 / d)))
    xs = t(x[:50_001])
    z_sh, i_sh = ConditionalVarianceInducingPointsSelector(comm=comm).compute_inducing_points(
        jax_prng.PRNGKey(3), xs, 64, kern, kp)
    if rank == 0:
        z_1, i_1 = ConditionalVarianceInducingPointsSelector().compute_inducing_points(jax_prng.PRNGKey(3), xs, 64, kern, kp)
        same = bool(torch.equal(i_sh, i_1) and torch.equal(z_sh, z_1))
        print(f"[multigpu] compute_inducing_points(comm=...) equals the single-GPU call: {same}")
        ok &= same
    # row-sharded calculate_gram
    blk = distributed.calculate_gram_row_sharded(kern, kp, t(x[lo:lo + 1000]), t(x[:77]), comm, gather=True)
    if rank == 0:
        full = torch.cat([kern.calculate_gram(kp, t(x[distributed.shard_bounds(n, r, world)[0]:][:1000]), t(x[:77]))
                          for r in range(world)])
        same = bool(torch.equal(blk, full))
        print(f"[multigpu] row-sharded calculate_gram (gathered) equals the per-block Grams: {same}")
        ok &= same
    # ---- fused step value + gradient -------------------------------------------------------------------------
    z = t(x[idx.cpu().numpy()])
    yz = t(y[idx.cpu().numpy()])
    fq = L.GpFactor(m, d).factor(z, ls, ll, 1e-5, False)
    fp = L.GpFactor(m, d).factor(z, ls, ll, 0.0, True, log_noise=t([0.0]), y_z=yz)
    pm = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
    out, dm = L.fused_step_grad(fq, fp, t(x[lo:hi]), t(y[lo:hi]), t(mq[lo:hi]), pm, None, None, 5, 0.5)
    comm.allreduce_sum(out)
    if rank == 0:
        pm_full = torch.zeros(n, dtype=torch.float64, device=dev)
        ref, dm_ref = L.fused_step_grad(fq, fp, t(x), t(y), t(mq), pm_full, None, None, 5, 0.5)
        rel = (out - ref).abs() / ref.abs().clamp_min(1e-300)
        err_sums, err_grad = float(rel[:3].max()), float(rel[3:].max())
        dm_err = float((dm - dm_ref[lo:hi]).abs().max() / dm_ref.abs().max())
        print(f"[multigpu] loss sums rel err vs single GPU {err_sums:.2e}; packed gradient rel err {err_grad:.2e}; "
              f"dm shard err {dm_err:.2e}; loss {float(distributed.loss_from_sums(out)):.12f}")
        # sums re-associate only (1e-12); the gradient passes through S = A^-1 C A^-1 whose rounding depends on
        # the order C was accumulated in (cond(A)^2 * eps), tolerance = the gradient parity tolerance
        ok &= err_sums < 1e-12 and err_grad < 1e-8 and dm_err < 1e-13
    # ---- data-parallel Trainer: N ranks must take the single-GPU optimisation path -----------------------------------
    ok &= check_trainer(rank, world, dev)
    comm.destroy()
    comm_nccl.destroy()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if flag.item() != 1.0:
        sys.exit(1)
    if rank == 0:
        print("[multigpu] OK")


if __name__ == "__main__":
    main()
