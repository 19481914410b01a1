"""Inducing-point selectors (reference: src/inducing_points_selection.py:15-176)."""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib_proxy as L
from .kernels.base import KernelBase
from .kernels.standard.ard_kernel import ARDKernel
from .utils import jax_prng
from .utils.device import as_device


class InducingPointsSelectorBase(ABC):
    @abstractmethod
    def compute_inducing_points(self, key, training_inputs, number_of_inducing_points, kernel=None,
                                kernel_parameters=None):
        raise NotImplementedError


class RandomInducingPointsSelector(InducingPointsSelectorBase):
    """reference lines 31-46: jax.random.choice over the row indices (with replacement)."""

    def compute_inducing_points(self, key, training_inputs, number_of_inducing_points, kernel=None,
                                kernel_parameters=None):
        x = as_device(training_inputs)
        idx = jax_prng.choice(np.asarray(key), x.shape[0], number_of_inducing_points)
        idx_d = torch.as_tensor(idx, device=x.device)
        return x[idx_d], idx_d


class ConditionalVarianceInducingPointsSelector(InducingPointsSelectorBase):
    """Greedy conditional-variance selection (reference lines 49-176) as a GPU pivoted Cholesky.

    jax.random's permutation is reproduced on the device (utils/jax_prng.permutation_device: threefry hash kernel +
    stable key sort), the permuted inputs are gathered there and the whole M-step loop runs in CUDA
    (gvi_cond_var_select_f64) with the C matrix resident in HBM.

    Multi-GPU (`comm`: a gvi_gaussian_process_b200.distributed.Comm): every rank passes the SAME full training_inputs
    (8 N D bytes - small; it is C, 8 N M bytes, that has to be sharded), works on its contiguous slice of the permuted
    points (gvi_cond_var_select_sharded_f64: one exchange per pivot inside the kernels) and receives the same
    (points, indices) as the single-GPU call."""

    def __init__(self, threshold: float = 0.0, comm=None):
        self.threshold = threshold
        self.comm = comm

    def compute_inducing_points(self, key, training_inputs, number_of_inducing_points, kernel=None,
                                kernel_parameters=None, jitter: float = 1e-12):
        assert number_of_inducing_points > 1, "Must have at least 2 inducing points"
        assert isinstance(kernel, KernelBase), f"kernel must be a kernel of this package, got {type(kernel)=}"
        x = as_device(training_inputs)
        n = x.shape[0]
        _, subkey = jax_prng.split(np.asarray(key))
        perm = jax_prng.permutation_device(subkey, n, x.device)
        x_perm = x[perm].contiguous()
        kernel_parameters = kernel._to_parameters(kernel_parameters)
        if self.comm is not None and self.comm.world > 1:
            assert ARDKernel.is_fast_path(kernel), (
                "the sharded selector evaluates the ARD kernel (no preprocess_function) inside its CUDA kernels")
            from ..distributed import ShardedSelector, shard_bounds

            ls, ll = ARDKernel.device_parameters(kernel_parameters)
            lo, hi = shard_bounds(n, self.comm.rank, self.comm.world)
            x_flat = x_perm.reshape(n, -1)
            idx, trace = ShardedSelector(self.comm).select(x_flat[lo:hi].contiguous(), lo, number_of_inducing_points,
                                                           ls, ll, jitter)
            assert int(idx.min()) >= 0, "sharded selection: the per-pivot exchange timed out (a rank never answered)"
        elif ARDKernel.is_fast_path(kernel):
            ls, ll = ARDKernel.device_parameters(kernel_parameters)
            idx, trace = L.cond_var_select(x_perm.reshape(n, -1), number_of_inducing_points, ls, ll, jitter)
        else:  # any other kernel: Gram columns from the kernel itself, the pivoted Cholesky in the same CUDA kernels
            idx, trace = L.cond_var_select_any_kernel(x_perm, number_of_inducing_points, kernel, kernel_parameters,
                                                      jitter)
        if self.threshold > 0.0:
            # reference lines 167-172: the first step m whose tr(K - Q) < threshold truncates to indices[:m]
            below = torch.nonzero(trace < self.threshold)
            if below.numel() > 0:
                m = int(below[0, 0].item())
                idx = idx[:m]
                warnings.warn("ConditionalVariance: Terminating selection of inducing points early.")
        return x_perm[idx], perm[idx]
