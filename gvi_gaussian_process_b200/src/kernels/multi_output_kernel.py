"""MultiOutputKernel (reference: src/kernels/multi_output_kernel.py:11-70): K single-output kernels evaluated on the same
inputs, stacked to [K, m1, m2] (or [K, m1] in diagonal mode, src/kernels/base.py:132-147).  Each member runs on its
own CUDA state (its own inducing set and factorisation when it is a sparse posterior kernel)."""
from __future__ import annotations

import torch

from .. import _lib_proxy as L
from .base import KernelBase, KernelBaseParameters


class MultiOutputKernelParameters(KernelBaseParameters):
    """kernels: list of the member kernels' parameters."""


class MultiOutputKernel(KernelBase):
    Parameters = MultiOutputKernelParameters

    def __init__(self, kernels):
        assert all(kernel.number_output_dimensions == 1 for kernel in kernels)
        self.kernels = list(kernels)
        super().__init__(preprocess_function=None)
        self.number_output_dimensions = len(self.kernels)

    def generate_parameters(self, parameters) -> MultiOutputKernelParameters:
        assert len(parameters["kernels"]) == self.number_output_dimensions
        return MultiOutputKernel.Parameters(
            kernels=[kernel._to_parameters(parameters_)
                     for kernel, parameters_ in zip(self.kernels, parameters["kernels"])])

    def _calculate_gram(self, parameters, x1, x2):
        return torch.stack([kernel.calculate_gram(parameters_, x1, x2)
                            for kernel, parameters_ in zip(self.kernels, parameters.kernels)])

    def _calculate_gram_diagonal(self, parameters, x):
        # the K members are independent: their factorisations and predictive launches overlap on side streams
        jobs = [(lambda kernel=kernel, parameters_=parameters_: kernel.calculate_gram(parameters_, x, x,
                                                                                    full_covariance=False))
                for kernel, parameters_ in zip(self.kernels, parameters.kernels)]
        return torch.stack(L.run_concurrently(jobs))
