"""ARDKernel on the GPU (reference: src/kernels/standard/ard_kernel.py:12-66, standard/base.py:73-103)."""
from __future__ import annotations

import torch

from ... import _lib_proxy as L
from ..._autograd import wants_grad
from ...utils.device import as_device, empty, scalar
from .base import StandardKernelBase, StandardKernelBaseParameters


class ARDKernelParameters(StandardKernelBaseParameters):
    """log_scaling: scalar; log_lengthscales: [D] (or a scalar when D == 1, README.md:92-94)."""


class ARDKernel(StandardKernelBase):
    Parameters = ARDKernelParameters

    def __init__(self, number_of_dimensions: int, preprocess_function=None):
        self.number_of_dimensions = number_of_dimensions
        super().__init__(preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> ARDKernelParameters:
        return ARDKernel.Parameters(log_scaling=parameters["log_scaling"],
                                    log_lengthscales=parameters["log_lengthscales"])

    @staticmethod
    def is_fast_path(kernel) -> bool:
        """True when `kernel` is an ARD kernel the CUDA tile kernels may evaluate from RAW inputs: no preprocess_function
        (the reference applies it inside every calculate_gram call, src/kernels/base.py:45-63)."""
        return isinstance(kernel, ARDKernel) and not kernel.has_preprocess_function

    # -- device views of the hyper-parameters (no host sync) --------------------------------------------
    @staticmethod
    def device_parameters(parameters):
        ls = scalar(parameters.log_scaling)
        ll = as_device(parameters.log_lengthscales).reshape(-1)
        return ls, ll

    def _flat(self, x: torch.Tensor) -> torch.Tensor:
        return x.reshape(x.shape[0], -1) if x.ndim > 2 else x

    def _calculate_gram(self, parameters, x1, x2):
        x1, x2 = self._flat(x1), self._flat(x2)
        ls, ll = self.device_parameters(parameters)
        d = x1.shape[1]
        assert ll.numel() in (1, d), f"log_lengthscales has {ll.numel()} entries, inputs have {d} dimensions"
        out = empty(x1.shape[0], x2.shape[0])
        L.ard_gram(x1, x2, ls, ll, out)
        return out

    def _calculate_gram_diagonal(self, parameters, x):
        ls, _ = self.device_parameters(parameters)
        if wants_grad(ls):
            return (torch.exp(ls) ** 2).expand(x.shape[0])  # k(x,x) = exp(log_scaling)^2, differentiable
        out = empty(x.shape[0])
        L.ard_gram_diag(x.shape[0], ls, out)
        return out

    def _calculate_gram_pairs(self, parameters, x1, x2):
        x1, x2 = self._flat(x1), self._flat(x2)
        ls, ll = self.device_parameters(parameters)
        out = empty(x1.shape[0])
        L.kernel_pairs(L.KERNEL_ARD, x1, x2, ls, ll, 1.0, out)
        return out
