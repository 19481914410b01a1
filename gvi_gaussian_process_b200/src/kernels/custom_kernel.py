"""CustomKernel: a wrapper around any host-framework kernel function, e.g. an NNGP kernel (reference:
src/kernels/custom_kernel.py:19-82).  The reference vmaps `kernel_function(custom, x1_, x2_)` over single points; here the
function is called ONCE on the batches (x1 [m1, d], x2 [m2, d]) and must return the [m1, m2] Gram as a torch tensor -
functions written for the vmapped call are batch functions already.  If the batched call does not yield an [m1, m2]
result the function is mapped over the pairs with torch.func.vmap.  The Gram itself is host-framework work; predictives
over it run on the tensor pipe through the caller-supplied-Gram entry points."""
from __future__ import annotations

import torch

from ..utils.device import as_device
from .base import KernelBase, KernelBaseParameters


class CustomKernelParameters(KernelBaseParameters):
    """custom: anything the kernel function understands."""


class CustomKernel(KernelBase):
    Parameters = CustomKernelParameters

    def __init__(self, kernel_function, preprocess_function=None):
        self.kernel_function = kernel_function
        super().__init__(preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> CustomKernelParameters:
        return CustomKernel.Parameters(custom=parameters["custom"])

    def _calculate_gram(self, parameters, x1, x2):
        m1, m2 = x1.shape[0], x2.shape[0]
        out = self.kernel_function(parameters.custom, x1, x2)
        if not (isinstance(out, torch.Tensor) and out.numel() == m1 * m2):
            pair = lambda a, b: torch.as_tensor(self.kernel_function(parameters.custom, a[None], b[None])).reshape(())
            out = torch.func.vmap(lambda a: torch.func.vmap(lambda b: pair(a, b))(x2))(x1)
        return as_device(out).reshape(m1, m2)

    def _calculate_gram_pairs(self, parameters, x1, x2):
        pair = lambda a, b: torch.as_tensor(self.kernel_function(parameters.custom, a[None], b[None])).reshape(())
        return as_device(torch.func.vmap(pair)(x1, x2)).reshape(-1)

    def _calculate_gram_diagonal(self, parameters, x):
        return self._calculate_gram_pairs(parameters, x, x)
