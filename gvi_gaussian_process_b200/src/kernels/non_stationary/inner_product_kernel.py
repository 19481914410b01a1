"""InnerProductKernel: k(x, y) = exp(log_scaling) <x, y> (reference:
src/kernels/non_stationary/inner_product_kernel.py:15-50)."""
from __future__ import annotations

from .base import NonStationaryKernelBase, NonStationaryKernelBaseParameters


class InnerProductKernelParameters(NonStationaryKernelBaseParameters):
    """log_scaling: scalar."""


class InnerProductKernel(NonStationaryKernelBase):
    Parameters = InnerProductKernelParameters

    def __init__(self, preprocess_function=None):
        super().__init__(preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> InnerProductKernelParameters:
        return InnerProductKernel.Parameters(log_scaling=parameters["log_scaling"])

    def _dot_parameters(self, parameters):
        return self._scalar(parameters.log_scaling), None, 1.0
