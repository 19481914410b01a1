"""PolynomialKernel: k(x, y) = (exp(log_scaling) <x, y> + exp(log_constant))^degree (reference:
src/kernels/non_stationary/polynomial_kernel.py:15-62).  jnp.power evaluates a Python-int degree by repeated
multiplication and a float degree by pow; the CUDA epilogue does the same for integral / non-integral degrees."""
from __future__ import annotations

from .base import NonStationaryKernelBase, NonStationaryKernelBaseParameters


class PolynomialKernelParameters(NonStationaryKernelBaseParameters):
    """log_constant, log_scaling: scalars."""


class PolynomialKernel(NonStationaryKernelBase):
    Parameters = PolynomialKernelParameters

    def __init__(self, polynomial_degree: float = 1, preprocess_function=None):
        self.polynomial_degree = polynomial_degree
        super().__init__(preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> PolynomialKernelParameters:
        return PolynomialKernel.Parameters(log_constant=parameters["log_constant"],
                                           log_scaling=parameters["log_scaling"])

    def _dot_parameters(self, parameters):
        return (self._scalar(parameters.log_scaling), self._scalar(parameters.log_constant),
                float(self.polynomial_degree))
