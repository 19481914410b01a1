"""TemperedKernel (reference: src/kernels/tempered_kernel.py:14-81): exp(log_tempering_factor) * base Gram with the
base kernel's parameters frozen - the calibration stage of the reference's pipeline.  The base Gram / predictive
variance comes from the CUDA path; the scalar factor is applied to its O(N) output."""
from __future__ import annotations

from ..utils.device import as_device, scalar
from .base import KernelBase, KernelBaseParameters


class TemperedKernelParameters(KernelBaseParameters):
    """log_tempering_factor: scalar."""


class TemperedKernel(KernelBase):
    Parameters = TemperedKernelParameters

    def __init__(self, base_kernel: KernelBase, base_kernel_parameters, number_output_dimensions: int = 1,
                 preprocess_function=None):
        self.base_kernel = base_kernel
        self.base_kernel_parameters = base_kernel._to_parameters(base_kernel_parameters)
        super().__init__(preprocess_function=preprocess_function)
        self.number_output_dimensions = number_output_dimensions

    def generate_parameters(self, parameters) -> TemperedKernelParameters:
        return TemperedKernelParameters(log_tempering_factor=parameters["log_tempering_factor"])

    def _factor(self, parameters, ndim):
        # reference lines 76-80: atleast_1d(exp(log_tempering_factor))[:, None, None] * atleast_3d(gram)
        if self.number_output_dimensions == 1:
            return scalar(parameters.log_tempering_factor).exp()
        t = as_device(parameters.log_tempering_factor).reshape(-1).exp()
        return t.reshape((-1,) + (1,) * (ndim - 1))

    def _calculate_gram(self, parameters, x1, x2):
        gram = self.base_kernel.calculate_gram(self.base_kernel_parameters, x1, x2)
        return self._factor(parameters, gram.ndim) * gram

    def _calculate_gram_diagonal(self, parameters, x):
        gram = self.base_kernel.calculate_gram(self.base_kernel_parameters, x, x, full_covariance=False)
        return self._factor(parameters, gram.ndim) * gram
