"""CustomMappingKernel: a feature map (e.g. a neural network of the host framework) followed by a base kernel on the
features (reference: src/kernels/custom_mapping_kernel.py:23-92).  The reference maps point by point under vmap; here the
feature map is applied to the whole batch (rows are points) and the base kernel's CUDA Gram runs on the features."""
from __future__ import annotations

from ..utils.device import as_device
from .base import KernelBase, KernelBaseParameters


class CustomMappingKernelParameters(KernelBaseParameters):
    """base_kernel: parameters of the base kernel; feature_mapping: anything the feature map understands."""


class CustomMappingKernel(KernelBase):
    Parameters = CustomMappingKernelParameters

    def __init__(self, base_kernel: KernelBase, feature_mapping, preprocess_function=None):
        self.base_kernel = base_kernel
        self.feature_mapping = feature_mapping
        super().__init__(preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> CustomMappingKernelParameters:
        return CustomMappingKernel.Parameters(
            base_kernel=self.base_kernel._to_parameters(parameters["base_kernel"]),
            feature_mapping=parameters["feature_mapping"])

    def _features(self, parameters, x):
        phi = as_device(self.feature_mapping(parameters.feature_mapping, x))
        return phi.reshape(x.shape[0], -1)

    def _calculate_gram(self, parameters, x1, x2):
        return self.base_kernel.calculate_gram(parameters.base_kernel, self._features(parameters, x1),
                                               self._features(parameters, x2))

    def _calculate_gram_pairs(self, parameters, x1, x2):
        return self.base_kernel._calculate_gram_pairs(self.base_kernel._to_parameters(parameters.base_kernel),
                                                      self._features(parameters, x1), self._features(parameters, x2))

    def _calculate_gram_diagonal(self, parameters, x):
        phi = self._features(parameters, x)
        return self.base_kernel.calculate_gram(parameters.base_kernel, phi, phi, full_covariance=False)
