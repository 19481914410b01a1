"""KernelBase (reference: src/kernels/base.py:38-147)."""
from __future__ import annotations

from ..module import Module, ModuleParameters
from ..utils.checks import check_matching_dimensions, check_maximum_dimension
from ..utils.device import as_2d


class KernelBaseParameters(ModuleParameters):
    pass


class KernelBase(Module):
    Parameters = KernelBaseParameters
    number_output_dimensions = 1

    def __init__(self, preprocess_function=None):
        # reference: every calculate_gram call applies the kernel's preprocess_function first (src/kernels/base.py:45-63).
        # The ARD fast paths (K tiles evaluated inside the CUDA tile kernels from raw inputs) are only taken for kernels
        # WITHOUT one (`has_preprocess_function` False); otherwise the general Gram-supplied path runs, which calls
        # calculate_gram and therefore preprocesses exactly as the reference does.
        self.has_preprocess_function = preprocess_function is not None
        self.preprocess_function = preprocess_function if preprocess_function is not None else (lambda x: x)

    def preprocess_inputs(self, x, y=None):
        # reference: src/kernels/base.py:45-63
        x1 = as_2d(self.preprocess_function(x))
        if y is None or y is x:
            return x1, x1  # one device copy: the diagonal mode recognises k(x_i, x_i)
        return x1, as_2d(self.preprocess_function(y))

    @staticmethod
    def check_inputs(x, y) -> None:
        # reference: src/kernels/base.py:65-80
        check_matching_dimensions(x, y)
        check_maximum_dimension(x, maximum_dimensionality=2)
        check_maximum_dimension(y, maximum_dimensionality=2)

    def _calculate_gram(self, parameters, x1, x2):
        raise NotImplementedError

    def _calculate_gram_diagonal(self, parameters, x):
        raise NotImplementedError

    def _calculate_gram_pairs(self, parameters, x1, x2):
        """k(x1_i, x2_i) for two DIFFERENT inputs of equal length (the reference's diagonal mode vmaps over the pairs,
        src/kernels/base.py:136-147)."""
        raise NotImplementedError(
            f"{type(self).__name__}: diagonal mode with two different inputs is off the hot path of this build")

    def calculate_gram(self, parameters, x1, x2=None, full_covariance: bool = True):
        """reference: src/kernels/base.py:104-147 (same argument order and assertion behaviour)."""
        x1, x2 = self.preprocess_inputs(x1, x2)
        self.check_inputs(x1, x2)
        parameters = self._to_parameters(parameters)
        if full_covariance:
            return self._calculate_gram(parameters, x1, x2)
        assert x1.shape[0] == x2.shape[0], (
            f"{x1.shape[0]=} must be equal to {x2.shape[0]=} for diagonal only calculation")
        if x1 is x2 or (x1.data_ptr() == x2.data_ptr() and x1.shape == x2.shape):
            return self._calculate_gram_diagonal(parameters, x1)
        return self._calculate_gram_pairs(parameters, x1, x2)
