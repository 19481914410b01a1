"""MeanBase (reference: src/means/base.py:78-100).  Mean functions stay in the host framework (torch on device)."""
from __future__ import annotations

from ..module import Module, ModuleParameters
from ..utils.device import as_2d


class MeanBaseParameters(ModuleParameters):
    pass


class MeanBase(Module):
    Parameters = MeanBaseParameters

    def __init__(self, number_output_dimensions: int = 1, preprocess_function=None):
        self.number_output_dimensions = number_output_dimensions
        self.preprocess_function = preprocess_function if preprocess_function is not None else (lambda x: x)

    def preprocess_input(self, x):
        return as_2d(self.preprocess_function(x))

    def _predict(self, parameters, x):
        raise NotImplementedError

    def predict(self, parameters, x):
        x = self.preprocess_input(x)
        parameters = self._to_parameters(parameters)
        out = self._predict(parameters=parameters, x=x)
        if self.number_output_dimensions == 1:
            return out.reshape(-1)
        return out.reshape(self.number_output_dimensions, -1)
