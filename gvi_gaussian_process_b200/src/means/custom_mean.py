"""CustomMean (reference: src/means/custom_mean.py:20-41): wraps any callable(parameters, x) -> tensor, e.g. a
torch neural network evaluated on the device."""
from __future__ import annotations

from ..utils.device import as_device
from .base import MeanBase, MeanBaseParameters


class CustomMeanParameters(MeanBaseParameters):
    """custom: whatever the wrapped callable takes as parameters."""


class CustomMean(MeanBase):
    Parameters = CustomMeanParameters

    def __init__(self, mean_function, number_output_dimensions: int = 1, preprocess_function=None):
        self.mean_function = mean_function
        super().__init__(number_output_dimensions=number_output_dimensions, preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> CustomMeanParameters:
        return CustomMean.Parameters(custom=parameters["custom"])

    def _predict(self, parameters, x):
        return as_device(self.mean_function(parameters.custom, x))
