"""MultiOutputMean (reference: src/means/multi_output_mean.py:11-80): K single-output means stacked to [K, N]."""
from __future__ import annotations

import torch

from .base import MeanBase, MeanBaseParameters


class MultiOutputMeanParameters(MeanBaseParameters):
    """means: list of the member means' parameters."""


class MultiOutputMean(MeanBase):
    Parameters = MultiOutputMeanParameters

    def __init__(self, means):
        assert all(mean.number_output_dimensions == 1 for mean in means)
        self.means = list(means)
        super().__init__(number_output_dimensions=len(self.means), preprocess_function=None)

    def generate_parameters(self, parameters) -> MultiOutputMeanParameters:
        assert len(parameters["means"]) == self.number_output_dimensions
        return MultiOutputMean.Parameters(
            means=[mean._to_parameters(parameters_) for mean, parameters_ in zip(self.means, parameters["means"])])

    def _predict(self, parameters, x):
        return torch.stack([mean.predict(parameters_, x) for mean, parameters_ in zip(self.means, parameters.means)])
