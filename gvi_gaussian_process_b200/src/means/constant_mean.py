"""ConstantMean (reference: src/means/constant_mean.py:50-66)."""
from __future__ import annotations

from ..utils.device import as_device
from .base import MeanBase, MeanBaseParameters


class ConstantMeanParameters(MeanBaseParameters):
    """constant: scalar (or [K] for multi-output)."""


class ConstantMean(MeanBase):
    Parameters = ConstantMeanParameters

    def generate_parameters(self, parameters) -> ConstantMeanParameters:
        if self.number_output_dimensions > 1:
            assert as_device(parameters["constant"]).shape[0] == self.number_output_dimensions
        return ConstantMean.Parameters(**parameters)

    def _predict(self, parameters, x):
        # jnp.tile(constant, n) (reference line 66): for K > 1 this is [c_0..c_{K-1}, c_0.., ...] and MeanBase.predict
        # then RESHAPES it to [K, n] (means/base.py:97-100) - reproduced as is, it only matters when the constants differ
        c = as_device(parameters.constant).reshape(-1)
        return c.repeat(x.shape[0]) if c.numel() > 1 else c.expand(x.shape[0]).contiguous()
