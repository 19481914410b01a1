"""Gaussian negative log-likelihood (reference: src/empirical_risks/negative_log_likelihood.py:17-55, 80-91):
mean_i [ 0.5 log(2 pi) + 0.5 log v_i + (y_i - m_i)^2 / (2 v_i) ] as one vectorised CUDA reduction."""
from __future__ import annotations

from .. import _lib_proxy as L
from ..utils.device import as_device
from .base import EmpiricalRiskBase


class NegativeLogLikelihood(EmpiricalRiskBase):
    def _calculate_empirical_risk(self, parameters, x, y):
        gaussian = self._gp.predict_probability(parameters, x)
        mean = gaussian.mean.reshape(-1).contiguous()
        variance = gaussian.covariance.reshape(-1).contiguous()
        y = as_device(y).reshape(-1)
        assert y.numel() == mean.numel(), f"{y.numel()=} responses for {mean.numel()=} predictions"
        out3 = L.risk(L.REG_KINDS["none"], 0.0, y, mean, variance, None, None)
        return out3[0] / out3[2]
