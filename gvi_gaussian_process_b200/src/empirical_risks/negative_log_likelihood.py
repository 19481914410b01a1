"""Negative log-likelihood (reference: src/empirical_risks/negative_log_likelihood.py:17-91).
Regression: mean_i [ 0.5 log(2 pi) + 0.5 log v_i + (y_i - m_i)^2 / (2 v_i) ] as one vectorised CUDA reduction
(lines 17-55).  Classification: -sum_i log sum_k p_ik y_ik (lines 57-78; a SUM over points, as in the reference)."""
from __future__ import annotations

from .. import _lib_proxy as L
from .._autograd import MultinomialRisk, PointwiseRisk, wants_grad
from ..gps.classification_base import GPClassificationBase
from ..utils.device import as_device
from .base import EmpiricalRiskBase


class NegativeLogLikelihood(EmpiricalRiskBase):
    def _calculate_multinomial_negative_log_likelihood(self, parameters, x, y):
        probabilities = self._gp.predict_probability(parameters, x).probabilities
        y = as_device(y).reshape(probabilities.shape)
        if wants_grad(probabilities):
            return MultinomialRisk.apply(probabilities, y, None, 2)[0]
        return L.multinomial_risk(probabilities, y, None, 2)[0]

    def _calculate_empirical_risk(self, parameters, x, y):
        if isinstance(self._gp, GPClassificationBase):
            return self._calculate_multinomial_negative_log_likelihood(parameters, x, y)
        gaussian = self._gp.predict_probability(parameters, x)
        mean = gaussian.mean.reshape(-1).contiguous()
        variance = gaussian.covariance.reshape(-1).contiguous()
        y = as_device(y)
        k = getattr(self._gp.kernel, "number_output_dimensions", 1)
        if k > 1 and y.ndim == 2 and y.shape[1] == k:
            y = y.T  # reference negative_log_likelihood.py:27-40 pairs y.T [K,N] with mean / covariance [K,N]
        y = y.reshape(-1).contiguous()
        assert y.numel() == mean.numel(), f"{y.numel()=} responses for {mean.numel()=} predictions"
        if wants_grad(mean, variance):
            out3 = PointwiseRisk.apply(L.REG_KINDS["none"], 0.0, y, None, None, mean, variance)
        else:
            out3 = L.risk(L.REG_KINDS["none"], 0.0, y, mean, variance, None, None)
        return out3[0] / out3[2]
