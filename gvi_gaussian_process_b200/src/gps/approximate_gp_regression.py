"""Approximate GP regression (reference: src/gps/approximate_gp_regression.py:29-68,
src/gps/base/approximate_base.py:16,31-39).  The predictive is the "prior" of the variational mean + kernel; the
observation noise defaults to log(0) and is dropped by generate_parameters, so it contributes exp(-inf) = 0."""
from __future__ import annotations

from .base import ApproximateGPBase, GPBaseParameters, GPRegressionBase, LOG_ZERO


ApproximateGPRegressionParameters = GPBaseParameters  # the reference's name for this GP's parameter container


class ApproximateGPRegression(ApproximateGPBase, GPRegressionBase):
    def generate_parameters(self, parameters):
        return self.Parameters(
            log_observation_noise=LOG_ZERO,
            mean=self.mean._to_parameters(parameters["mean"]),
            kernel=self.kernel._to_parameters(parameters["kernel"]),
        )

    def _to_parameters(self, parameters):
        # The reference round-trips parameters through .dict() at its jit boundary and re-generates them, so the
        # observation noise of an approximate GP is dropped even when a Parameters object carries one
        # (pinned by tests/regularisations/projected/test_projected_kl.py: q = (1, 1) with log_noise = log 1).
        if isinstance(parameters, self.Parameters):
            if getattr(parameters, "_noise_dropped", False):
                return parameters  # already regenerated: idempotent, so one objective can share a predictive
            parameters = parameters.dict_shallow()
        out = self.generate_parameters(dict(parameters))
        out._noise_dropped = True
        return out

    def _calculate_prediction_gaussian(self, parameters, x, full_covariance: bool):
        return self.calculate_prior(parameters, x, full_covariance)
