"""ConformalGPRegression (reference: src/conformal_calibration/regression/conformal_gp_regression.py:11-71): Gaussian
intervals mean +- ndtri((coverage + 1) / 2) sqrt(var) of a GP's predictive, conformally calibrated."""
from __future__ import annotations

import torch

from ...utils.device import as_2d
from .base import ConformalRegressionBase


class ConformalGPRegression(ConformalRegressionBase):
    def __init__(self, x_calibration, y_calibration, gp, gp_parameters):
        self.gp = gp
        self.gp_parameters = gp_parameters
        super().__init__(x_calibration=x_calibration, y_calibration=y_calibration)

    def _predict_uncalibrated_coverage(self, x, coverage):
        gaussian = self.gp.predict_probability(parameters=self.gp_parameters, x=x)
        scale = torch.special.ndtri(torch.as_tensor((float(coverage) + 1.0) / 2.0, dtype=torch.float64))
        mean, sd = gaussian.mean.reshape(1, -1), torch.sqrt(gaussian.covariance.reshape(1, -1))
        return mean - scale * sd, mean + scale * sd

    def _predict_median(self, x):
        return as_2d(self.gp.predict_probability(parameters=self.gp_parameters, x=x).mean.reshape(1, -1))
