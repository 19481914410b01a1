"""ConformalRegressionBase (reference: src/conformal_calibration/regression/base.py:20-179): split-conformal calibration
of predictive intervals.  A consumer of the hot path: the predictives come from the CUDA kernels, the O(n) score /
quantile arithmetic stays in the host framework on the device."""
from __future__ import annotations

import numpy as np
import torch

from ...module import Module, ModuleParameters
from ...utils.device import as_2d, as_device


class ConformalRegressionBaseParameters(ModuleParameters):
    """There are no learnable parameters for conformal regression."""


class ConformalRegressionBase(Module):
    Parameters = ConformalRegressionBaseParameters

    def __init__(self, x_calibration, y_calibration):
        self.x_calibration = as_2d(x_calibration)
        self.y_calibration = as_device(y_calibration)
        self.number_of_calibration_points = self.x_calibration.shape[0]

    def _predict_uncalibrated_coverage(self, x, coverage):
        """-> (lower, upper), each [1, n]."""
        raise NotImplementedError

    def _predict_median(self, x):
        """-> [1, n]."""
        raise NotImplementedError

    def _calculate_calibration(self, coverage):
        # reference lines 76-109: score_i = max(lower_i - y_i, y_i - upper_i); the (n+1) coverage / n quantile
        lower, upper = self._predict_uncalibrated_coverage(x=self.x_calibration, coverage=coverage)
        y = self.y_calibration.reshape(1, -1)
        scores = torch.max(torch.cat([lower - y, y - upper], dim=0), dim=0).values
        n = self.number_of_calibration_points
        q = min(max((n + 1) * float(coverage) / n, 0.0), 1.0)
        return torch.quantile(scores, q)

    def _predict_coverage(self, x, coverage):
        # reference lines 111-137
        calibration = self._calculate_calibration(coverage)
        lower, upper = self._predict_uncalibrated_coverage(x=x, coverage=coverage)
        median = self._predict_median(x)
        return (torch.min(torch.cat([lower - calibration, median], dim=0), dim=0).values,
                torch.max(torch.cat([upper + calibration, median], dim=0), dim=0).values)

    def predict_coverage(self, x, coverage):
        """-> (lower, upper), each [m, n] for m coverage percentages (reference lines 139-160)."""
        x = as_2d(x)
        coverages = np.atleast_1d(coverage.detach().cpu().numpy() if isinstance(coverage, torch.Tensor)
                                  else np.asarray(coverage, dtype=np.float64)).tolist()
        pairs = [self._predict_coverage(x, c) for c in coverages]
        return torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])

    def calculate_average_interval_width(self, x, coverage):
        lower, upper = self.predict_coverage(x=x, coverage=coverage)
        return torch.mean(upper - lower, dim=1)

    def generate_parameters(self, parameters=None):
        return ConformalRegressionBaseParameters()
