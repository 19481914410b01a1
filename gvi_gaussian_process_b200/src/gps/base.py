"""GPBase (reference: src/gps/base/base.py:17-24, 128-151, 173-212, 246-276, 589-634)."""
from __future__ import annotations

import math

from ..distributions import Gaussian
from ..kernels.base import KernelBase
from ..means.base import MeanBase
from ..module import Module, ModuleParameters
from ..utils.device import as_2d, scalar


class GPBaseParameters(ModuleParameters):
    """log_observation_noise, mean, kernel."""


class GPBase(Module):
    Parameters = GPBaseParameters

    def __init__(self, mean: MeanBase, kernel: KernelBase):
        self.mean = mean
        self.kernel = kernel

    # -- parameters -----------------------------------------------------------------------------------------
    def generate_parameters(self, parameters) -> GPBaseParameters:
        return self.Parameters(
            log_observation_noise=parameters["log_observation_noise"],
            mean=self.mean._to_parameters(parameters["mean"]),
            kernel=self.kernel._to_parameters(parameters["kernel"]),
        )

    # -- prior (reference base.py:173-212, 246-276) -----------------------------------------------------------
    def calculate_prior_covariance(self, parameters, x, full_covariance: bool):
        parameters = self._to_parameters(parameters)
        x = as_2d(x)
        covariance = self.kernel.calculate_gram(parameters.kernel, x, x, full_covariance=full_covariance)
        noise = scalar(parameters.log_observation_noise).exp()
        if full_covariance:
            covariance = covariance.clone()
            covariance.diagonal().add_(noise)
            return covariance
        return covariance + noise

    def calculate_prior(self, parameters, x, full_covariance: bool):
        parameters = self._to_parameters(parameters)
        x = as_2d(x)
        mean = self.mean.predict(parameters.mean, x)
        covariance = self.calculate_prior_covariance(parameters, x, full_covariance)
        return mean, covariance

    # -- prediction -------------------------------------------------------------------------------------------
    def _calculate_prediction_gaussian(self, parameters, x, full_covariance: bool):
        raise NotImplementedError

    def calculate_prediction_gaussian(self, parameters, x, full_covariance: bool) -> Gaussian:
        parameters = self._to_parameters(parameters)
        mean, covariance = self._calculate_prediction_gaussian(parameters, as_2d(x), full_covariance)
        return Gaussian(mean=mean, covariance=covariance)

    def _predict_probability(self, parameters, x):
        # regression: diagonal predictive (reference regression_base.py:34-43)
        return self._calculate_prediction_gaussian(parameters, x, full_covariance=False)

    def predict_probability(self, parameters, x) -> Gaussian:
        """reference: src/gps/base/base.py:128-151."""
        parameters = self._to_parameters(parameters)
        mean, covariance = self._predict_probability(parameters, as_2d(x))
        return Gaussian(mean=mean, covariance=covariance)


LOG_ZERO = -math.inf
