"""GPBase (reference: src/gps/base/base.py:17-24, 128-151, 173-212, 246-276, 589-634)."""
from __future__ import annotations

import math

from ...distributions import Gaussian
from ...kernels.base import KernelBase
from ...means.base import MeanBase
from ...module import Module, ModuleParameters
from ...utils.device import as_2d, as_device, scalar


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
        k_out = self.kernel.number_output_dimensions
        if k_out > 1:
            # reference lines 200-211: [K, n] + atleast_1d(exp(log_observation_noise))[:, None]
            noise = as_device(parameters.log_observation_noise).reshape(-1).exp()
            if full_covariance:
                covariance = covariance.clone()
                covariance.diagonal(dim1=-2, dim2=-1).add_(noise.reshape(-1, 1))
                return covariance
            return covariance.reshape(k_out, x.shape[0]) + noise.reshape(-1, 1)
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
        mean, covariance = self._prediction_gaussian(parameters, as_2d(x), full_covariance)
        return Gaussian(mean=mean, covariance=covariance)

    def _predict_probability(self, parameters, x):
        # regression: diagonal predictive (reference regression_base.py:34-43)
        return self._prediction_gaussian(parameters, x, full_covariance=False)

    def _construct_distribution(self, probabilities):
        mean, covariance = probabilities
        return Gaussian(mean=mean, covariance=covariance)

    def predict_probability(self, parameters, x):
        """reference: src/gps/base/base.py:128-151 (Gaussian for regression, Multinomial for classification)."""
        parameters = self._to_parameters(parameters)
        return self._memoised("probability", parameters, x, lambda: self._construct_distribution(
            self._predict_probability(parameters, as_2d(x))))

    def _prediction_gaussian(self, parameters, x, full_covariance: bool):
        return self._memoised(("gaussian", full_covariance), parameters, x,
                              lambda: self._calculate_prediction_gaussian(parameters, x, full_covariance))

    def _memoised(self, tag, parameters, x, fn):
        if self._memo is None:
            return fn()
        key = (tag, id(parameters), x.data_ptr() if hasattr(x, "data_ptr") else id(x), tuple(getattr(x, "shape", ())))
        if key not in self._memo:
            self._memo[key] = (parameters, x, fn())  # keeps the keyed objects alive so that their ids stay unique
        return self._memo[key][2]

    # One objective evaluates the same predictive in the empirical risk and in the regulariser (the reference relies
    # on XLA's common-subexpression elimination inside its jitted loss); `with gp.memoise():` shares it explicitly.
    _memo = None

    def memoise(self):
        gp = self

        class _Memo:
            def __enter__(self):
                gp._memo = {}

            def __exit__(self, *exc):
                gp._memo = None

        return _Memo()


LOG_ZERO = -math.inf
