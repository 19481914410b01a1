"""GPClassificationBase (reference: src/gps/base/classification_base.py:23-153).

predict_probability = Gauss-Hermite quadrature of prod_{l != j} Phi(.) over the K predictive Gaussians, then the
epsilon-smoothing of lines 70-73.  The K x K x H normal-CDF products per point run in csrc/classify.cu
(gvi_class_probabilities_f64, one warp per point); the backward pass is its CUDA vector-Jacobian product."""
from __future__ import annotations

import numpy as np

from ... import _lib_proxy as L
from ..._autograd import ClassProbabilities, wants_grad
from ...distributions import Multinomial
from ...utils.device import as_device
from .base import GPBase, GPBaseParameters


class GPClassificationBase(GPBase):
    def _init_classification(self, mean, kernel, epsilon: float, hermite_polynomial_order: int,
                             cdf_lower_bound: float):
        assert mean.number_output_dimensions > 1
        assert mean.number_output_dimensions == kernel.number_output_dimensions
        self.number_output_dimensions = mean.number_output_dimensions
        self.epsilon = epsilon
        self.cdf_lower_bound = cdf_lower_bound
        # scipy.special.roots_hermite in the reference (line 44); numpy's hermgauss is the same Golub-Welsch rule
        roots, weights = np.polynomial.hermite.hermgauss(hermite_polynomial_order)
        self._hermite_roots_h, self._hermite_weights_h = roots, weights
        self._hermite = None

    def _hermite_device(self):
        if self._hermite is None:
            self._hermite = (as_device(self._hermite_roots_h), as_device(self._hermite_weights_h))
        return self._hermite

    def _construct_distribution(self, probabilities) -> Multinomial:
        return Multinomial(probabilities=probabilities)

    def _predict_probability(self, parameters, x):
        mean, covariance = self._prediction_gaussian(parameters, x, full_covariance=False)
        k = self.number_output_dimensions
        mean = mean.reshape(k, -1).contiguous()
        covariance = covariance.reshape(k, -1).contiguous()
        roots, weights = self._hermite_device()
        if wants_grad(mean, covariance):
            return ClassProbabilities.apply(mean, covariance, roots, weights, self.epsilon, self.cdf_lower_bound)
        return L.class_probabilities(mean, covariance, roots, weights, self.epsilon, self.cdf_lower_bound)


GPClassificationBaseParameters = GPBaseParameters
