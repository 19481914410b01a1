"""ApproximateGPClassification (reference: src/gps/approximate_gp_classification.py:23-76,
src/gps/base/approximate_base.py:16,31-39): the K-output variational mean + kernel, class probabilities by
quadrature.  As in regression the observation noise is dropped by generate_parameters (pinned by the reference's
tempered golden 0.1690632.. of tests/gps/test_classification.py, which passes a noise that has no effect)."""
from __future__ import annotations

from .approximate_gp_regression import ApproximateGPRegression
from .base import GPBaseParameters, GPClassificationBase


ApproximateGPClassificationParameters = GPBaseParameters  # the reference's name for this GP's parameter container


class ApproximateGPClassification(GPClassificationBase, ApproximateGPRegression):
    def __init__(self, mean, kernel, epsilon: float = 0.01, hermite_polynomial_order: int = 50,
                 cdf_lower_bound: float = 1e-10):
        ApproximateGPRegression.__init__(self, mean=mean, kernel=kernel)
        self._init_classification(mean, kernel, epsilon, hermite_polynomial_order, cdf_lower_bound)
