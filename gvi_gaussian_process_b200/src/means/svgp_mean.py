"""SVGPMean (reference: src/means/svgp_mean.py:16-100): regulariser mean + K(x, x) @ weights, exactly as coded (the Gram
is taken between x and x, so `weights` has one entry per evaluated point; the inducing points are stored but unused).  The
Gram comes from the regulariser kernel's CUDA Gram routine, the matrix-vector product is a library GEMV."""
from __future__ import annotations

from ..utils.device import as_device
from .base import MeanBase, MeanBaseParameters


class SVGPMeanParameters(MeanBaseParameters):
    """weights: [n]."""


class SVGPMean(MeanBase):
    Parameters = SVGPMeanParameters

    def __init__(self, regulariser_mean_parameters, regulariser_mean, regulariser_kernel_parameters, regulariser_kernel,
                 inducing_points, number_output_dimensions: int = 1, preprocess_function=None):
        self.regulariser_mean_parameters = regulariser_mean_parameters
        self.regulariser_mean = regulariser_mean
        self.regulariser_kernel_parameters = regulariser_kernel_parameters
        self.regulariser_kernel = regulariser_kernel
        self.inducing_points = inducing_points
        super().__init__(number_output_dimensions=number_output_dimensions, preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> SVGPMeanParameters:
        return SVGPMean.Parameters(**parameters)

    def _predict(self, parameters, x):
        gram = self.regulariser_kernel.calculate_gram(parameters=self.regulariser_kernel_parameters, x1=x, x2=x,
                                                      full_covariance=True)
        product = gram @ as_device(parameters.weights)
        return (self.regulariser_mean.predict(parameters=self.regulariser_mean_parameters, x=x)
                + (product.mT if product.ndim >= 2 else product))
