"""reference: src/regularisations/projected/projected_gaussian_wasserstein_regularisation.py"""
from .projected_regularisations import ProjectedGaussianWassersteinRegularisation  # noqa: F401
