"""reference: src/regularisations/projected/projected_kl_regularisation.py"""
from .projected_regularisations import ProjectedKLRegularisation  # noqa: F401
