"""reference: src/regularisations/projected/projected_renyi_regularisation.py"""
from .projected_regularisations import ProjectedRenyiRegularisation  # noqa: F401
