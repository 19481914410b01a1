"""reference: src/regularisations/projected/projected_hellinger_regularisation.py"""
from .projected_regularisations import ProjectedHellingerRegularisation  # noqa: F401
