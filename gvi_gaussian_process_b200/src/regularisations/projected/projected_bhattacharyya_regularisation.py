"""reference: src/regularisations/projected/projected_bhattacharyya_regularisation.py"""
from .projected_regularisations import ProjectedBhattacharyyaRegularisation  # noqa: F401
