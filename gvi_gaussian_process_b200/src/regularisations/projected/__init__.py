from .base import ProjectedRegularisationBase  # noqa: F401
from .projected_regularisations import (  # noqa: F401
    ProjectedBhattacharyyaRegularisation,
    ProjectedGaussianWassersteinRegularisation,
    ProjectedHellingerRegularisation,
    ProjectedKLRegularisation,
    ProjectedRenyiRegularisation,
)
