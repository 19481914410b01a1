"""The five projected (pGVI) regularisers.  The D0 bodies live in csrc/common.cuh (gvi_reg_term); references:
projected_bhattacharyya_regularisation.py:34-36, projected_gaussian_wasserstein_regularisation.py:34,
projected_hellinger_regularisation.py:34-45, projected_kl_regularisation.py:34-38,
projected_renyi_regularisation.py:12-45."""
from __future__ import annotations

from ..schemas import RegularisationMode
from .base import ProjectedRegularisationBase


class ProjectedBhattacharyyaRegularisation(ProjectedRegularisationBase):
    kind = "bhattacharyya"


class ProjectedGaussianWassersteinRegularisation(ProjectedRegularisationBase):
    kind = "gaussian_wasserstein"


class ProjectedHellingerRegularisation(ProjectedRegularisationBase):
    kind = "hellinger"


class ProjectedKLRegularisation(ProjectedRegularisationBase):
    kind = "kl"


class ProjectedRenyiRegularisation(ProjectedRegularisationBase):
    kind = "renyi"

    def __init__(self, gp, regulariser, regulariser_parameters, mode=RegularisationMode.prior, alpha: float = 0.5):
        self.alpha = alpha
        super().__init__(gp=gp, regulariser=regulariser, regulariser_parameters=regulariser_parameters, mode=mode)
