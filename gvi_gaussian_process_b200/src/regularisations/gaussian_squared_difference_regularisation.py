"""Squared-difference regulariser on diagonal covariances (reference:
src/regularisations/gaussian_squared_difference_regularisation.py:17-86 with full_covariance=False):
mean((m_p-m_q)^2) + mean((c_p-c_q)^2)."""
from __future__ import annotations

from .base import RegularisationBase
from .schemas import RegularisationMode


class GaussianSquaredDifferenceRegularisation(RegularisationBase):
    kind = "squared_difference"

    def __init__(self, gp, regulariser, regulariser_parameters, mode=RegularisationMode.prior,
                 full_covariance: bool = True):
        if full_covariance:
            raise NotImplementedError(
                "full N x N covariances are outside the data-parallel hot path; pass full_covariance=False "
                "(SURVEY.md section 8, a11)")
        self.full_covariance = full_covariance
        super().__init__(gp=gp, regulariser=regulariser, regulariser_parameters=regulariser_parameters, mode=mode)
