"""GWI regulariser without the eigendecomposition term (reference:
src/regularisations/gaussian_wasserstein_regularisation.py:29-54, 143-160, 230-251; default
include_eigendecomposition=False): mean((m_p-m_q)^2) + mean(c_p) + mean(c_q)."""
from __future__ import annotations

from .base import RegularisationBase
from .schemas import RegularisationMode


class GaussianWassersteinRegularisation(RegularisationBase):
    kind = "gwi_no_eig"

    def __init__(self, gp, regulariser, regulariser_parameters, mode=RegularisationMode.prior,
                 eigenvalue_regularisation: float = 1e-8, is_eigenvalue_regularisation_absolute_scale: bool = False,
                 use_symmetric_matrix_eigendecomposition: bool = True, include_eigendecomposition: bool = False):
        if include_eigendecomposition:
            raise NotImplementedError(
                "the O(N^3) eigendecomposition term is outside the data-parallel hot path (SURVEY.md section 8, a11)")
        self.eigenvalue_regularisation = eigenvalue_regularisation
        self.is_eigenvalue_regularisation_absolute_scale = is_eigenvalue_regularisation_absolute_scale
        self.use_symmetric_matrix_eigendecomposition = use_symmetric_matrix_eigendecomposition
        self.include_eigendecomposition = include_eigendecomposition
        super().__init__(gp=gp, regulariser=regulariser, regulariser_parameters=regulariser_parameters, mode=mode)
