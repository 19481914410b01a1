"""RegularisationBase (reference: src/regularisations/base.py:15-190).

Every regulariser on the hot path is a mean over points of a pointwise function of (m_p, c_p, m_q, c_q); the
CUDA side therefore needs only a `kind` code (include/gvigp.h GVI_REG_*) and an optional alpha."""
from __future__ import annotations

from .. import _lib_proxy as L
from .._autograd import PointwiseRisk, wants_grad
from ..distributions import Gaussian
from ..gps.base import GPBase
from .schemas import RegularisationMode


class RegularisationBase:
    kind: str = "none"
    alpha: float = 0.0

    def __init__(self, gp: GPBase, regulariser: GPBase, regulariser_parameters, mode=RegularisationMode.prior):
        self._gp = gp
        self._regulariser = regulariser
        self._regulariser_parameters = regulariser_parameters
        self._mode = RegularisationMode(mode)

    # the reference re-jits on every setter (base.py:42-99); nothing to rebuild here
    @property
    def gp(self):
        return self._gp

    @gp.setter
    def gp(self, gp):
        self._gp = gp

    @property
    def regulariser(self):
        return self._regulariser

    @regulariser.setter
    def regulariser(self, regulariser):
        self._regulariser = regulariser

    @property
    def regulariser_parameters(self):
        return self._regulariser_parameters

    @regulariser_parameters.setter
    def regulariser_parameters(self, regulariser_parameters):
        self._regulariser_parameters = regulariser_parameters

    @property
    def mode(self):
        return self._mode

    @mode.setter
    def mode(self, mode):
        self._mode = RegularisationMode(mode)

    def _calculate_regulariser_gaussian(self, x, full_covariance: bool) -> Gaussian:
        """reference base.py:109-139: prior -> regulariser.calculate_prior, posterior -> prediction gaussian."""
        if self._mode == RegularisationMode.prior:
            mean_p, covariance_p = self._regulariser.calculate_prior(
                parameters=self._regulariser_parameters, x=x, full_covariance=full_covariance)
            return Gaussian(mean=mean_p, covariance=covariance_p)
        return self._regulariser.calculate_prediction_gaussian(
            parameters=self._regulariser_parameters, x=x, full_covariance=full_covariance)

    def _pointwise_sums(self, parameters, x):
        """out3 = [-, sum_i D0_i, N] through the CUDA reduction.  K-output GPs give [K, N] arrays; with equal N per
        output the reference's mean over outputs of per-output means (projected/base.py:75-92) is the mean over K*N."""
        gaussian_p = self._calculate_regulariser_gaussian(x, full_covariance=False)
        gaussian_q = self._gp.calculate_prediction_gaussian(parameters, x, full_covariance=False)
        m_p = gaussian_p.mean.reshape(-1).contiguous()
        c_p = gaussian_p.covariance.reshape(-1).contiguous()
        m_q = gaussian_q.mean.reshape(-1).contiguous()
        c_q = gaussian_q.covariance.reshape(-1).contiguous()
        if wants_grad(m_q, c_q):
            return PointwiseRisk.apply(L.REG_KINDS[self.kind], self.alpha, None, m_p.detach(), c_p.detach(), m_q, c_q)
        return L.risk(L.REG_KINDS[self.kind], self.alpha, None, m_q, c_q, m_p, c_p)

    def _calculate_regularisation(self, parameters, x):
        out3 = self._pointwise_sums(parameters, x)
        return out3[1] / out3[2]

    def calculate_regularisation(self, parameters, x):
        """reference: src/regularisations/base.py:169-190."""
        parameters = self._gp._to_parameters(parameters)
        return self._calculate_regularisation(parameters, x)
