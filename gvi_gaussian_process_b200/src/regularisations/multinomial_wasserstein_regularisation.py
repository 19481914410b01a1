"""MultinomialWassersteinRegularisation (reference: src/regularisations/multinomial_wasserstein_regularisation.py:10-76):
mean_i (sum_k |p_ik - q_ik|^power)^(1/power) between the class probabilities of the regulariser GP (posterior mode
only; prior mode raises NotImplementedError in the reference too) and of the approximate GP."""
from __future__ import annotations

from .. import _lib_proxy as L
from .._autograd import MultinomialRisk, wants_grad
from .base import RegularisationBase
from .schemas import RegularisationMode


class MultinomialWassersteinRegularisation(RegularisationBase):
    kind = "multinomial_wasserstein"

    def __init__(self, gp, regulariser, regulariser_parameters, mode=RegularisationMode.prior, power: int = 2):
        self.power = power
        super().__init__(gp=gp, regulariser=regulariser, regulariser_parameters=regulariser_parameters, mode=mode)

    def _calculate_regulariser_multinomial(self, x):
        if self._mode == RegularisationMode.prior:
            raise NotImplementedError
        return self._regulariser.predict_probability(self._regulariser_parameters, x)

    def _calculate_regularisation(self, parameters, x):
        p = self._calculate_regulariser_multinomial(x).probabilities.detach()
        q = self._gp.predict_probability(parameters, x).probabilities
        if wants_grad(q):
            out4 = MultinomialRisk.apply(q, None, p, self.power)
        else:
            out4 = L.multinomial_risk(q, None, p, self.power)
        return out4[1] / out4[2]
