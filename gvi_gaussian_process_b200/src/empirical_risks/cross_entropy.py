"""CrossEntropy (reference: src/empirical_risks/cross_entropy.py:11-30).  The reference hands the class
PROBABILITIES to jax_metrics.losses.Crossentropy(), whose default treats `preds` as logits: the value is
mean_i -sum_k y_ik log_softmax(p_i)_k (pinned by the reference's own golden 1.306689,
tests/test_empirical_risks.py:215-269; plain -sum y log p would give 1.1186)."""
from __future__ import annotations

from .. import _lib_proxy as L
from .._autograd import MultinomialRisk, wants_grad
from ..gps.classification_base import GPClassificationBase
from ..utils.device import as_device
from .base import EmpiricalRiskBase


class CrossEntropy(EmpiricalRiskBase):
    def __init__(self, gp: GPClassificationBase):
        super().__init__(gp)

    def _calculate_empirical_risk(self, parameters, x, y):
        probabilities = self._gp.predict_probability(parameters, x).probabilities
        y = as_device(y).reshape(probabilities.shape)
        if wants_grad(probabilities):
            out4 = MultinomialRisk.apply(probabilities, y, None, 2)
        else:
            out4 = L.multinomial_risk(probabilities, y, None, 2)
        return out4[3] / out4[2]
