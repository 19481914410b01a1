"""EmpiricalRiskBase (reference: src/empirical_risks/base.py:13-73)."""
from __future__ import annotations

from ..gps.base import GPBase


class EmpiricalRiskBase:
    def __init__(self, gp: GPBase):
        self._gp = gp

    @property
    def gp(self) -> GPBase:
        return self._gp

    @gp.setter
    def gp(self, gp: GPBase) -> None:
        self._gp = gp

    def _calculate_empirical_risk(self, parameters, x, y):
        raise NotImplementedError

    def calculate_empirical_risk(self, parameters, x, y):
        parameters = self._gp._to_parameters(parameters)
        return self._calculate_empirical_risk(parameters, x, y)
