"""Gaussian / Multinomial containers (reference: src/distributions.py:14-40)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any


@dataclass
class Distribution:
    pass


@dataclass
class Gaussian(Distribution):
    mean: Any
    covariance: Any

    @property
    def full_covariance(self) -> bool:
        # reference src/distributions.py:24-29 compares with `!=`: the exact GP's quirk (mean [1,N], diagonal [N]) reads True
        return self.covariance.ndim != self.mean.ndim

    def dict(self):
        # reference: Gaussian is a ModuleParameters, so `Gaussian(**gp.predict_probability(...).dict())` works
        return {"mean": self.mean, "covariance": self.covariance}


@dataclass
class Multinomial(Distribution):
    probabilities: Any

    def dict(self):
        return {"probabilities": self.probabilities}
