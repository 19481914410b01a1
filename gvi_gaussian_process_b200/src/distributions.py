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
        return self.covariance.ndim > self.mean.ndim


@dataclass
class Multinomial(Distribution):
    probabilities: Any

    def dict(self):
        return {"probabilities": self.probabilities}
