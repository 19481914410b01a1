"""Gaussian container (reference: src/distributions.py:14-29)."""
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
