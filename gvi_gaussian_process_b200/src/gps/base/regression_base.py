"""GPRegressionBase (reference: src/gps/base/regression_base.py:14-43): GPs whose predict_probability returns a
Gaussian.  Common base named by the reference's metrics / plotters (experiments/regression/metrics.py:11,33)."""
from __future__ import annotations

from .base import GPBase


class GPRegressionBase(GPBase):
    pass
