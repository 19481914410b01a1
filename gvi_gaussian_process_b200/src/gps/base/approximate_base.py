"""ApproximateGPBase (reference: src/gps/base/approximate_base.py:13-39): GPs whose predictive is the "prior" of a
variational mean + kernel, with the observation noise fixed at exp(-inf) = 0 (approximate_base.py:16).  The algebra
lives in ApproximateGPRegression / ApproximateGPClassification; this is the common base the reference's type
annotations name (experiments/shared/trainers.py:11,20)."""
from __future__ import annotations

from .base import GPBase, GPBaseParameters


class ApproximateGPBase(GPBase):
    pass


ApproximateGPBaseParameters = GPBaseParameters
