"""ExactGPBase (reference: src/gps/base/exact_base.py:16-49): the family of GPs conditioned on training data (x, y).
In the reference this class holds the posterior algebra; here that algebra is the CUDA factor / predict path of
GPRegression and GPClassification, and this class is the common base the reference's type annotations and
`isinstance` checks name."""
from __future__ import annotations

from .base import GPBase


class ExactGPBase(GPBase):
    pass
