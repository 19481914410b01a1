"""ProjectedRegularisationBase (reference: src/regularisations/projected/base.py:44-92): mean over points of
D0(m_p, c_p, m_q, c_q) with both Gaussians evaluated pointwise."""
from __future__ import annotations

from ..base import RegularisationBase


class ProjectedRegularisationBase(RegularisationBase):
    pass
