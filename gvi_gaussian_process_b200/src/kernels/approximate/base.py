"""ApproximateBaseKernel (reference: src/kernels/approximate/base.py:9-60): common base of the kernels defined through a
set of inducing points.  The sparse-posterior algebra itself lives in the subclasses (CUDA factor / predict path)."""
from __future__ import annotations

from ..base import KernelBase, KernelBaseParameters


class ApproximateBaseKernelParameters(KernelBaseParameters):
    pass


class ApproximateBaseKernel(KernelBase):
    pass
