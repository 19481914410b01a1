"""StandardKernelBase (reference: src/kernels/standard/base.py:13-103): kernels defined by a scalar k(x, z), whose Gram
the reference builds with vmap(vmap(.)).  Here the Gram is a CUDA kernel of the subclass; this is the common base."""
from __future__ import annotations

from ..base import KernelBase, KernelBaseParameters


class StandardKernelBaseParameters(KernelBaseParameters):
    pass


class StandardKernelBase(KernelBase):
    pass
