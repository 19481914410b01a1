from .base import KernelBase, KernelBaseParameters  # noqa: F401
