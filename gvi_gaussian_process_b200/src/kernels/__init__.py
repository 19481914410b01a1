from .base import KernelBase, KernelBaseParameters  # noqa: F401
from .multi_output_kernel import MultiOutputKernel, MultiOutputKernelParameters  # noqa: F401
from .tempered_kernel import TemperedKernel, TemperedKernelParameters  # noqa: F401
