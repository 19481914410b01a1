from .base import KernelBase, KernelBaseParameters  # noqa: F401
from .multi_output_kernel import MultiOutputKernel, MultiOutputKernelParameters  # noqa: F401
from .tempered_kernel import TemperedKernel, TemperedKernelParameters  # noqa: F401
from .custom_kernel import CustomKernel, CustomKernelParameters  # noqa: F401
from .custom_mapping_kernel import CustomMappingKernel, CustomMappingKernelParameters  # noqa: F401
