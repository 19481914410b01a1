from .sparse_posterior_kernel import SparsePosteriorKernel, SparsePosteriorKernelParameters  # noqa: F401
from .fixed_sparse_posterior_kernel import (  # noqa: F401
    FixedSparsePosteriorKernel,
    FixedSparsePosteriorKernelParameters,
)
from .svgp import (  # noqa: F401
    CholeskySVGPKernel,
    CholeskySVGPKernelParameters,
    DiagonalSVGPKernel,
    DiagonalSVGPKernelParameters,
    KernelisedSVGPKernel,
    KernelisedSVGPKernelParameters,
    LogSVGPKernel,
    LogSVGPKernelParameters,
)
