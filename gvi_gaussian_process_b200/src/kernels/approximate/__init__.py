from .sparse_posterior_kernel import SparsePosteriorKernel, SparsePosteriorKernelParameters  # noqa: F401
from .fixed_sparse_posterior_kernel import (  # noqa: F401
    FixedSparsePosteriorKernel,
    FixedSparsePosteriorKernelParameters,
)
from .svgp import CholeskySVGPKernel, DiagonalSVGPKernel, KernelisedSVGPKernel, LogSVGPKernel  # noqa: F401
