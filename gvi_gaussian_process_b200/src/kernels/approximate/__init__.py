from .sparse_posterior_kernel import SparsePosteriorKernel, SparsePosteriorKernelParameters  # noqa: F401
