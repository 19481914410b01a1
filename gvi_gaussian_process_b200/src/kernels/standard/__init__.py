from .ard_kernel import ARDKernel, ARDKernelParameters  # noqa: F401
