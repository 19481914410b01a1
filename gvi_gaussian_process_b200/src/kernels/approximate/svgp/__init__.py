from .svgp_kernels import (  # noqa: F401
    CholeskySVGPKernel,
    CholeskySVGPKernelParameters,
    DiagonalSVGPKernel,
    DiagonalSVGPKernelParameters,
    KernelisedSVGPKernel,
    KernelisedSVGPKernelParameters,
    LogSVGPKernel,
    LogSVGPKernelParameters,
    SVGPBaseKernel,
    SVGPBaseKernelParameters,
)
