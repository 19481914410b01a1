"""reference: src/kernels/approximate/svgp/kernelised_svgp_kernel.py"""
from .svgp_kernels import KernelisedSVGPKernel, KernelisedSVGPKernelParameters  # noqa: F401
