"""reference: src/kernels/approximate/svgp/diagonal_svgp_kernel.py"""
from .svgp_kernels import DiagonalSVGPKernel, DiagonalSVGPKernelParameters  # noqa: F401
