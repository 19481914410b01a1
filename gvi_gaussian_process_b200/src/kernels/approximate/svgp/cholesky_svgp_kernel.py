"""reference: src/kernels/approximate/svgp/cholesky_svgp_kernel.py"""
from .svgp_kernels import CholeskySVGPKernel, CholeskySVGPKernelParameters  # noqa: F401
