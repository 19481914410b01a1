"""reference: src/kernels/approximate/svgp/log_svgp_kernel.py"""
from .svgp_kernels import LogSVGPKernel, LogSVGPKernelParameters  # noqa: F401
