"""reference: src/kernels/approximate/svgp/base.py"""
from .svgp_kernels import SVGPBaseKernel, SVGPBaseKernelParameters  # noqa: F401
