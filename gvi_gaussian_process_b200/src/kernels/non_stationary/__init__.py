from .base import NonStationaryKernelBase, NonStationaryKernelBaseParameters  # noqa: F401
from .inner_product_kernel import InnerProductKernel, InnerProductKernelParameters  # noqa: F401
from .polynomial_kernel import PolynomialKernel, PolynomialKernelParameters  # noqa: F401
