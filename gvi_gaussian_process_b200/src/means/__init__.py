from .base import MeanBase, MeanBaseParameters  # noqa: F401
from .constant_mean import ConstantMean, ConstantMeanParameters  # noqa: F401
from .multi_output_mean import MultiOutputMean, MultiOutputMeanParameters  # noqa: F401
from .custom_mean import CustomMean, CustomMeanParameters  # noqa: F401
from .svgp_mean import SVGPMean, SVGPMeanParameters  # noqa: F401
