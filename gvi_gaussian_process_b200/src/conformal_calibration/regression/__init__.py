from .base import ConformalRegressionBase, ConformalRegressionBaseParameters  # noqa: F401
from .conformal_gp_regression import ConformalGPRegression  # noqa: F401
