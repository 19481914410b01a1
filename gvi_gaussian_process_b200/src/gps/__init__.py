from .base import GPBase, GPBaseParameters  # noqa: F401
from .gp_regression import GPRegression  # noqa: F401
from .approximate_gp_regression import ApproximateGPRegression  # noqa: F401
from .gp_classification import GPClassification  # noqa: F401
from .approximate_gp_classification import ApproximateGPClassification  # noqa: F401
