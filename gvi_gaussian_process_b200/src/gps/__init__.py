from .base import GPBase, GPBaseParameters  # noqa: F401
from .gp_regression import GPRegression, GPRegressionParameters  # noqa: F401
from .approximate_gp_regression import ApproximateGPRegression, ApproximateGPRegressionParameters  # noqa: F401
from .gp_classification import GPClassification, GPClassificationParameters  # noqa: F401
from .approximate_gp_classification import ApproximateGPClassification, ApproximateGPClassificationParameters  # noqa: F401
