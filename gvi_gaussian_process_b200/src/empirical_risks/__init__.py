from .base import EmpiricalRiskBase  # noqa: F401
from .negative_log_likelihood import NegativeLogLikelihood  # noqa: F401
from .cross_entropy import CrossEntropy  # noqa: F401
