"""src.gps.base - same module layout as the reference's package (src/gps/base/{base,exact_base,approximate_base,
regression_base,classification_base}.py), so `from src.gps.base.base import GPBase, GPBaseParameters` and the other
import paths the reference's experiments, mockers and tests use keep working."""
from .base import GPBase, GPBaseParameters, LOG_ZERO  # noqa: F401
from .exact_base import ExactGPBase  # noqa: F401
from .approximate_base import ApproximateGPBase  # noqa: F401
from .regression_base import GPRegressionBase  # noqa: F401
from .classification_base import GPClassificationBase  # noqa: F401
