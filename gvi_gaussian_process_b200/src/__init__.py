"""Mirror of the reference's `src` package (hot-path subset); same module layout and import paths."""
from .generalised_variational_inference import GeneralisedVariationalInference  # noqa: F401  (reference: src/__init__.py)

__all__ = ["GeneralisedVariationalInference"]
