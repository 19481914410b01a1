"""Earlier location of GPClassificationBase; it now lives where the reference keeps it (src/gps/base/classification_base.py)."""
from .base.classification_base import GPClassificationBase  # noqa: F401
