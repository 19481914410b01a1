from .data import Data  # noqa: F401
from .schemas import OptimiserSchema  # noqa: F401
from .trainer import Trainer, TrainerSettings  # noqa: F401
