from .data import Data  # noqa: F401
from .schemas import OptimiserSchema  # noqa: F401
from .trainer import Trainer, TrainerSettings  # noqa: F401
from .schemas import (ActionSchema, EmpiricalRiskSchema, InducingPointsSelectorSchema,  # noqa: F401
                      RegularisationSchema)
from .resolvers import (empirical_risk_resolver, inducing_points_selector_resolver,  # noqa: F401
                        regularisation_resolver)
from .trainers import train_approximate_gp, train_tempered_gp  # noqa: F401
from .utils import calculate_inducing_points  # noqa: F401
