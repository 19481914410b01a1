from .base import RegularisationBase  # noqa: F401
from .schemas import RegularisationMode  # noqa: F401
from .gaussian_wasserstein_regularisation import GaussianWassersteinRegularisation  # noqa: F401
from .multinomial_wasserstein_regularisation import MultinomialWassersteinRegularisation  # noqa: F401
from .gaussian_squared_difference_regularisation import GaussianSquaredDifferenceRegularisation  # noqa: F401
