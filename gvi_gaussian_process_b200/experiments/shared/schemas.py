"""reference: experiments/shared/schemas.py (the enums the hot path's callers use)."""
import enum


class OptimiserSchema(str, enum.Enum):
    adam = "adam"
    adabelief = "adabelief"
    rmsprop = "rmsprop"


class EmpiricalRiskSchema(str, enum.Enum):
    negative_log_likelihood = "negative_log_likelihood"
    cross_entropy = "cross_entropy"


class RegularisationSchema(str, enum.Enum):
    projected_gaussian_wasserstein = "projected_gaussian_wasserstein"
    projected_kl = "projected_kl"
    projected_bhattacharyya = "projected_bhattacharyya"
    projected_hellinger = "projected_hellinger"
    projected_renyi = "projected_renyi"
    gaussian_squared_difference = "gaussian_squared_difference"
    gaussian_wasserstein = "gaussian_wasserstein"
    multinomial_wasserstein = "multinomial_wasserstein"


class InducingPointsSelectorSchema(str, enum.Enum):
    random = "random"
    conditional_variance = "conditional_variance"


class ActionSchema(str, enum.Enum):
    build_data = "build_data"
    train_regulariser = "train_regulariser"
    train_approximate = "train_approximate"
    temper_approximate = "temper_approximate"
