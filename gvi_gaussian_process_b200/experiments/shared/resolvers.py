"""Resolvers of the objective's parts (reference: experiments/shared/resolvers/empirical_risk.py:9-19,
regularisation.py:23-103, inducing_points_selector.py:5-15): schema value -> object of the src-compatible shim.  The
kernel / mean / neural-network resolvers of the reference build flax and neural-tangents objects from YAML and are not
part of the hot path; callers construct kernels and means directly."""
from __future__ import annotations

from ...src.empirical_risks import CrossEntropy, NegativeLogLikelihood
from ...src.gps.base.classification_base import GPClassificationBase
from ...src.inducing_points_selection import (ConditionalVarianceInducingPointsSelector,
                                               RandomInducingPointsSelector)
from ...src.regularisations import (GaussianSquaredDifferenceRegularisation, GaussianWassersteinRegularisation,
                                    MultinomialWassersteinRegularisation)
from ...src.regularisations import projected
from . import schemas


def empirical_risk_resolver(empirical_risk_schema, gp):
    empirical_risk_schema = schemas.EmpiricalRiskSchema(empirical_risk_schema)
    if empirical_risk_schema == schemas.EmpiricalRiskSchema.negative_log_likelihood:
        return NegativeLogLikelihood(gp=gp)
    assert isinstance(gp, GPClassificationBase), "CrossEntropy is only for classification"
    return CrossEntropy(gp=gp)


_REGULARISATIONS = {
    schemas.RegularisationSchema.gaussian_squared_difference: GaussianSquaredDifferenceRegularisation,
    schemas.RegularisationSchema.gaussian_wasserstein: GaussianWassersteinRegularisation,
    schemas.RegularisationSchema.projected_gaussian_wasserstein: projected.ProjectedGaussianWassersteinRegularisation,
    schemas.RegularisationSchema.projected_kl: projected.ProjectedKLRegularisation,
    schemas.RegularisationSchema.projected_bhattacharyya: projected.ProjectedBhattacharyyaRegularisation,
    schemas.RegularisationSchema.projected_hellinger: projected.ProjectedHellingerRegularisation,
    schemas.RegularisationSchema.projected_renyi: projected.ProjectedRenyiRegularisation,
    schemas.RegularisationSchema.multinomial_wasserstein: MultinomialWassersteinRegularisation,
}


def regularisation_resolver(regularisation_config, gp, regulariser, regulariser_parameters):
    assert "regularisation_schema" in regularisation_config, "Regularisation schema must be specified"
    assert "regularisation_kwargs" in regularisation_config, "Regularisation kwargs must be specified"
    schema = schemas.RegularisationSchema(regularisation_config["regularisation_schema"])
    if schema == schemas.RegularisationSchema.multinomial_wasserstein:
        assert isinstance(gp, GPClassificationBase), "GP must be a classification GP"
        assert isinstance(regulariser, GPClassificationBase), "Regulariser must be a classification GP"
    return _REGULARISATIONS[schema](gp=gp, regulariser=regulariser, regulariser_parameters=regulariser_parameters,
                                    **dict(regularisation_config["regularisation_kwargs"]))


def inducing_points_selector_resolver(inducing_points_schema):
    schema = schemas.InducingPointsSelectorSchema(inducing_points_schema)
    if schema == schemas.InducingPointsSelectorSchema.random:
        return RandomInducingPointsSelector()
    return ConditionalVarianceInducingPointsSelector()
