"""train_approximate_gp / train_tempered_gp (reference: experiments/shared/trainers.py:15-114): the callers of the hot
path in the reference's pipeline.  Same signatures and return values; `Trainer` is the device-resident one."""
from __future__ import annotations

from typing import Dict, List, Tuple

from ...src.generalised_variational_inference import GeneralisedVariationalInference
from .data import Data
from .resolvers import empirical_risk_resolver, regularisation_resolver
from .trainer import Trainer, TrainerSettings


def train_approximate_gp(data: Data, empirical_risk_schema, regularisation_config: Dict,
                         trainer_settings: TrainerSettings, approximate_gp, approximate_gp_parameters, regulariser,
                         regulariser_parameters, save_checkpoint_frequency: int,
                         checkpoint_path: str) -> Tuple[object, List[Dict[str, float]]]:
    empirical_risk = empirical_risk_resolver(empirical_risk_schema=empirical_risk_schema, gp=approximate_gp)
    regularisation = regularisation_resolver(regularisation_config=regularisation_config, gp=approximate_gp,
                                             regulariser=regulariser, regulariser_parameters=regulariser_parameters)
    gvi = GeneralisedVariationalInference(empirical_risk=empirical_risk, regularisation=regularisation)
    trainer = Trainer(
        save_checkpoint_frequency=save_checkpoint_frequency, checkpoint_path=checkpoint_path,
        post_epoch_callback=lambda parameters: {
            "empirical-risk": empirical_risk.calculate_empirical_risk(parameters, data.x, data.y),
            "regularisation": regularisation.calculate_regularisation(parameters, data.x),
            "gvi-objective": gvi.calculate_loss(parameters, data.x, data.y),
        })
    gp_parameters, post_epoch_history = trainer.train(
        trainer_settings=trainer_settings, parameters=approximate_gp_parameters, data=data,
        loss_function=lambda parameters_dict, x, y: gvi.calculate_loss(parameters=parameters_dict, x=x, y=y))
    return approximate_gp.generate_parameters(gp_parameters.dict()), post_epoch_history


def train_tempered_gp(data: Data, empirical_risk_schema, trainer_settings: TrainerSettings, tempered_gp,
                      tempered_gp_parameters, save_checkpoint_frequency: int,
                      checkpoint_path: str) -> Tuple[object, List[Dict[str, float]]]:
    """Only the kernel's parameters (the tempering factor) are trained; mean and noise stay fixed."""
    empirical_risk = empirical_risk_resolver(empirical_risk_schema=empirical_risk_schema, gp=tempered_gp)
    tempered_gp_parameters = tempered_gp._to_parameters(tempered_gp_parameters)

    def with_kernel(kernel_parameters_dict):
        return tempered_gp.Parameters(
            log_observation_noise=tempered_gp_parameters.log_observation_noise, mean=tempered_gp_parameters.mean,
            kernel=tempered_gp_parameters.kernel.construct(**kernel_parameters_dict))

    trainer = Trainer(
        save_checkpoint_frequency=save_checkpoint_frequency, checkpoint_path=checkpoint_path,
        post_epoch_callback=lambda parameters: {
            "empirical-risk": empirical_risk.calculate_empirical_risk(
                parameters=with_kernel(parameters.dict()), x=data.x, y=data.y)})
    tempered_kernel_parameters, post_epoch_history = trainer.train(
        trainer_settings=trainer_settings, parameters=tempered_gp_parameters.kernel, data=data,
        loss_function=lambda parameters_dict, x, y: empirical_risk.calculate_empirical_risk(
            parameters=with_kernel(parameters_dict), x=x, y=y))
    return (tempered_gp.Parameters(log_observation_noise=tempered_gp_parameters.log_observation_noise,
                                   mean=tempered_gp_parameters.mean, kernel=tempered_kernel_parameters),
            post_epoch_history)
