"""train_regulariser_gp / meta_train_regulariser_gp (reference: experiments/regression/trainers.py:19-145, without the
plotting): greedy inducing-point selection, then the exact GP on the inducing data trained on its own NLL
(value and gradient through gvi_exact_gp_predict_grad_f64)."""
from __future__ import annotations

import math
import os
from typing import Dict, List, Tuple

from ...src.gps import GPRegression
from ...src.means import ConstantMean
from ...src.utils import jax_prng
from ..shared.data import Data
from ..shared.resolvers import empirical_risk_resolver
from ..shared.trainer import Trainer, TrainerSettings
from ..shared.utils import calculate_inducing_points


def train_regulariser_gp(data: Data, empirical_risk_schema, trainer_settings: TrainerSettings, kernel,
                         kernel_parameters, inducing_points_selector, number_of_inducing_points: int,
                         empirical_risk_break_condition: float, save_checkpoint_frequency: int,
                         checkpoint_path: str) -> Tuple[GPRegression, object, List[Dict[str, float]]]:
    inducing_data = calculate_inducing_points(
        key=jax_prng.PRNGKey(trainer_settings.seed), inducing_points_selector=inducing_points_selector, data=data,
        number_of_inducing_points=number_of_inducing_points, kernel=kernel, kernel_parameters=kernel_parameters)
    gp = GPRegression(x=inducing_data.x, y=inducing_data.y, kernel=kernel, mean=ConstantMean())
    gp_parameters = gp.generate_parameters({
        "log_observation_noise": math.log(1.0), "mean": {"constant": 0.0},
        "kernel": kernel._to_parameters(kernel_parameters).dict()})
    empirical_risk = empirical_risk_resolver(empirical_risk_schema=empirical_risk_schema, gp=gp)
    risk_on_inducing = lambda parameters: empirical_risk.calculate_empirical_risk(
        parameters, inducing_data.x, inducing_data.y)
    trainer = Trainer(
        save_checkpoint_frequency=save_checkpoint_frequency, checkpoint_path=checkpoint_path,
        post_epoch_callback=lambda parameters: {"empirical-risk": risk_on_inducing(parameters)},
        break_condition_function=lambda parameters: bool(risk_on_inducing(parameters)
                                                         < empirical_risk_break_condition))
    gp_parameters, post_epoch_history = trainer.train(
        trainer_settings=trainer_settings, parameters=gp_parameters, data=inducing_data,
        loss_function=lambda parameters_dict, x, y: empirical_risk.calculate_empirical_risk(
            parameters=parameters_dict, x=x, y=y))
    return gp, gp.generate_parameters(gp_parameters.dict()), post_epoch_history


def meta_train_regulariser_gp(data: Data, empirical_risk_schema, trainer_settings: TrainerSettings, kernel,
                              kernel_parameters, inducing_points_selector, number_of_inducing_points: int,
                              number_of_iterations: int, save_checkpoint_frequency: int, checkpoint_path: str,
                              empirical_risk_break_condition: float = -float("inf")):
    """Alternates selection and training: each iteration re-selects the inducing points with the kernel parameters the
    previous one learned (reference lines 85-145; the 1-D plots are left to the caller)."""
    post_epoch_histories = []
    gp, gp_parameters = None, None
    for i in range(number_of_iterations):
        gp, gp_parameters, post_epoch_history = train_regulariser_gp(
            data=data, empirical_risk_schema=empirical_risk_schema, trainer_settings=trainer_settings, kernel=kernel,
            kernel_parameters=kernel_parameters, inducing_points_selector=inducing_points_selector,
            number_of_inducing_points=number_of_inducing_points,
            empirical_risk_break_condition=empirical_risk_break_condition,
            save_checkpoint_frequency=save_checkpoint_frequency,
            checkpoint_path=os.path.join(checkpoint_path, f"iteration-{i}"))
        kernel_parameters = gp_parameters.kernel
        post_epoch_histories.append(post_epoch_history)
    return gp, gp_parameters, post_epoch_histories
