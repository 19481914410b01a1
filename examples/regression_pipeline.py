"""The reference's regression pipeline (experiments/regression: train_regulariser -> train_approximate ->
temper_approximate -> metrics) on a toy curve, every stage on the CUDA path through the src-compatible shim.

    python examples/regression_pipeline.py        # needs a B200; prints the per-stage histories
"""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gvi_gaussian_process_b200.experiments.regression.metrics import calculate_metrics  # noqa: E402
from gvi_gaussian_process_b200.experiments.regression.trainers import train_regulariser_gp  # noqa: E402
from gvi_gaussian_process_b200.experiments.shared import (Data, TrainerSettings,  # noqa: E402
                                                          inducing_points_selector_resolver, train_approximate_gp,
                                                          train_tempered_gp)
from gvi_gaussian_process_b200.src.gps import ApproximateGPRegression  # noqa: E402
from gvi_gaussian_process_b200.src.kernels import TemperedKernel  # noqa: E402
from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel  # noqa: E402
from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel  # noqa: E402
from gvi_gaussian_process_b200.src.means import CustomMean  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    n, m = 20_000, 128
    x = np.sort(rng.uniform(-3.0, 3.0, (n, 1)), axis=0)
    y = np.sin(2.0 * x[:, 0]) + 0.15 * rng.standard_normal(n)
    dev = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")
    train = Data(x=dev(x[::2]), y=dev(y[::2]), name="train")
    test = Data(x=dev(x[1::2]), y=dev(y[1::2]), name="test")
    out = tempfile.mkdtemp()

    kernel = ARDKernel(number_of_dimensions=1)
    kernel_parameters = kernel.Parameters.construct(log_scaling=0.0, log_lengthscales=np.array([0.0]))
    exact, exact_parameters, history = train_regulariser_gp(
        data=train, empirical_risk_schema="negative_log_likelihood",
        trainer_settings=TrainerSettings(seed=0, optimiser_schema="adam", learning_rate=0.05, number_of_epochs=30,
                                         batch_size=m, batch_shuffle=True, batch_drop_last=False),
        kernel=kernel, kernel_parameters=kernel_parameters,
        inducing_points_selector=inducing_points_selector_resolver("conditional_variance"),
        number_of_inducing_points=m, empirical_risk_break_condition=-float("inf"), save_checkpoint_frequency=0,
        checkpoint_path=os.path.join(out, "regulariser"))
    print("regulariser NLL:", [round(float(h["empirical-risk"]), 4) for h in history][::5])

    torch.manual_seed(0)
    mean = CustomMean(mean_function=lambda p, xx: torch.tanh(xx @ p["w1"] + p["b1"]) @ p["w2"])
    sparse = SparsePosteriorKernel(base_kernel=kernel, inducing_points=exact.x)
    approximate = ApproximateGPRegression(mean=mean, kernel=sparse)
    parameters = approximate.Parameters.construct(
        mean=mean.Parameters.construct(custom={"w1": 0.5 * torch.randn(1, 32, dtype=torch.float64, device="cuda"),
                                               "b1": torch.zeros(32, dtype=torch.float64, device="cuda"),
                                               "w2": 0.1 * torch.randn(32, dtype=torch.float64, device="cuda")}),
        kernel=sparse.Parameters.construct(
            base_kernel=kernel.Parameters.construct(**exact_parameters.kernel.dict())))
    parameters, history = train_approximate_gp(
        data=train, empirical_risk_schema="negative_log_likelihood",
        regularisation_config={"regularisation_schema": "projected_renyi",
                               "regularisation_kwargs": {"mode": "posterior", "alpha": 0.5}},
        trainer_settings=TrainerSettings(seed=1, optimiser_schema="adabelief", learning_rate=0.02,
                                         number_of_epochs=20, batch_size=2000, batch_shuffle=True,
                                         batch_drop_last=False),
        approximate_gp=approximate, approximate_gp_parameters=parameters, regulariser=exact,
        regulariser_parameters=exact_parameters, save_checkpoint_frequency=0,
        checkpoint_path=os.path.join(out, "approximate"))
    print("GVI objective:", [round(float(h["gvi-objective"]), 4) for h in history][::4])

    tempered_kernel = TemperedKernel(base_kernel=sparse, base_kernel_parameters=parameters.kernel)
    tempered = ApproximateGPRegression(mean=mean, kernel=tempered_kernel)
    tempered_parameters = tempered.generate_parameters({"mean": parameters.mean,
                                                        "kernel": {"log_tempering_factor": 0.0}})
    print("before tempering:", calculate_metrics({"train": train, "test": test}, tempered, tempered_parameters))
    tempered_parameters, _ = train_tempered_gp(
        data=test, empirical_risk_schema="negative_log_likelihood",
        trainer_settings=TrainerSettings(seed=2, optimiser_schema="adam", learning_rate=0.1, number_of_epochs=30,
                                         batch_size=2000, batch_shuffle=True, batch_drop_last=False),
        tempered_gp=tempered, tempered_gp_parameters=tempered_parameters, save_checkpoint_frequency=0,
        checkpoint_path=os.path.join(out, "tempered"))
    print("after tempering: ", calculate_metrics({"train": train, "test": test}, tempered, tempered_parameters),
          "log_tempering_factor =", round(float(tempered_parameters.kernel.log_tempering_factor), 3))


if __name__ == "__main__":
    main()
