"""The reference README walkthrough (README.md:55-404 of jswu18/gvi-gaussian-process) on the B200 path:
noisy sine -> greedy conditional-variance inducing points -> exact GP regulariser -> approximate GP with a tanh-FCN
mean + sparse posterior kernel -> GVI objective (Gaussian NLL + projected Renyi) trained with torch autograd / Adam
(the reference uses jax.grad / optax.adabelief).  Run on a machine with a B200:  python examples/readme_regression.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood
from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference
from gvi_gaussian_process_b200.src.gps import ApproximateGPRegression, GPRegression
from gvi_gaussian_process_b200.src.inducing_points_selection import ConditionalVarianceInducingPointsSelector
from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel
from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean
from gvi_gaussian_process_b200.src.regularisations.projected import ProjectedRenyiRegularisation
from gvi_gaussian_process_b200.src.utils.jax_prng import PRNGKey, split


def main(steps: int = 200):
    dev = torch.device("cuda")
    number_of_points, noise = 100, 0.1
    key = PRNGKey(0)
    key, subkey = split(key)
    rng = np.random.default_rng(0)
    x = np.linspace(-1, 1, number_of_points).reshape(-1, 1)
    y = (np.sin(np.pi * x) + noise * rng.standard_normal(x.shape)).reshape(-1)

    kernel = ARDKernel(number_of_dimensions=1)
    kernel_parameters = kernel.Parameters.construct(log_scaling=np.log(10.0), log_lengthscales=np.log(10.0))
    key, subkey = split(key)
    inducing_points, inducing_idx = ConditionalVarianceInducingPointsSelector().compute_inducing_points(
        key=subkey, training_inputs=x, number_of_inducing_points=int(np.sqrt(number_of_points)), kernel=kernel,
        kernel_parameters=kernel_parameters)
    inducing_responses = torch.as_tensor(y, device=dev)[inducing_idx]

    mean = ConstantMean()
    exact_gp = GPRegression(mean=mean, kernel=kernel, x=inducing_points, y=inducing_responses)
    exact_parameters = exact_gp.Parameters.construct(log_observation_noise=np.log(1.0),
                                                     mean=mean.Parameters.construct(constant=0.0),
                                                     kernel=kernel_parameters)
    nll_exact = NegativeLogLikelihood(gp=exact_gp)
    print("exact GP NLL on the inducing set:",
          float(nll_exact.calculate_empirical_risk(exact_parameters, inducing_points, inducing_responses)))

    torch.manual_seed(0)
    fcnn = {"w1": (0.5 * torch.randn(1, 10, dtype=torch.float64, device=dev)).requires_grad_(True),
            "b1": torch.zeros(10, dtype=torch.float64, device=dev, requires_grad=True),
            "w2": (0.3 * torch.randn(10, 1, dtype=torch.float64, device=dev)).requires_grad_(True),
            "b2": torch.zeros(1, dtype=torch.float64, device=dev, requires_grad=True)}
    approximate_mean = CustomMean(mean_function=lambda p, xx: torch.tanh(xx @ p["w1"] + p["b1"]) @ p["w2"] + p["b2"])
    log_scaling = torch.tensor(np.log(10.0), dtype=torch.float64, device=dev, requires_grad=True)
    log_lengthscales = torch.tensor(np.log(10.0), dtype=torch.float64, device=dev, requires_grad=True)
    approximate_kernel = SparsePosteriorKernel(base_kernel=kernel, inducing_points=inducing_points)
    approximate_gp = ApproximateGPRegression(mean=approximate_mean, kernel=approximate_kernel)
    approximate_parameters = approximate_gp.Parameters.construct(
        mean=approximate_mean.Parameters.construct(custom=fcnn),
        kernel=approximate_kernel.Parameters.construct(
            base_kernel=kernel.Parameters.construct(log_scaling=log_scaling, log_lengthscales=log_lengthscales)))

    gvi = GeneralisedVariationalInference(
        empirical_risk=NegativeLogLikelihood(gp=approximate_gp),
        regularisation=ProjectedRenyiRegularisation(gp=approximate_gp, regulariser=exact_gp,
                                                    regulariser_parameters=exact_parameters, mode="posterior",
                                                    alpha=0.5))
    optimiser = torch.optim.Adam(list(fcnn.values()) + [log_scaling, log_lengthscales], lr=1e-2)
    for step in range(steps):
        optimiser.zero_grad()
        loss = gvi.calculate_loss(approximate_parameters, x, y)
        loss.backward()
        optimiser.step()
        if step % 50 == 0 or step == steps - 1:
            print(f"step {step:4d}  GVI loss {float(loss.detach()):.6f}")
    prediction = approximate_gp.predict_probability(approximate_parameters, x)
    rmse = float(((prediction.mean - torch.as_tensor(np.sin(np.pi * x).reshape(-1), device=dev)) ** 2).mean().sqrt())
    print(f"approximate GP: RMSE to the noiseless sine {rmse:.4f}, mean predictive std "
          f"{float(prediction.covariance.sqrt().mean()):.4f}")


if __name__ == "__main__":
    main()
