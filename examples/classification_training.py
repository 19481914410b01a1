"""Multi-class GVI on the B200 path, end to end (the shape of experiments/classification/mnist/main.py of the reference):

  1. per class: greedy conditional-variance inducing points (pivoted-Cholesky kernel),
  2. regulariser: exact GP classifier on the combined inducing data, trained by NLL for a few epochs
     (train_regulariser_gp; exact-GP VJP + device-resident AdaBelief),
  3. approximate GP: K sparse-posterior ARD kernels over the per-class inducing sets + a small network mean, trained on
     multinomial NLL + projected KL against the regulariser posterior (train_approximate_gp),
  4. accuracy of both on held-out points.

Run on a B200:  python examples/classification_training.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvi_gaussian_process_b200.experiments.shared import Data, OptimiserSchema, Trainer, TrainerSettings  # noqa: E402
from gvi_gaussian_process_b200.src.empirical_risks import NegativeLogLikelihood  # noqa: E402
from gvi_gaussian_process_b200.src.generalised_variational_inference import GeneralisedVariationalInference  # noqa: E402
from gvi_gaussian_process_b200.src.gps import ApproximateGPClassification, GPClassification  # noqa: E402
from gvi_gaussian_process_b200.src.inducing_points_selection import ConditionalVarianceInducingPointsSelector  # noqa: E402
from gvi_gaussian_process_b200.src.kernels import MultiOutputKernel  # noqa: E402
from gvi_gaussian_process_b200.src.kernels.approximate import SparsePosteriorKernel  # noqa: E402
from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel  # noqa: E402
from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean  # noqa: E402
from gvi_gaussian_process_b200.src.regularisations.projected import ProjectedKLRegularisation  # noqa: E402
from gvi_gaussian_process_b200.src.utils.jax_prng import PRNGKey  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    k, d, n_per_class, m_per_class = 3, 2, 600, 20
    centres = np.array([[0.0, 1.5], [1.3, -0.8], [-1.3, -0.8]])
    x = np.concatenate([c + 0.7 * rng.standard_normal((n_per_class, d)) for c in centres])
    labels = np.repeat(np.arange(k), n_per_class)
    order = rng.permutation(len(x))
    x, labels = x[order], labels[order]
    y = np.eye(k)[labels]
    n_train = int(0.8 * len(x))
    x_tr, y_tr, x_te, lab_te = x[:n_train], y[:n_train], x[n_train:], labels[n_train:]

    ard = lambda: ARDKernel(number_of_dimensions=d)
    kp0 = {"log_scaling": 0.0, "log_lengthscales": np.zeros(d)}
    # 1. inducing points per class (experiments/classification/trainers.py:41-55)
    selector = ConditionalVarianceInducingPointsSelector()
    z_list, yz_list = [], []
    for c in range(k):
        xc = x_tr[y_tr[:, c] == 1]
        z, idx = selector.compute_inducing_points(PRNGKey(c), xc, m_per_class, ard(), kp0)
        z_list.append(z)
        yz_list.append(torch.as_tensor(np.eye(k)[np.full(m_per_class, c)], dtype=torch.float64, device=z.device))
    z_all, yz_all = torch.cat(z_list), torch.cat(yz_list)

    # 2. regulariser: exact GP classifier on the inducing data (NLL training)
    reg_kernel = MultiOutputKernel(kernels=[ard() for _ in range(k)])
    regulariser = GPClassification(mean=ConstantMean(number_output_dimensions=k), kernel=reg_kernel, x=z_all, y=yz_all)
    reg_params = regulariser.generate_parameters({"log_observation_noise": np.zeros(k), "mean": {"constant": np.zeros(k)},
                                                  "kernel": {"kernels": [kp0] * k}})
    accuracy = lambda gp, p: float((gp.predict_probability(p, x_te).probabilities.argmax(1).cpu().numpy()
                                    == lab_te).mean())
    print(f"regulariser GP accuracy (untrained hyper-parameters): {accuracy(regulariser, reg_params):.3f}")

    # 3. approximate GP: network mean + per-class sparse posterior kernels; NLL + projected KL
    torch.manual_seed(0)
    w1 = (torch.randn(d, 16, dtype=torch.float64, device="cuda") / np.sqrt(d))
    w2 = (torch.randn(16, k, dtype=torch.float64, device="cuda") / 4.0)
    mean = CustomMean(mean_function=lambda p, xx: (torch.tanh(xx @ p[0]) @ p[1]).T, number_output_dimensions=k)
    kernel = MultiOutputKernel(kernels=[SparsePosteriorKernel(base_kernel=ard(), inducing_points=z_list[c],
                                                              diagonal_regularisation=1e-6) for c in range(k)])
    approx = ApproximateGPClassification(mean=mean, kernel=kernel)
    params = approx.Parameters.construct(mean={"custom": (w1, w2)},
                                         kernel={"kernels": [{"base_kernel": dict(kp0)} for _ in range(k)]})
    gvi = GeneralisedVariationalInference(
        empirical_risk=NegativeLogLikelihood(gp=approx),
        regularisation=ProjectedKLRegularisation(gp=approx, regulariser=regulariser, regulariser_parameters=reg_params,
                                                 mode="posterior"))
    settings = TrainerSettings(seed=0, optimiser_schema=OptimiserSchema.adabelief, learning_rate=2e-2,
                               number_of_epochs=15, batch_size=480, batch_shuffle=True, batch_drop_last=False)
    trainer = Trainer(save_checkpoint_frequency=0, checkpoint_path="",
                      post_epoch_callback=lambda p: {"gvi-objective": float(gvi.calculate_loss(p, x_tr, y_tr))})
    print(f"approximate GP accuracy before training: {accuracy(approx, params):.3f}")
    params, history = trainer.train(settings, params, Data(x=x_tr, y=y_tr),
                                    lambda pd, xb, yb: gvi.calculate_loss(pd, xb, yb))
    print("gvi objective per epoch:", " ".join(f"{h['gvi-objective']:.1f}" for h in history))
    acc = accuracy(approx, params)
    print(f"approximate GP accuracy after training: {acc:.3f}")
    assert history[-1]["gvi-objective"] < history[0]["gvi-objective"] and acc > 0.8


if __name__ == "__main__":
    main()
