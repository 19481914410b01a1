"""reference: experiments/shared/utils.py:12-31."""
from __future__ import annotations

import torch

from ...src.utils.device import as_device
from .data import Data


def calculate_inducing_points(key, inducing_points_selector, data: Data, number_of_inducing_points: int, kernel,
                              kernel_parameters) -> Data:
    x = as_device(data.x)
    x = x.reshape(x.shape[0], -1) if x.ndim > 1 else x.reshape(1, -1)
    x_inducing, inducing_indices = inducing_points_selector.compute_inducing_points(
        key=key, training_inputs=x, number_of_inducing_points=number_of_inducing_points, kernel=kernel,
        kernel_parameters=kernel_parameters)
    y = as_device(data.y)
    return Data(x=x_inducing, y=y[torch.as_tensor(inducing_indices, device=y.device)])
