"""generate_batch (reference: src/utils/data.py:11-62): batches in the order of `jax.random.permutation(key, n)` -
restated host-side in utils/jax_prng.py so a seed gives the reference's batch order - gathered on the device."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import jax_prng


def generate_batch(key, data, batch_size: int, shuffle: bool = True, drop_last: bool = True):
    dataset_size = data[0].shape[0] if isinstance(data, tuple) else data.shape[0]
    batch_idx = jax_prng.permutation(key, dataset_size) if shuffle else np.arange(dataset_size)
    if drop_last:
        steps_per_epoch = math.floor(dataset_size / batch_size)
    else:
        steps_per_epoch = math.ceil(dataset_size / batch_size)
    steps_per_epoch = max(steps_per_epoch, 1)  # ensure at least one step

    def take(a, idx):
        if isinstance(a, torch.Tensor):
            return a[torch.as_tensor(idx, dtype=torch.int64, device=a.device)]
        return a[idx, ...]

    for idx in range(steps_per_epoch):
        curr_idx = np.asarray(batch_idx[idx * batch_size: (idx + 1) * batch_size], dtype=np.int64)
        if isinstance(data, tuple):
            yield tuple(take(a, curr_idx) for a in data)
        else:
            yield take(data, curr_idx)
