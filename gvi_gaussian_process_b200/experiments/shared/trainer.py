"""Trainer / TrainerSettings (reference: experiments/shared/trainer.py:19-100).

Same control flow as the reference - per epoch: optional checkpoint, `key, subkey = split(key)`, batches in
`jax.random.permutation(subkey, n)` order, NaN check on the loss, gradient step, post-epoch callback, break condition -
with everything device-resident: the trainable leaves of `parameters.dict()` are views into ONE flat float64 device
vector, `loss_function(parameters_dict, x, y).backward()` (the CUDA vector-Jacobian products behind the torch autograd
front end) accumulates straight into a flat gradient vector, and the optimiser (optax adam / adabelief / rmsprop rules,
experiments/shared/resolvers/optimiser.py:6-16) is one CUDA launch over the flat vectors (gvi_optimiser_step_f64).
Only the loss scalar visits the host (the reference's NaN test does the same).

Multi-GPU (one process per GPU, torch.distributed initialised): every rank walks the SAME batch sequence (same seed ->
same permutation) and evaluates its contiguous slice of each batch (distributed.shard_bounds); the flat gradient vector
is combined with ONE all-reduce per step - weighted by the slice sizes for objectives that are means over points
(`gradient_reduction="mean"`, the default), plainly summed for objectives that are sums over points (the multinomial
NLL) - so N ranks take exactly the single-GPU optimisation path."""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ... import _lib
from ...distributed import shard_bounds
from ...src.module import ModuleParameters
from ...src.utils import jax_prng
from ...src.utils.data import generate_batch
from ...src.utils.device import device
from .data import Data
from .schemas import OptimiserSchema

# optax 0.1.7 defaults of adam / adabelief / rmsprop: (kind, b1, b2 or decay, eps, eps_root)
_OPTIMISERS = {
    OptimiserSchema.adam: (0, 0.9, 0.999, 1e-8, 0.0),
    OptimiserSchema.adabelief: (1, 0.9, 0.999, 1e-16, 1e-16),
    OptimiserSchema.rmsprop: (2, 0.0, 0.9, 1e-8, 0.0),
}


@dataclass
class TrainerSettings:
    seed: int
    optimiser_schema: OptimiserSchema
    learning_rate: float
    number_of_epochs: int
    batch_size: int
    batch_shuffle: bool
    batch_drop_last: bool


def _is_trainable(v) -> bool:
    if isinstance(v, torch.Tensor):
        return v.is_floating_point() and bool(torch.isfinite(v).all())
    if isinstance(v, (float, np.floating)):
        return math.isfinite(v)
    if isinstance(v, np.ndarray):
        return v.dtype.kind == "f" and bool(np.isfinite(v).all())
    return False


class FlatParameters:
    """The float leaves of a (nested dict / list / tuple) parameter tree as views into one flat device vector."""

    def __init__(self, tree: Dict[str, Any]):
        self.shapes: List[Tuple[int, ...]] = []
        values = []
        self._collect(tree, values)
        n = sum(int(np.prod(s)) if len(s) else 1 for s in self.shapes)
        self.flat = torch.empty(n, dtype=torch.float64, device=device())
        self.grad = torch.zeros_like(self.flat)
        self.leaves = []
        o = 0
        for v, s in zip(values, self.shapes):
            k = int(np.prod(s)) if len(s) else 1
            self.flat[o:o + k] = torch.as_tensor(np.asarray(v.detach().cpu() if isinstance(v, torch.Tensor) else v),
                                                 dtype=torch.float64).reshape(-1).to(self.flat.device)
            leaf = self.flat[o:o + k].view(s).detach().requires_grad_(True)
            leaf.grad = self.grad[o:o + k].view(s)  # backward accumulates in place into the flat gradient
            self.leaves.append(leaf)
            o += k
        self._it = None
        self.tree = self._rebuild(tree, iter(self.leaves))

    def _collect(self, node, values):
        if isinstance(node, dict):
            for v in node.values():
                self._collect(v, values)
        elif isinstance(node, (list, tuple)):
            for v in node:
                self._collect(v, values)
        elif _is_trainable(node):
            values.append(node)
            self.shapes.append(tuple(node.shape) if hasattr(node, "shape") else ())

    def _rebuild(self, node, it):
        if isinstance(node, dict):
            return {k: self._rebuild(v, it) for k, v in node.items()}
        if isinstance(node, (list, tuple)):
            return type(node)(self._rebuild(v, it) for v in node)
        if _is_trainable(node):
            return next(it)
        return node

    def detached_tree(self):
        """A copy of the tree whose leaves are plain (non-view, no-grad) tensors - what train() returns."""
        return self._rebuild(self.tree, iter([leaf.detach().clone() for leaf in self.leaves]))


class Trainer:
    def __init__(self, save_checkpoint_frequency: int, checkpoint_path: str,
                 post_epoch_callback: Callable[[ModuleParameters], Dict[str, float]],
                 break_condition_function: Callable[[ModuleParameters], bool] = None,
                 gradient_reduction: str = "mean", group=None):
        assert gradient_reduction in ("mean", "sum")
        self.gradient_reduction = gradient_reduction
        self.group = group
        self.save_checkpoint_frequency = save_checkpoint_frequency
        self.checkpoint_path = checkpoint_path
        self.post_epoch_callback = post_epoch_callback
        self.break_condition_function = break_condition_function

    def train(self, trainer_settings: TrainerSettings, parameters: ModuleParameters, data: Data,
              loss_function: Callable[[Dict, Any, Any], torch.Tensor], disable_tqdm: bool = True):
        lib = _lib.load()
        post_epoch_history = []
        kind, b1, b2, eps, eps_root = _OPTIMISERS[OptimiserSchema(trainer_settings.optimiser_schema)]
        flat = FlatParameters(parameters.dict())
        mu, nu = torch.zeros_like(flat.flat), torch.zeros_like(flat.flat)
        x_all = data.x if isinstance(data.x, torch.Tensor) else torch.as_tensor(
            np.asarray(data.x), dtype=torch.float64, device=device())
        y_all = data.y if isinstance(data.y, torch.Tensor) else torch.as_tensor(
            np.asarray(data.y), dtype=torch.float64, device=device())
        current = lambda: parameters.construct(**flat.detached_tree())
        key = jax_prng.PRNGKey(trainer_settings.seed)
        count = 0
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if world > 1 else 0
        packed = torch.zeros(flat.flat.numel() + 2, dtype=torch.float64, device=flat.flat.device) if world > 1 else None
        for epoch in range(trainer_settings.number_of_epochs):
            if self.save_checkpoint_frequency and epoch % self.save_checkpoint_frequency == 0:
                current().save(path=os.path.join(self.checkpoint_path, f"epoch-{epoch}.ckpt"))
            key, subkey = jax_prng.split(key)
            for x_batch, y_batch in generate_batch(key=subkey, data=(x_all, y_all),
                                                   batch_size=trainer_settings.batch_size,
                                                   shuffle=trainer_settings.batch_shuffle,
                                                   drop_last=trainer_settings.batch_drop_last):
                flat.grad.zero_()
                if world > 1:
                    lo, hi = shard_bounds(x_batch.shape[0], rank, world)
                    n_local = hi - lo
                    weight = float(n_local) if self.gradient_reduction == "mean" else 1.0
                    packed.zero_()
                    if n_local > 0:
                        loss = loss_function(flat.tree, x_batch[lo:hi], y_batch[lo:hi])
                        loss.backward()
                        packed[:-2] = flat.grad * weight
                        packed[-2] = weight
                        packed[-1] = loss.detach() * weight
                    dist.all_reduce(packed, group=self.group)  # gradient, normaliser and loss in one collective
                    norm = packed[-2] if self.gradient_reduction == "mean" else 1.0
                    flat.grad.copy_(packed[:-2] / norm)
                    if bool(torch.isnan(packed[-1])):
                        return current(), post_epoch_history
                else:
                    loss = loss_function(flat.tree, x_batch, y_batch)
                    if bool(torch.isnan(loss.detach())):
                        return current(), post_epoch_history
                    loss.backward()
                count += 1
                _lib.check(lib.gvi_optimiser_step_f64(kind, flat.flat.data_ptr(), flat.grad.data_ptr(), mu.data_ptr(),
                                                      nu.data_ptr(), flat.flat.numel(), count,
                                                      float(trainer_settings.learning_rate), b1, b2, eps, eps_root,
                                                      _lib.current_stream()), "gvi_optimiser_step_f64")
            post_epoch_history.append(self.post_epoch_callback(current()))
            if self.break_condition_function and self.break_condition_function(current()):
                break
        return current(), post_epoch_history
