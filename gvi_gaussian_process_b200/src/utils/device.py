"""Host <-> device plumbing (torch is used for device memory and streams only)."""
from __future__ import annotations

import numpy as np
import torch

DTYPE = torch.float64


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("gvi_gaussian_process_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_device(x, dtype=DTYPE) -> torch.Tensor:
    """float64, C-contiguous CUDA tensor from a numpy array / torch tensor / python scalar."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.dtype != dtype or not t.is_cuda:
        t = t.to(device=device(), dtype=dtype, non_blocking=True)
    return t.contiguous()


def as_2d(x) -> torch.Tensor:
    """jnp.atleast_2d semantics: scalars -> [1,1], vectors -> [1,n]."""
    t = as_device(x)
    if t.ndim == 0:
        t = t.reshape(1, 1)
    elif t.ndim == 1:
        t = t.reshape(1, -1)
    return t


def scalar(v) -> torch.Tensor:
    """1-element device tensor holding a (possibly already device-resident) scalar parameter."""
    return as_device(v).reshape(-1)[:1].contiguous()


def empty(*shape) -> torch.Tensor:
    return torch.empty(*shape, dtype=DTYPE, device=device())
