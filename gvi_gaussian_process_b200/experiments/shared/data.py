"""Data container with the reference's on-disk layout (experiments/shared/data.py:13-35): `<name>.npz` holding the
arrays `x` and `y`."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch


def _host(a) -> np.ndarray:
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


@dataclass
class Data:
    x: Any
    y: Any
    name: str = "data"

    def __add__(self, other):
        cat = torch.cat if isinstance(self.x, torch.Tensor) else np.concatenate
        return Data(name=self.name + other.name, x=cat([self.x, other.x], 0), y=cat([self.y, other.y], 0))

    def save(self, path: str):
        np.savez(os.path.join(path, f"{self.name}.npz"), x=_host(self.x), y=_host(self.y))

    @staticmethod
    def load(path: str, name: str):
        data_path = os.path.join(path, f"{name}.npz")
        if os.path.isfile(data_path):
            data = np.load(data_path)
            return Data(name=name, x=np.array(data["x"]), y=np.array(data["y"]))
        return None
