"""Module / ModuleParameters (reference: src/module.py:19-94).

The reference uses pydantic-v1 models; here parameters are plain attribute containers with the same
`.construct(**kwargs)` / `.dict()` surface, so README-style code keeps working.
"""
from __future__ import annotations

from typing import Any, Dict


class ModuleParameters:
    _fields: tuple = ()

    def __init__(self, **kwargs: Any):
        for k, v in kwargs.items():
            setattr(self, k, v)
        self._keys = tuple(kwargs.keys())

    @classmethod
    def construct(cls, **kwargs: Any) -> "ModuleParameters":
        return cls(**kwargs)

    def dict(self) -> Dict[str, Any]:
        out = {}
        for k in self._keys:
            v = getattr(self, k)
            out[k] = v.dict() if isinstance(v, ModuleParameters) else v
        return out

    def dict_shallow(self) -> Dict[str, Any]:
        return {k: getattr(self, k) for k in self._keys}

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(f'{k}={getattr(self, k)!r}' for k in self._keys)})"


class Module:
    Parameters = ModuleParameters

    def generate_parameters(self, parameters) -> ModuleParameters:
        raise NotImplementedError

    @staticmethod
    def check_parameters(parameters: ModuleParameters, parameter_type) -> None:
        # reference: src/module.py:92-94
        assert isinstance(parameters, parameter_type), (
            f"Parameters is type: {type(parameters)=}, needs to be {parameter_type=}")

    def _to_parameters(self, parameters):
        if not isinstance(parameters, self.Parameters):
            parameters = self.generate_parameters(dict(parameters))
        Module.check_parameters(parameters, self.Parameters)
        return parameters
