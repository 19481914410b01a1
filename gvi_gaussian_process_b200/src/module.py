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

    # -- checkpoints (reference: src/module.py:27-55 writes orbax checkpoint directories; orbax is a third-party
    #    format that is not available here, so the same `.dict()` tree is stored as one .npz with "/"-joined keys) ----
    def save(self, path: str) -> None:
        import json
        import os

        import numpy as np

        arrays = {}

        def walk(node):
            """JSON description of the tree; leaves are stored as arrays a0, a1, ... (empty containers, None leaves, tuples
            and arbitrary dict keys survive the round trip)."""
            if isinstance(node, dict):
                return {"t": "dict", "k": [str(k) for k in node], "v": [walk(v) for v in node.values()]}
            if isinstance(node, (list, tuple)):
                return {"t": "tuple" if isinstance(node, tuple) else "list", "v": [walk(v) for v in node]}
            if node is None:
                return {"t": "none"}
            name = f"a{len(arrays)}"
            arrays[name] = node.detach().cpu().numpy() if hasattr(node, "detach") else np.asarray(node)
            return {"t": "array", "a": name}

        structure = json.dumps(walk(self.dict()))
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        np.savez(path if path.endswith(".npz") else path + ".npz", __structure__=np.array(structure), **arrays)

    def load(self, path: str) -> "ModuleParameters":
        import json

        import numpy as np

        data = np.load(path if path.endswith(".npz") else path + ".npz")
        if "__structure__" in data.files:

            def build(desc):
                kind = desc["t"]
                if kind == "dict":
                    return {k: build(v) for k, v in zip(desc["k"], desc["v"])}
                if kind in ("list", "tuple"):
                    out = [build(v) for v in desc["v"]]
                    return tuple(out) if kind == "tuple" else out
                if kind == "none":
                    return None
                return data[desc["a"]]

            return self.construct(**build(json.loads(str(data["__structure__"]))))
        # files written before the structure manifest existed: "/"-joined keys, "#i" list members
        tree: Dict[str, Any] = {}
        for key in data.files:
            node = tree
            parts = key.split("/")
            for part in parts[:-1]:
                node = node.setdefault(part, {})
            node[parts[-1]] = data[key]

        def lists(node):
            if isinstance(node, dict):
                if node and all(k.startswith("#") for k in node):
                    return [lists(node[f"#{i}"]) for i in range(len(node))]
                return {k: lists(v) for k, v in node.items()}
            return node

        return self.construct(**lists(tree))

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
        if isinstance(parameters, ModuleParameters) and not isinstance(parameters, self.Parameters):
            Module.check_parameters(parameters, self.Parameters)  # another module's parameters: AssertionError
        if not isinstance(parameters, self.Parameters):
            parameters = self.generate_parameters(dict(parameters))
        elif any(isinstance(v, dict) for v in parameters.dict_shallow().values()):
            # `Parameters.construct(**tree)` with plain-dict members (what the reference's trainer builds after every
            # optimiser step, experiments/shared/trainer.py:92-94): regenerate the nested parameter objects
            parameters = self.generate_parameters(parameters.dict_shallow())
        Module.check_parameters(parameters, self.Parameters)
        return parameters
