"""src.utils.custom_types (reference: src/utils/custom_types.py:16-51).  The reference's JaxArrayType / JaxFloatType are
pydantic field types that coerce a value to jnp.array(val, float64) / jnp.float64(val).  The parameter containers here
are plain attribute holders (src/module.py), so these are the matching coercions onto the CUDA device: float64 tensors
(0-d for JaxFloatType), usable as annotations and as `X.validate(value)`."""
from __future__ import annotations

from typing import Any, Generic, TypeVar

from .device import as_device

ArrayDType = TypeVar("ArrayDType")
FloatDType = TypeVar("FloatDType")
PRNGKey = Any  # pylint: disable=invalid-name


class JaxArrayType(Generic[ArrayDType]):
    @classmethod
    def validate(cls, val):
        return as_device(val)


class JaxFloatType(Generic[FloatDType]):
    @classmethod
    def validate(cls, val):
        return as_device(val).reshape(())
