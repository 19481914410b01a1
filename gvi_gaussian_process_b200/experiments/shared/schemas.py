"""reference: experiments/shared/schemas.py:4-8."""
import enum


class OptimiserSchema(str, enum.Enum):
    adam = "adam"
    adabelief = "adabelief"
    rmsprop = "rmsprop"
