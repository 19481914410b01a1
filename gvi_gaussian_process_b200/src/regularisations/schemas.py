"""reference: src/regularisations/schemas.py:4-10."""
import enum


class RegularisationMode(str, enum.Enum):
    prior = "prior"
    posterior = "posterior"
