"""Dimension asserts (reference: src/utils/checks.py:4-38)."""


def check_matching_dimensions(x, y) -> None:
    assert x.shape[1] == y.shape[1], (
        f"Dimension Mismatch: {x.shape[1]=} needs to match {y.shape[1]=}")


def check_maximum_dimension(x, maximum_dimensionality: int) -> None:
    # NB: the reference's check of the same name actually asserts ndim >= maximum (checks.py:36-38); mirrored.
    assert x.ndim >= maximum_dimensionality, (
        f"The number of dimensions of x must be at most {maximum_dimensionality}, got {x.ndim=}")
