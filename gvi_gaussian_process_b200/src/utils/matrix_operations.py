"""src.utils.matrix_operations (reference: src/utils/matrix_operations.py:7-28).

`add_diagonal_regulariser` is row a4 of the hot path: K + lambda' I with lambda' = lambda (absolute) or
lambda * trace(K) / M (relative), one CUDA launch (gvi_add_diagonal_regulariser_f64).  Like the reference it returns a
NEW matrix and leaves its argument untouched.  The two eigenvalue helpers (lines 46-92) serve only the Gaussian
Wasserstein regulariser WITH eigendecomposition, which is out of scope (O(N^3), not data-parallel)."""
from __future__ import annotations

from .. import _lib_proxy as L
from .device import as_2d


def add_diagonal_regulariser(matrix, diagonal_regularisation: float, is_diagonal_regularisation_absolute_scale: bool):
    out = as_2d(matrix).clone()
    assert out.shape[0] == out.shape[1], f"matrix must be square, got {tuple(out.shape)}"
    L.add_diagonal_regulariser(out, diagonal_regularisation, is_diagonal_regularisation_absolute_scale)
    return out


def compute_covariance_eigenvalues(matrix):
    raise NotImplementedError("eigenvalue helpers of the Gaussian Wasserstein regulariser with eigendecomposition are "
                              "out of scope (DESIGN.md section 8)")


def compute_product_eigenvalues(matrix_a, matrix_b):
    raise NotImplementedError("eigenvalue helpers of the Gaussian Wasserstein regulariser with eigendecomposition are "
                              "out of scope (DESIGN.md section 8)")
