"""gvi_gaussian_process_b200 - B200 (sm_100a) implementation of the GP hot path of jswu18/gvi-gaussian-process.

`gvi_gaussian_process_b200.src` mirrors the reference's `src/` API surface for the hot path
(ARDKernel, SparsePosteriorKernel, GPRegression, ApproximateGPRegression, NegativeLogLikelihood, the projected /
Wasserstein / squared-difference regularisations, GeneralisedVariationalInference,
ConditionalVarianceInducingPointsSelector).  All numerics run in hand-written CUDA behind the C ABI in
include/gvigp.h (libgvigp.so); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
