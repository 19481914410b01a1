"""Non-stationary kernels: k depends on x and y, not only on x - y (reference: src/kernels/non_stationary/base.py).
Both members are functions of the inner product <x, y>, so one CUDA Gram kernel (csrc/gram.cu:dot_gram_kernel) with a
fused (scale, constant, power) epilogue serves them; k(x, x) varies per point, so predictives over these kernels take
the caller-supplied-Gram entry points (gvi_gp_factor_gram_f64 / gvi_gp_predict_gram_f64)."""
from __future__ import annotations

import torch

from ... import _lib_proxy as L
from ..._autograd import DotGram, wants_grad
from ...utils.device import empty, scalar
from ..base import KernelBase, KernelBaseParameters


class NonStationaryKernelBaseParameters(KernelBaseParameters):
    pass


class NonStationaryKernelBase(KernelBase):
    Parameters = NonStationaryKernelBaseParameters

    # (log_scaling, log_constant or None, degree) of k = (exp(log_scaling) <x, y> + exp(log_constant))^degree
    def _dot_parameters(self, parameters):
        raise NotImplementedError

    @staticmethod
    def _flat(x):
        return x.reshape(x.shape[0], -1) if x.ndim > 2 else x

    def _calculate_gram(self, parameters, x1, x2):
        x1, x2 = self._flat(x1), self._flat(x2)
        log_scaling, log_constant, degree = self._dot_parameters(parameters)
        if wants_grad(log_scaling, log_constant, x1, x2):
            # forward = the same CUDA Gram kernel; backward w.r.t. the scalars and the inputs (feature maps)
            return DotGram.apply(x1, x2, log_scaling, log_constant, degree)
        out = empty(x1.shape[0], x2.shape[0])
        L.dot_gram(x1, x2, log_scaling, log_constant, degree, out)
        return out

    def _calculate_gram_pairs(self, parameters, x1, x2):
        x1, x2 = self._flat(x1), self._flat(x2)
        log_scaling, log_constant, degree = self._dot_parameters(parameters)
        if wants_grad(log_scaling, log_constant, x1, x2):
            # O(n d) pairwise values: host-framework expression, differentiated by torch
            base = torch.exp(log_scaling) * (x1 * x2).sum(-1)
            if log_constant is not None:
                base = base + torch.exp(log_constant)
            return base if degree == 1.0 else base ** degree
        out = empty(x1.shape[0])
        L.kernel_pairs(L.KERNEL_POLYNOMIAL, x1, x2, log_scaling, log_constant, degree, out)
        return out

    def _calculate_gram_diagonal(self, parameters, x):
        return self._calculate_gram_pairs(parameters, x, x)

    @staticmethod
    def _scalar(v):
        return scalar(v)
