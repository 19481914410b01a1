"""SparsePosteriorKernel (reference: src/kernels/approximate/sparse_posterior_kernel.py:14-96).

r(x1,x2) = k(x1,x2) - k(x1,Z) (K_ZZ + jitter I)^-1 k(Z,x2).  The diagonal mode - the one on the training hot
path (predict_probability, NLL, regularisers) - runs in the fused DMMA kernel and never materialises K(X,Z).
The full-covariance mode (off the hot path) materialises the three Grams with the CUDA Gram kernel and uses the
library DGEMM (cuBLAS through torch.matmul) against the L^-1 produced by the CUDA factorisation.
"""
from __future__ import annotations

from ... import _lib_proxy as L
from ..._autograd import PredictVariance, tensor_leaves, wants_grad
from ...utils.device import as_2d
from ..base import KernelBase
from .base import ApproximateBaseKernel, ApproximateBaseKernelParameters
from ..standard.ard_kernel import ARDKernel


class SparsePosteriorKernelParameters(ApproximateBaseKernelParameters):
    """base_kernel: parameters of the base kernel."""


class SparsePosteriorKernel(ApproximateBaseKernel):
    Parameters = SparsePosteriorKernelParameters

    def __init__(self, base_kernel: KernelBase, inducing_points, diagonal_regularisation: float = 1e-5,
                 is_diagonal_regularisation_absolute_scale: bool = False, preprocess_function=None):
        # ARD: the fused path (K tiles evaluated inside the tile kernel, gradients).  Any other base kernel: its own
        # Gram routine supplies K_ZZ and K(x_chunk, Z); factorisation and predictive still run in the CUDA kernels.
        assert isinstance(base_kernel, KernelBase), f"base_kernel must be a kernel, got {type(base_kernel)=}"
        self._ard = ARDKernel.is_fast_path(base_kernel)
        self.base_kernel = base_kernel
        self.inducing_points = as_2d(inducing_points)
        if self._ard:
            self.inducing_points = self.inducing_points.reshape(self.inducing_points.shape[0], -1)
        self.diagonal_regularisation = diagonal_regularisation
        self.is_diagonal_regularisation_absolute_scale = is_diagonal_regularisation_absolute_scale
        self._factor = None
        super().__init__(preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> SparsePosteriorKernelParameters:
        return SparsePosteriorKernel.Parameters(
            base_kernel=self.base_kernel._to_parameters(parameters["base_kernel"]))

    def factorise(self, parameters) -> "L.GpFactor":
        """cho_factor(add_diagonal_regulariser(K_ZZ)) of lines 63-74, plus L^-1 packed for the tensor pipe."""
        parameters = self._to_parameters(parameters)
        z = self.inducing_points
        if not self._ard:
            if self._factor is None:
                self._factor = L.GpFactor(z.shape[0], 1)
            k_zz = self.base_kernel.calculate_gram(parameters.base_kernel, z, z).contiguous()
            return self._factor.factor_from_gram(k_zz, self.diagonal_regularisation,
                                                 self.is_diagonal_regularisation_absolute_scale)
        if self._factor is None:
            self._factor = L.GpFactor(z.shape[0], z.shape[1])
        ls, ll = ARDKernel.device_parameters(self.base_kernel._to_parameters(parameters.base_kernel))
        return self._factor.factor(z, ls, ll, self.diagonal_regularisation,
                                   self.is_diagonal_regularisation_absolute_scale)

    def _calculate_gram_diagonal(self, parameters, x):
        if not self._ard and wants_grad(*tensor_leaves(parameters.base_kernel)):
            if self._factor is None:
                self._factor = L.GpFactor(self.inducing_points.shape[0], 1)
            _, var = L.gp_predict_any_kernel_grad(
                self._factor, self.base_kernel, parameters.base_kernel, self.inducing_points, x,
                self.diagonal_regularisation, self.is_diagonal_regularisation_absolute_scale, want_mean=False)
            return var
        f = self.factorise(parameters)
        if not self._ard:
            _, var = L.gp_predict_any_kernel(f, self.base_kernel, parameters.base_kernel, self.inducing_points, x,
                                             want_mean=False)
            return var
        x = x.reshape(x.shape[0], -1)
        base = self.base_kernel._to_parameters(parameters.base_kernel)
        if wants_grad(base.log_scaling, base.log_lengthscales):
            # torch autograd front end: backward = gvi_gp_predict_grad_f64 (the variance VJP on the tensor pipe)
            return PredictVariance.apply(f, lambda: self.factorise(parameters), x, False, base.log_scaling,
                                         base.log_lengthscales)
        _, var = L.gp_predict(f, x, want_mean=False)
        return var

    def _calculate_gram(self, parameters, x1, x2):
        f = self.factorise(parameters)
        base_parameters = parameters.base_kernel
        linv = f.inverse_cholesky()
        k1z = self.base_kernel.calculate_gram(base_parameters, x1, self.inducing_points)
        k2z = self.base_kernel.calculate_gram(base_parameters, x2, self.inducing_points)
        k12 = self.base_kernel.calculate_gram(base_parameters, x1, x2)
        return k12 - (k1z @ linv.T) @ (k2z @ linv.T).T
