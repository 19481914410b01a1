"""FixedSparsePosteriorKernel (reference: src/kernels/approximate/fixed_sparse_posterior_kernel.py:14-99):
r(x1,x2) = k_b(x1,x2) - k_b(x1,Z) cho_solve(chol(K_reg(Z,Z) + jitter), k_b(Z,x2)) - the Cholesky factor comes from
the regulariser kernel with FIXED parameters (computed once), the cross-Grams from the trainable base kernel."""
from __future__ import annotations

from ... import _lib_proxy as L
from ..._autograd import PredictVariance, tensor_leaves, wants_grad
from ...utils.device import as_2d
from ..base import KernelBase
from .base import ApproximateBaseKernel, ApproximateBaseKernelParameters
from ..standard.ard_kernel import ARDKernel


class FixedSparsePosteriorKernelParameters(ApproximateBaseKernelParameters):
    """base_kernel: parameters of the base kernel."""


class FixedSparsePosteriorKernel(ApproximateBaseKernel):
    Parameters = FixedSparsePosteriorKernelParameters

    def __init__(self, regulariser_kernel: KernelBase, regulariser_kernel_parameters, base_kernel: KernelBase,
                 inducing_points, diagonal_regularisation: float = 1e-5,
                 is_diagonal_regularisation_absolute_scale: bool = False, preprocess_function=None):
        assert isinstance(base_kernel, KernelBase) and isinstance(regulariser_kernel, KernelBase)
        # both ARD: the fused path (K tiles evaluated in the tile kernel, tensor-pipe gradients).  Otherwise the Grams
        # come from the kernels' own routines: frozen Cholesky from the regulariser kernel once, cross-Grams per chunk.
        self._ard = ARDKernel.is_fast_path(base_kernel) and ARDKernel.is_fast_path(regulariser_kernel)
        self.regulariser_kernel = regulariser_kernel
        self.regulariser_kernel_parameters = regulariser_kernel._to_parameters(regulariser_kernel_parameters)
        self.base_kernel = base_kernel
        self.inducing_points = as_2d(inducing_points)
        if self._ard:
            self.inducing_points = self.inducing_points.reshape(self.inducing_points.shape[0], -1)
        self.diagonal_regularisation = diagonal_regularisation
        self.is_diagonal_regularisation_absolute_scale = is_diagonal_regularisation_absolute_scale
        self._factor = None
        super().__init__(preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> FixedSparsePosteriorKernelParameters:
        return FixedSparsePosteriorKernel.Parameters(
            base_kernel=self.base_kernel._to_parameters(parameters["base_kernel"]))

    def factorise(self, parameters) -> "L.GpFactor":
        parameters = self._to_parameters(parameters)
        z = self.inducing_points
        if not self._ard:
            if self._factor is None:  # the regulariser kernel is frozen: one factorisation for the kernel's lifetime
                k_zz = self.regulariser_kernel.calculate_gram(self.regulariser_kernel_parameters, z, z)
                self._factor = L.GpFactor(z.shape[0], 1).factor_from_gram(
                    k_zz.detach().contiguous(), self.diagonal_regularisation,
                    self.is_diagonal_regularisation_absolute_scale)
            return self._factor
        if self._factor is None:
            self._factor = L.GpFactor(z.shape[0], z.shape[1])
        ls, ll = ARDKernel.device_parameters(self.base_kernel._to_parameters(parameters.base_kernel))
        cls_, cll = ARDKernel.device_parameters(self.regulariser_kernel_parameters)
        return self._factor.factor_frozen(z, ls, ll, cls_, cll, self.diagonal_regularisation,
                                          self.is_diagonal_regularisation_absolute_scale)

    def _calculate_gram_diagonal(self, parameters, x):
        f = self.factorise(parameters)
        if not self._ard:
            if wants_grad(*tensor_leaves(parameters.base_kernel)):
                _, var = L.gp_predict_any_kernel_grad(
                    f, self.base_kernel, parameters.base_kernel, self.inducing_points, x, self.diagonal_regularisation,
                    self.is_diagonal_regularisation_absolute_scale, want_mean=False, prefactored=True)
                return var
            _, var = L.gp_predict_any_kernel(f, self.base_kernel, parameters.base_kernel, self.inducing_points, x,
                                             want_mean=False)
            return var
        x = x.reshape(x.shape[0], -1)
        base = self.base_kernel._to_parameters(parameters.base_kernel)
        if wants_grad(base.log_scaling, base.log_lengthscales):
            # torch autograd front end: backward = gvi_gp_predict_grad_f64 (the variance VJP on the tensor pipe)
            return PredictVariance.apply(f, lambda: self.factorise(parameters), x, True, base.log_scaling,
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
