"""SVGP-family variational kernels (reference: src/kernels/approximate/svgp/base.py:24-73,
cholesky_svgp_kernel.py:75-150, diagonal_svgp_kernel.py:63-129).

All Grams use the regulariser kernel with FIXED parameters; the trainable part is Sigma = E^T E:
    r(x1,x2) = k(x1,x2) - k(x1,Z) (K_ZZ + jitter)^-1 k(Z,x2) + k(x1,Z) Sigma k(Z,x2)
  Cholesky parameterisation: E = tril(P, -1) + diag(exp(d));   diagonal: Sigma = diag(exp(d)), E = diag(exp(d/2)).
Diagonal mode (the training hot path) runs in the fused tile kernel: K(x,Z) tile once, two triangular products on
the FP64 tensor pipe (L^-1 k and E k), var = k(x,x) - ||L^-1 k||^2 + ||E k||^2.  The Cholesky of the fixed K_ZZ is
factorised once and cached."""
from __future__ import annotations

import torch

from .... import _lib_proxy as L
from ...._autograd import SVGPVariance, SVGPVarianceGram, wants_grad
from ....utils.device import as_2d, as_device, scalar
from ...base import KernelBase
from ..base import ApproximateBaseKernel, ApproximateBaseKernelParameters
from ...standard.ard_kernel import ARDKernel


class SVGPBaseKernelParameters(ApproximateBaseKernelParameters):
    pass


class SVGPBaseKernel(ApproximateBaseKernel):
    def __init__(self, regulariser_kernel: KernelBase, regulariser_kernel_parameters, log_observation_noise,
                 inducing_points, training_points, diagonal_regularisation: float = 1e-5,
                 is_diagonal_regularisation_absolute_scale: bool = False, preprocess_function=None):
        assert isinstance(regulariser_kernel, KernelBase)
        # ARD: K tiles evaluated inside the tile kernel.  Any other regulariser kernel: its own Gram routine supplies
        # K_ZZ (once) and K(x_chunk, Z); factorisation, both triangular products and the norms stay in the CUDA kernels.
        self._ard = ARDKernel.is_fast_path(regulariser_kernel)
        self.regulariser_kernel = regulariser_kernel
        self.regulariser_kernel_parameters = regulariser_kernel._to_parameters(regulariser_kernel_parameters)
        self.log_observation_noise = log_observation_noise
        self.inducing_points = as_2d(inducing_points)
        self.training_points = as_2d(training_points)
        if self._ard:
            self.inducing_points = self.inducing_points.reshape(self.inducing_points.shape[0], -1)
            self.training_points = self.training_points.reshape(self.training_points.shape[0], -1)
        self.number_of_dimensions = self.inducing_points.shape[1]
        self.diagonal_regularisation = diagonal_regularisation
        self.is_diagonal_regularisation_absolute_scale = is_diagonal_regularisation_absolute_scale
        self._factor = None
        super().__init__(preprocess_function=preprocess_function)

    # ---- one-time state (fixed regulariser kernel): svgp/base.py:57-73 ---------------------------------------
    def _frozen_factor(self) -> "L.GpFactor":
        if self._factor is None:
            z = self.inducing_points
            if not self._ard:
                k_zz = self.regulariser_kernel.calculate_gram(self.regulariser_kernel_parameters, z, z)
                self._factor = L.GpFactor(z.shape[0], 1).factor_from_gram(
                    k_zz.detach().contiguous(), self.diagonal_regularisation,
                    self.is_diagonal_regularisation_absolute_scale)
                return self._factor
            ls, ll = ARDKernel.device_parameters(self.regulariser_kernel_parameters)
            self._factor = L.GpFactor(z.shape[0], z.shape[1]).factor(
                z, ls, ll, self.diagonal_regularisation, self.is_diagonal_regularisation_absolute_scale)
        return self._factor

    def _variance_over_supplied_grams(self, f, x, sigma=None):
        """Non-ARD regulariser kernel: K(x_chunk, Z) and k(x, x) from the kernel, the tile kernel does the rest; with
        `sigma` (requires grad) every chunk goes through SVGPVarianceGram."""
        k, kp, z = self.regulariser_kernel, self.regulariser_kernel_parameters, self.inducing_points
        rows = max(24, L.GRAM_CHUNK_BYTES // (8 * f.m) // 24 * 24)
        out = []
        for r0 in range(0, x.shape[0], rows):
            xc = x[r0:r0 + rows]
            k_xz = k.calculate_gram(kp, xc, z).detach().contiguous()
            k_diag = k.calculate_gram(kp, xc, xc, full_covariance=False).detach().reshape(-1).contiguous()
            if sigma is not None:
                out.append(SVGPVarianceGram.apply(f, k_xz, k_diag, sigma))
            else:
                var = torch.empty(xc.shape[0], dtype=torch.float64, device=x.device)
                L.gp_predict_gram(f, k_xz, k_diag, None, None, None, var)
                out.append(var)
        return torch.cat(out)

    def _initial_inverse_cholesky(self) -> torch.Tensor:
        """inv(chol(K_ZZ + exp(-log_noise) K_ZX K_XZ)) - initialisation only, M^3 once (library linear algebra)."""
        k = self.regulariser_kernel
        kzz = k.calculate_gram(self.regulariser_kernel_parameters, self.inducing_points, self.inducing_points)
        kzx = k.calculate_gram(self.regulariser_kernel_parameters, self.inducing_points, self.training_points)
        precision = 1.0 / scalar(self.log_observation_noise).exp()
        chol = torch.linalg.cholesky(kzz + precision * (kzx @ kzx.T))
        return torch.linalg.inv(chol)

    def _el_matrix(self, parameters) -> torch.Tensor:
        raise NotImplementedError

    def factorise(self, parameters) -> "L.GpFactor":
        """State for the tile kernel: the cached frozen factor with this step's E attached."""
        parameters = self._to_parameters(parameters)
        f = self._frozen_factor()
        f.set_extra(self._el_matrix(parameters).contiguous())
        return f

    def _sigma_matrix(self, parameters) -> torch.Tensor:
        """Sigma = E^T E as a (host-framework differentiable) function of the trainable parameters."""
        e = self._el_matrix(parameters)
        return e.T @ e

    def _calculate_gram_diagonal(self, parameters, x):
        f = self.factorise(parameters)
        if not self._ard:
            sigma = self._sigma_matrix(parameters) if wants_grad(*self.grad_leaves(parameters)) else None
            return self._variance_over_supplied_grams(f, x, sigma)
        x = x.reshape(x.shape[0], -1)
        if wants_grad(*self.grad_leaves(parameters)):
            # general (non-fused) objectives, e.g. classification: dLoss/dSigma from the weighted-Gram kernel, the
            # M x M map parameters -> Sigma differentiated by the host framework
            return SVGPVariance.apply(f, x, self._sigma_matrix(parameters))
        _, var = L.gp_predict(f, x, want_mean=False)
        return var

    def _calculate_gram(self, parameters, x1, x2):
        f = self._frozen_factor()
        linv = f.inverse_cholesky()
        k = self.regulariser_kernel
        kp = self.regulariser_kernel_parameters
        k1z = k.calculate_gram(kp, x1, self.inducing_points)
        k2z = k.calculate_gram(kp, x2, self.inducing_points)
        k12 = k.calculate_gram(kp, x1, x2)
        return k12 - (k1z @ linv.T) @ (k2z @ linv.T).T + k1z @ self._sigma_matrix(parameters) @ k2z.T


class CholeskySVGPKernelParameters(SVGPBaseKernelParameters):
    """el_matrix_lower_triangle [M,M], el_matrix_log_diagonal [M]."""


class CholeskySVGPKernel(SVGPBaseKernel):
    Parameters = CholeskySVGPKernelParameters

    def initialise_el_matrix_parameters(self):
        """reference: cholesky_svgp_kernel.py:75-100."""
        inv_chol = self._initial_inverse_cholesky()
        lower = torch.tril(inv_chol, diagonal=-1)
        log_diag = torch.log(torch.clamp(torch.diagonal(inv_chol), min=self.diagonal_regularisation))
        return lower, log_diag

    def generate_parameters(self, parameters=None) -> CholeskySVGPKernelParameters:
        if parameters is None:
            lower, log_diag = self.initialise_el_matrix_parameters()
            return CholeskySVGPKernelParameters(el_matrix_lower_triangle=lower, el_matrix_log_diagonal=log_diag)
        return CholeskySVGPKernelParameters(el_matrix_lower_triangle=parameters["el_matrix_lower_triangle"],
                                            el_matrix_log_diagonal=parameters["el_matrix_log_diagonal"])

    def _el_matrix(self, parameters) -> torch.Tensor:
        lower = torch.tril(as_device(parameters.el_matrix_lower_triangle), diagonal=-1)
        return lower + torch.diag(torch.exp(as_device(parameters.el_matrix_log_diagonal).reshape(-1)))

    # gradient plumbing: E = tril(P,-1) + diag(exp(d))  =>  dP = tril(dE,-1), dd = diag(dE) exp(d)
    def grad_leaves(self, parameters):
        return [parameters.el_matrix_lower_triangle, parameters.el_matrix_log_diagonal]

    def parameter_gradients(self, parameters, d_e):
        d = as_device(parameters.el_matrix_log_diagonal).reshape(-1)
        return [torch.tril(d_e, diagonal=-1), torch.diagonal(d_e) * torch.exp(d)]


class DiagonalSVGPKernelParameters(SVGPBaseKernelParameters):
    """log_el_matrix_diagonal [M]."""


class DiagonalSVGPKernel(SVGPBaseKernel):
    Parameters = DiagonalSVGPKernelParameters

    def initialise_diagonal_parameters(self):
        """reference: diagonal_svgp_kernel.py:63-87 - diag(log(clip(L L^T))), L = inv(chol(...))."""
        inv_chol = self._initial_inverse_cholesky()
        sigma = inv_chol @ inv_chol.T
        return torch.log(torch.clamp(torch.diagonal(sigma), min=self.diagonal_regularisation))

    def generate_parameters(self, parameters=None) -> DiagonalSVGPKernelParameters:
        if parameters is None:
            return DiagonalSVGPKernelParameters(log_el_matrix_diagonal=self.initialise_diagonal_parameters())
        return DiagonalSVGPKernelParameters(log_el_matrix_diagonal=parameters["log_el_matrix_diagonal"])

    def _el_matrix(self, parameters) -> torch.Tensor:
        return torch.diag(torch.exp(0.5 * as_device(parameters.log_el_matrix_diagonal).reshape(-1)))

    # E = diag(exp(d/2))  =>  dd = diag(dE) * exp(d/2) / 2
    def grad_leaves(self, parameters):
        return [parameters.log_el_matrix_diagonal]

    def parameter_gradients(self, parameters, d_e):
        d = as_device(parameters.log_el_matrix_diagonal).reshape(-1)
        return [torch.diagonal(d_e) * 0.5 * torch.exp(0.5 * d)]


class LogSVGPKernelParameters(SVGPBaseKernelParameters):
    """log_el_matrix [M,M]."""


class LogSVGPKernel(SVGPBaseKernel):
    """reference: src/kernels/approximate/svgp/log_svgp_kernel.py:20-124 - El = exp(P) exp(P)^T (elementwise exp, dense
    and symmetric), Sigma = El^T El.  The tile kernel wants a LOWER-triangular E with E^T E = Sigma: the triangular
    factor of a QL factorisation of El, one M^3 library factorisation per step."""

    Parameters = LogSVGPKernelParameters

    def initialise_el_matrix_parameters(self):
        """reference lines 60-79: log(clip(inv(chol(K_ZZ + K_ZX K_XZ / noise)), diagonal_regularisation))."""
        return torch.log(torch.clamp(self._initial_inverse_cholesky(), min=self.diagonal_regularisation))

    def generate_parameters(self, parameters=None) -> LogSVGPKernelParameters:
        if parameters is None:
            return LogSVGPKernelParameters(log_el_matrix=self.initialise_el_matrix_parameters())
        return LogSVGPKernelParameters(log_el_matrix=parameters["log_el_matrix"])

    def _sigma_matrix(self, parameters) -> torch.Tensor:
        ex = torch.exp(as_device(parameters.log_el_matrix))
        el = ex @ ex.T
        return el.T @ el

    def _el_matrix(self, parameters) -> torch.Tensor:
        # "QL" factorisation El = Q Lo (Lo lower triangular) gives Sigma = El^T El = Lo^T Lo without squaring the
        # condition number (a Cholesky of Sigma breaks down for the reference's own initialisation, whose El spans
        # ~10 orders of magnitude).  QL through the library QR of the flipped matrix: J El J = Q R => Lo = J R J.
        ex = torch.exp(as_device(parameters.log_el_matrix).detach())
        el = ex @ ex.T
        r = torch.linalg.qr(torch.flip(el, dims=(0, 1)), mode="r").R
        return torch.flip(r, dims=(0, 1)).contiguous()

    def grad_leaves(self, parameters):
        return [parameters.log_el_matrix]

    def parameter_gradients(self, parameters, d_e):
        raise NotImplementedError("LogSVGPKernel trains through the general objective path (loss.backward())")


class KernelisedSVGPKernelParameters(SVGPBaseKernelParameters):
    """base_kernel: parameters of the (trainable) base kernel."""


class KernelisedSVGPKernel(SVGPBaseKernel):
    """reference: src/kernels/approximate/svgp/kernelised_svgp_kernel.py:20-99 -
    r = k_reg - k_reg(x,Z) (K_ZZ + jitter)^-1 k_reg(Z,x') + k_base(x,x'): the frozen sparse-posterior part comes from
    the cached factorisation, the trainable base kernel is added on top."""

    Parameters = KernelisedSVGPKernelParameters

    def __init__(self, base_kernel: KernelBase, regulariser_kernel: KernelBase, regulariser_kernel_parameters,
                 log_observation_noise, inducing_points, training_points, diagonal_regularisation: float = 1e-5,
                 is_diagonal_regularisation_absolute_scale: bool = False, preprocess_function=None):
        self.base_kernel = base_kernel
        super().__init__(regulariser_kernel=regulariser_kernel,
                         regulariser_kernel_parameters=regulariser_kernel_parameters,
                         log_observation_noise=log_observation_noise, inducing_points=inducing_points,
                         training_points=training_points, diagonal_regularisation=diagonal_regularisation,
                         is_diagonal_regularisation_absolute_scale=is_diagonal_regularisation_absolute_scale,
                         preprocess_function=preprocess_function)

    def generate_parameters(self, parameters) -> KernelisedSVGPKernelParameters:
        return KernelisedSVGPKernelParameters(base_kernel=self.base_kernel._to_parameters(parameters["base_kernel"]))

    # no extra matrix: the objective goes through the predict + risk kernels (not the single-launch fused step)
    factorise = None

    def _calculate_gram_diagonal(self, parameters, x):
        f = self._frozen_factor()
        f.set_extra(None)
        if not self._ard:
            var = self._variance_over_supplied_grams(f, x)
        else:
            _, var = L.gp_predict(f, x.reshape(x.shape[0], -1), want_mean=False)
        return var + self.base_kernel.calculate_gram(parameters.base_kernel, x, x, full_covariance=False)

    def grad_leaves(self, parameters):
        from ...._autograd import tensor_leaves

        return list(tensor_leaves(self.base_kernel._to_parameters(parameters.base_kernel)))

    def _calculate_gram(self, parameters, x1, x2):
        f = self._frozen_factor()
        linv = f.inverse_cholesky()
        k = self.regulariser_kernel
        kp = self.regulariser_kernel_parameters
        k1z = k.calculate_gram(kp, x1, self.inducing_points)
        k2z = k.calculate_gram(kp, x2, self.inducing_points)
        return (k.calculate_gram(kp, x1, x2) - (k1z @ linv.T) @ (k2z @ linv.T).T
                + self.base_kernel.calculate_gram(parameters.base_kernel, x1, x2))
