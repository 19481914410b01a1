"""Exact GP regression (reference: src/gps/gp_regression.py, src/gps/base/exact_base.py:16-49,
src/gps/base/base.py:384-526).

Posterior at x given (x_train, y_train):  C = K(x_train,x_train) + exp(log_noise) I,
  mean = k(x,x_train) C^-1 y_train + m(x)     (y is NOT centred by the prior mean, base.py:519-521, 492)
  var  = k(x,x) - k(x,x_train) C^-1 k(x_train,x)   (no observation noise added, base.py:522-525)
The K=1 predictive mean has shape [1,N] as in the reference (only the covariance is squeezed, base.py:410-412).
"""
from __future__ import annotations

from .. import _lib_proxy as L
from .._autograd import ExactPredict, tensor_leaves, wants_grad
from ..kernels.standard.ard_kernel import ARDKernel
from ..utils.device import as_2d, as_device, scalar
from .base import ExactGPBase, GPBaseParameters, GPRegressionBase


GPRegressionParameters = GPBaseParameters  # the reference's name for this GP's parameter container


class GPRegression(ExactGPBase, GPRegressionBase):
    def __init__(self, mean, kernel, x, y):
        # ARD: K tiles evaluated inside the tile kernel (and the differentiable exact-GP step).  Any other kernel: its own
        # Gram routine supplies K(x_train, x_train) and K(x_chunk, x_train) to the same factorisation / predictive kernels.
        self._ard = ARDKernel.is_fast_path(kernel)
        self.x = as_2d(x)
        if self._ard:
            self.x = self.x.reshape(self.x.shape[0], -1)
        self.y = as_device(y).reshape(-1)
        self._factor = None
        super().__init__(mean=mean, kernel=kernel)

    def factorise(self, parameters) -> "L.GpFactor":
        parameters = self._to_parameters(parameters)
        if not self._ard:
            if self._factor is None:
                self._factor = L.GpFactor(self.x.shape[0], 1)
            k_zz = self.kernel.calculate_gram(parameters.kernel, self.x, self.x).contiguous()
            return self._factor.factor_from_gram(k_zz, 0.0, True,
                                                 log_noise=scalar(parameters.log_observation_noise), y_z=self.y)
        if self._factor is None:
            self._factor = L.GpFactor(self.x.shape[0], self.x.shape[1])
        ls, ll = ARDKernel.device_parameters(parameters.kernel)
        return self._factor.factor(self.x, ls, ll, 0.0, True, log_noise=scalar(parameters.log_observation_noise),
                                   y_z=self.y)

    def calculate_posterior(self, parameters, x_train=None, y_train=None, x=None, full_covariance: bool = False):
        if full_covariance:
            raise NotImplementedError("full-covariance exact posterior is off the hot path (SURVEY.md section 8)")
        parameters = self._to_parameters(parameters)
        x = as_2d(x)
        if not self._ard:
            prior_mean = self.mean.predict(parameters.mean, x).reshape(-1).contiguous()
            if wants_grad(prior_mean, parameters.log_observation_noise, *tensor_leaves(parameters.kernel)):
                # training the exact GP over a non-ARD kernel (e.g. a feature-map kernel, train_regulariser_gp)
                if self._factor is None:
                    self._factor = L.GpFactor(self.x.shape[0], 1)
                mean, var = L.gp_predict_any_kernel_grad(
                    self._factor, self.kernel, parameters.kernel, self.x, x, 0.0, True, prior_mean=prior_mean,
                    log_noise=as_device(parameters.log_observation_noise), y_z=self.y)
                return mean.reshape(1, -1), var
            f = self.factorise(parameters)
            mean, var = L.gp_predict_any_kernel(f, self.kernel, parameters.kernel, self.x, x, prior_mean=prior_mean)
            return mean.reshape(1, -1), var
        x = x.reshape(x.shape[0], -1)
        f = self.factorise(parameters)
        prior_mean = self.mean.predict(parameters.mean, x).contiguous()
        kp = self.kernel._to_parameters(parameters.kernel)
        if wants_grad(prior_mean, parameters.log_observation_noise, kp.log_scaling, kp.log_lengthscales):
            # training the exact GP itself (train_regulariser, experiments/regression/trainers.py:19-85)
            noise = as_device(parameters.log_observation_noise)
            mean, var = ExactPredict.apply(f, lambda: self.factorise(parameters), x, prior_mean, noise,
                                           kp.log_scaling, kp.log_lengthscales)
        else:
            mean, var = L.gp_predict(f, x, prior_mean=prior_mean)
        return mean.reshape(1, -1), var

    def _calculate_prediction_gaussian(self, parameters, x, full_covariance: bool):
        return self.calculate_posterior(parameters, self.x, self.y, x, full_covariance)
