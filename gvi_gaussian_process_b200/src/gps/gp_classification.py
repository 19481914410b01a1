"""Exact GP classification (reference: src/gps/gp_classification.py:32-94, src/gps/base/exact_base.py:16-49,
src/gps/base/base.py:384-526 vmapped over the K outputs): K independent exact posteriors - one CUDA factorisation
state per output, with that output's kernel parameters, observation noise and label column - feeding the
class-probability quadrature."""
from __future__ import annotations

import torch

from .. import _lib_proxy as L
from .._autograd import ExactPredict, wants_grad
from ..kernels.multi_output_kernel import MultiOutputKernel
from ..kernels.standard.ard_kernel import ARDKernel
from ..kernels.tempered_kernel import TemperedKernel
from ..utils.device import as_2d, as_device
from .base import ExactGPBase, GPBaseParameters, GPClassificationBase


GPClassificationParameters = GPBaseParameters  # the reference's name for this GP's parameter container


class GPClassification(ExactGPBase, GPClassificationBase):
    def __init__(self, mean, kernel, x, y, epsilon: float = 0.01, hermite_polynomial_order: int = 50,
                 cdf_lower_bound: float = 1e-10):
        base = kernel.base_kernel if isinstance(kernel, TemperedKernel) else kernel
        if not isinstance(base, MultiOutputKernel) or not all(ARDKernel.is_fast_path(k) for k in base.kernels):
            raise NotImplementedError("the B200 path implements exact GP classification over a MultiOutputKernel of "
                                      "ARDKernels without preprocess_function (optionally tempered)")
        self.x = as_2d(x)
        self.x = self.x.reshape(self.x.shape[0], -1)
        self.y = as_device(y).reshape(self.x.shape[0], -1)
        self._factors = None
        GPClassificationBase.__init__(self, mean=mean, kernel=kernel)
        self._init_classification(mean, kernel, epsilon, hermite_polynomial_order, cdf_lower_bound)

    def _member_parameters(self, parameters):
        """(ARD parameters, log tempering factor or None) per output."""
        k = self.number_output_dimensions
        if isinstance(self.kernel, TemperedKernel):
            members = self.kernel.base_kernel_parameters.kernels
            t = as_device(parameters.kernel.log_tempering_factor).reshape(-1)
            return [(members[i], t[i:i + 1] if t.numel() > 1 else t[:1]) for i in range(k)]
        return [(p, None) for p in parameters.kernel.kernels]

    def factorise(self, parameters):
        parameters = self._to_parameters(parameters)
        self._ensure_state()
        for i in range(self.number_output_dimensions):
            self._factor_member(parameters, i)
        return self._factors

    def _ensure_state(self):
        if self._factors is None:
            self._factors = [L.GpFactor(self.x.shape[0], self.x.shape[1]) for _ in range(self.number_output_dimensions)]
            self._y_cols = self.y.t().contiguous()  # [K, n] (base.py:480)

    def _member_leaves(self, parameters, i):
        """(log_scaling [1], log_lengthscales, log_noise [1]) of output i as (possibly differentiable) device tensors."""
        noise = as_device(parameters.log_observation_noise).reshape(-1)
        kp, temper = self._member_parameters(parameters)[i]
        ls, ll = ARDKernel.device_parameters(self.kernel_member(i)._to_parameters(kp))
        if temper is not None:
            ls = ls + 0.5 * temper  # exp(t) k = exp(ls + t/2)^2 exp(..): tempering scales the amplitude
        ni = noise[i:i + 1] if noise.numel() > 1 else noise[:1]
        return ls, ll, ni.contiguous()

    def _factor_member(self, parameters, i):
        ls, ll, ni = self._member_leaves(parameters, i)
        return self._factors[i].factor(self.x, ls, ll, 0.0, True, log_noise=ni, y_z=self._y_cols[i])

    def _predict_member(self, parameters, i, x, prior_mean_i):
        f = self._factor_member(parameters, i)
        ls, ll, ni = self._member_leaves(parameters, i)
        if wants_grad(prior_mean_i, ls, ll, ni):
            # training the exact classifier itself (experiments/classification/trainers.py:20-100): exact-GP VJP
            return ExactPredict.apply(f, lambda: self._factor_member(parameters, i), x, prior_mean_i, ni, ls, ll)
        return L.gp_predict(f, x, prior_mean=prior_mean_i)

    def kernel_member(self, i):
        base = self.kernel.base_kernel if isinstance(self.kernel, TemperedKernel) else self.kernel
        return base.kernels[i]

    def calculate_posterior(self, parameters, x_train=None, y_train=None, x=None, full_covariance: bool = False):
        if full_covariance:
            raise NotImplementedError("full-covariance exact posterior is off the hot path (SURVEY.md section 8)")
        parameters = self._to_parameters(parameters)
        x = as_2d(x)
        x = x.reshape(x.shape[0], -1)
        prior_mean = self.mean.predict(parameters.mean, x).reshape(self.number_output_dimensions, -1).contiguous()
        # K independent exact posteriors: factorisation + predictive of each output overlap on side streams
        self._ensure_state()
        jobs = [(lambda i=i: self._predict_member(parameters, i, x, prior_mean[i]))
                for i in range(self.number_output_dimensions)]
        outs = L.run_concurrently(jobs)
        return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])

    def _calculate_prediction_gaussian(self, parameters, x, full_covariance: bool):
        return self.calculate_posterior(parameters, self.x, self.y, x, full_covariance)
