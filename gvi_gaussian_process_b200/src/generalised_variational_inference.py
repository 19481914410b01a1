"""GeneralisedVariationalInference (reference: src/generalised_variational_inference.py:18-119):
loss = empirical risk + regularisation (no weighting).

When the objective is the hot-path configuration - Gaussian NLL of an ApproximateGPRegression over a
SparsePosteriorKernel(ARD) plus a pointwise regulariser against an exact GPRegression(ARD) (posterior mode) or any
GP prior (prior mode) on the same GP - the whole step runs as ONE fused CUDA launch (gvi_fused_step_f64): K(x,Z)
tiles are built in shared memory for q and p, both triangular products run on the FP64 tensor pipe and the
pointwise loss terms are reduced in the epilogue.  Otherwise the two terms are evaluated separately (still CUDA).
"""
from __future__ import annotations

import torch

from . import _lib_proxy as L
from .empirical_risks.base import EmpiricalRiskBase
from .empirical_risks.negative_log_likelihood import NegativeLogLikelihood
from .gps.approximate_gp_regression import ApproximateGPRegression
from .gps.base.classification_base import GPClassificationBase
from .gps.gp_regression import GPRegression
from .kernels.approximate.fixed_sparse_posterior_kernel import FixedSparsePosteriorKernel
from .kernels.approximate.sparse_posterior_kernel import SparsePosteriorKernel
from .kernels.approximate.svgp.svgp_kernels import LogSVGPKernel, SVGPBaseKernel
from .regularisations.base import RegularisationBase
from .regularisations.schemas import RegularisationMode
from .utils.device import as_2d, as_device


class GeneralisedVariationalInference:
    def __init__(self, regularisation: RegularisationBase, empirical_risk: EmpiricalRiskBase):
        self._regularisation = regularisation
        self._empirical_risk = empirical_risk
        self._side_stream = None

    def _factorise_regulariser(self, reg, reg_parameters, x, refactor: bool):
        """The two M x M factorisations (q and p) are latency-bound chains of small kernels that fill a fraction
        of the GPU, so the regulariser's runs on a side stream concurrently with q's (joined before the fused launch)."""
        cur = torch.cuda.current_stream()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        side = self._side_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            f_p = reg.factorise(reg_parameters) if (refactor or reg._factor is None) else reg._factor
            prior_mean_p = reg.mean.predict(reg_parameters.mean, x).detach().contiguous()
        prior_mean_p.record_stream(cur)  # allocated on the side stream, consumed by the fused launch on `cur`
        return f_p, prior_mean_p, side

    @property
    def regularisation(self):
        return self._regularisation

    @regularisation.setter
    def regularisation(self, regularisation):
        self._regularisation = regularisation

    @property
    def empirical_risk(self):
        return self._empirical_risk

    @empirical_risk.setter
    def empirical_risk(self, empirical_risk):
        self._empirical_risk = empirical_risk

    # ---------------------------------------------------------------------------------------------------------
    def _fusable(self) -> bool:
        er, rg = self._empirical_risk, self._regularisation
        if not isinstance(er, NegativeLogLikelihood) or not isinstance(rg, RegularisationBase):
            return False
        gp = er.gp
        if gp is not rg.gp or not isinstance(gp, ApproximateGPRegression) or isinstance(gp, GPClassificationBase):
            return False  # classification objectives take the general path (K predictives + quadrature epilogue)
        if getattr(gp.kernel, "factorise", None) is None or rg.kind == "none":
            return False  # SparsePosteriorKernel, FixedSparsePosteriorKernel and the SVGP kernels provide factorise()
        if not getattr(gp.kernel, "_ard", True):
            return False  # non-ARD base kernel (or one with a preprocess_function): term by term over supplied Grams
        if gp.kernel.has_preprocess_function:
            return False  # the wrapper kernel's own preprocess_function is applied by calculate_gram on the general path
        if rg.mode == RegularisationMode.posterior:
            reg = rg.regulariser
            return isinstance(reg, GPRegression) and reg._ard and reg.x.shape == gp.kernel.inducing_points.shape
        return True

    def loss_sums(self, parameters, x, y, refactor_regulariser: bool = True):
        """Shard-local sums [sum NLL_i, sum D0_i, N] of the fused step (device tensor, no host sync).
        Multi-GPU callers all-reduce this vector and divide (see gvi_gaussian_process_b200.distributed)."""
        er, rg = self._empirical_risk, self._regularisation
        gp = er.gp
        parameters = gp._to_parameters(parameters)
        x = as_2d(x)
        x = x.reshape(x.shape[0], -1)
        y = as_device(y).reshape(-1)
        assert y.numel() == x.shape[0], f"{y.numel()=} responses for {x.shape[0]=} points"
        kind = L.REG_KINDS[rg.kind]
        if rg.mode == RegularisationMode.posterior:
            reg = rg.regulariser
            reg_parameters = reg._to_parameters(rg.regulariser_parameters)
            f_p, prior_mean_p, side = self._factorise_regulariser(reg, reg_parameters, x, refactor_regulariser)
            m_q = gp.mean.predict(parameters.mean, x).contiguous()
            f_q = gp.kernel.factorise(parameters.kernel)
            torch.cuda.current_stream().wait_stream(side)
            return L.fused_step(f_q, f_p, x, y, m_q, prior_mean_p, None, None, kind, rg.alpha)
        m_q = gp.mean.predict(parameters.mean, x).contiguous()
        f_q = gp.kernel.factorise(parameters.kernel)
        m_p, c_p = rg.regulariser.calculate_prior(rg.regulariser_parameters, x, full_covariance=False)
        return L.fused_step(f_q, None, x, y, m_q, None, m_p.reshape(-1).contiguous(),
                            c_p.reshape(-1).contiguous(), kind, rg.alpha)

    # ---- gradient (what the reference obtains with jax.grad of calculate_loss, trainer.py:83-89) ----------------
    def loss_and_grad_sums(self, parameters, x, y, refactor_regulariser: bool = True):
        """Shard-local, SUM-form value and gradient of the fused step (device tensors, no host sync):
        returns (out, dm, m_q) with out = [sum NLL, sum D0, N, d/dlog_scaling, d/dlog_lengthscales[0..D)],
        dm[i] = d(sum loss)/dm_q[i] and m_q the mean-function output (so the host framework can back-propagate
        dm through it).  Multi-GPU: all-reduce `out`, keep dm sharded, divide both by out[2]."""
        assert self._fusable() and isinstance(self._empirical_risk.gp.kernel,
                                              (SparsePosteriorKernel, FixedSparsePosteriorKernel)), (
            "loss_and_grad_sums: SparsePosteriorKernel / FixedSparsePosteriorKernel over ARD (SVGP kernels: "
            "svgp_loss_and_grad_sums)")
        frozen = isinstance(self._empirical_risk.gp.kernel, FixedSparsePosteriorKernel)
        er, rg = self._empirical_risk, self._regularisation
        gp = er.gp
        parameters = gp._to_parameters(parameters)
        x = as_2d(x)
        x = x.reshape(x.shape[0], -1)
        y = as_device(y).reshape(-1)
        assert y.numel() == x.shape[0], f"{y.numel()=} responses for {x.shape[0]=} points"
        kind = L.REG_KINDS[rg.kind]
        side = None
        if rg.mode == RegularisationMode.posterior:
            reg = rg.regulariser
            reg_parameters = reg._to_parameters(rg.regulariser_parameters)
            f_p, prior_mean_p, side = self._factorise_regulariser(reg, reg_parameters, x, refactor_regulariser)
        m_q = gp.mean.predict(parameters.mean, x)
        m_q_c = m_q.detach().contiguous()
        f_q = gp.kernel.factorise(parameters.kernel)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            out, dm = L.fused_step_grad(f_q, f_p, x, y, m_q_c, prior_mean_p, None, None, kind, rg.alpha, frozen)
        else:
            m_p, c_p = rg.regulariser.calculate_prior(rg.regulariser_parameters, x, full_covariance=False)
            out, dm = L.fused_step_grad(f_q, None, x, y, m_q_c, None, m_p.detach().reshape(-1).contiguous(),
                                        c_p.detach().reshape(-1).contiguous(), kind, rg.alpha, frozen)
        return out, dm, m_q

    def svgp_loss_and_grad_sums(self, parameters, x, y, refactor_regulariser: bool = True):
        """SVGP kernels (hyper-parameters frozen, Sigma = E^T E trained): shard-local SUM-form
        (out3 = [sum NLL, sum D0, N], dm[N], m_q, dE[M,M]) with dE = 2 tril(E C)."""
        assert self._fusable() and isinstance(self._empirical_risk.gp.kernel, SVGPBaseKernel)
        er, rg = self._empirical_risk, self._regularisation
        gp = er.gp
        parameters = gp._to_parameters(parameters)
        x = as_2d(x)
        x = x.reshape(x.shape[0], -1)
        y = as_device(y).reshape(-1)
        kind = L.REG_KINDS[rg.kind]
        side = None
        if rg.mode == RegularisationMode.posterior:
            reg = rg.regulariser
            reg_parameters = reg._to_parameters(rg.regulariser_parameters)
            f_p, prior_mean_p, side = self._factorise_regulariser(reg, reg_parameters, x, refactor_regulariser)
        m_q = gp.mean.predict(parameters.mean, x)
        m_q_c = m_q.detach().contiguous()
        e_matrix = gp.kernel._el_matrix(parameters.kernel).detach().contiguous()
        f_q = gp.kernel.factorise(parameters.kernel)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            out3, dm, d_e = L.svgp_step_grad(f_q, f_p, x, y, m_q_c, prior_mean_p, None, None, kind, rg.alpha, e_matrix)
        else:
            m_p, c_p = rg.regulariser.calculate_prior(rg.regulariser_parameters, x, full_covariance=False)
            out3, dm, d_e = L.svgp_step_grad(f_q, None, x, y, m_q_c, None, m_p.detach().reshape(-1).contiguous(),
                                             c_p.detach().reshape(-1).contiguous(), kind, rg.alpha, e_matrix)
        return out3, dm, m_q, d_e

    def calculate_loss_and_gradients(self, parameters, x, y, allreduce=None):
        """loss and d loss / d {kernel.base_kernel.log_scaling, kernel.base_kernel.log_lengthscales, m_q}.
        Returns (loss, grads) with grads = {"log_scaling": 0-d, "log_lengthscales": same shape as the parameter,
        "mean_output": [N] = d loss / d m_q (feed to autograd of the mean function)}.
        `allreduce` (e.g. distributed.allreduce_loss_sums) sums the packed vector over ranks before normalising."""
        gp = self._empirical_risk.gp
        parameters = gp._to_parameters(parameters)
        out, dm, _ = self.loss_and_grad_sums(parameters, x, y)
        if allreduce is not None:
            out = allreduce(out)
        n = out[2]
        ll_shape = as_device(parameters.kernel.base_kernel.log_lengthscales).shape
        dll = out[4:] / n
        if as_device(parameters.kernel.base_kernel.log_lengthscales).numel() == 1:
            dll = dll.sum()  # a scalar lengthscale is broadcast over the D dimensions
        grads = {"log_scaling": out[3] / n, "log_lengthscales": dll.reshape(ll_shape), "mean_output": dm / n}
        return (out[0] + out[1]) / n, grads

    # the gradient tile kernel keeps the b tile of T >= 8 points in shared memory: M <= 3456; larger inducing sets are
    # differentiated term by term (variance VJP on the GEMM route of csrc/grad_wide.cu)
    FUSED_GRADIENT_MAX_M = 3456
    # ... and evaluates the kernel inside the tile: D <= 32 (wider inputs: K rows from the tensor-pipe Gram kernel; the
    # forward objective stays one call - gvi_fused_step_f64 walks chunks of <= 28 416 points - the gradient goes term by term)
    FUSED_GRADIENT_MAX_D = 32

    def _calculate_loss(self, parameters, x, y):
        if self._fusable():
            wants = torch.is_grad_enabled() and self._wants_grad(parameters)
            z_shape = self._empirical_risk.gp.kernel.inducing_points.shape
            if wants and (z_shape[0] > self.FUSED_GRADIENT_MAX_M or z_shape[1] > self.FUSED_GRADIENT_MAX_D):
                return self._calculate_loss_term_by_term(parameters, x, y)
            if not wants:
                out3 = self.loss_sums(parameters, x, y)
                return (out3[0] + out3[1]) / out3[2]
            if not isinstance(self._empirical_risk.gp.kernel, LogSVGPKernel):
                return _FusedLoss.apply(self, parameters, x, y, *self._grad_leaves(parameters))
            # LogSVGPKernel (dense El): trained through the general path below (dLoss/dSigma from the weighted-Gram
            # kernel, the M x M parameterisation differentiated by the host framework)
        return self._calculate_loss_term_by_term(parameters, x, y)

    def _calculate_loss_term_by_term(self, parameters, x, y):
        gp = self._empirical_risk.gp
        parameters = gp._to_parameters(parameters)
        x = as_device(x)  # one device copy, so both terms see the same tensor
        with gp.memoise():  # one predictive of q for both terms (XLA's CSE does this inside the reference's jit)
            return (self._empirical_risk.calculate_empirical_risk(parameters, x, y)
                    + self._regularisation.calculate_regularisation(parameters, x))

    # ---- torch autograd front end: `gvi.calculate_loss(...).backward()` plays the role of jax.grad ------------------
    def _grad_leaves(self, parameters):
        """Trainable kernel leaves - (log_scaling, log_lengthscales) of q's base kernel, or the E parameters of an
        SVGP kernel - followed by the tensors of the mean parameters."""
        gp = self._empirical_risk.gp
        parameters = gp._to_parameters(parameters)
        if isinstance(gp.kernel, SVGPBaseKernel):
            leaves = list(gp.kernel.grad_leaves(parameters.kernel))
        else:
            base = parameters.kernel.base_kernel
            leaves = [base.log_scaling, base.log_lengthscales]
        self._n_kernel_leaves = len(leaves)
        leaves += [t for t in _tensor_leaves(parameters.mean.dict()) if t.requires_grad]
        return leaves

    def _wants_grad(self, parameters) -> bool:
        if not isinstance(self._empirical_risk.gp.kernel,
                          (SparsePosteriorKernel, FixedSparsePosteriorKernel, SVGPBaseKernel)):
            return False
        return any(isinstance(t, torch.Tensor) and t.requires_grad for t in self._grad_leaves(parameters))

    def calculate_loss(self, parameters, x, y):
        """reference: src/generalised_variational_inference.py:96-119.  When any of q's kernel hyper-parameters or
        mean parameters is a torch tensor with requires_grad, the returned loss carries a backward that runs the
        CUDA gradient step (the role jax.grad plays in the reference's trainer)."""
        return self._calculate_loss(parameters, x, y)


def _tensor_leaves(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _tensor_leaves(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _tensor_leaves(v)


class _FusedLoss(torch.autograd.Function):
    """loss = (sum NLL + sum D0) / N of the fused step; backward = the CUDA gradient step, with dR/dm_q pushed
    through the (host-framework) mean function by torch autograd."""

    @staticmethod
    def forward(ctx, gvi, parameters, x, y, *leaves):
        gp = gvi._empirical_risk.gp
        nk = gvi._n_kernel_leaves
        kernel_leaves, mean_leaves = leaves[:nk], leaves[nk:]
        with torch.enable_grad():
            if isinstance(gp.kernel, SVGPBaseKernel):
                out, dm, m_q, d_e = gvi.svgp_loss_and_grad_sums(parameters, x, y)
                n = out[2]
                raw = gp.kernel.parameter_gradients(gp._to_parameters(parameters).kernel, d_e / n)
            else:
                out, dm, m_q = gvi.loss_and_grad_sums(parameters, x, y)
                n = out[2]
                d_ll = out[4:] / n
                ll = kernel_leaves[1]
                n_ll = ll.numel() if isinstance(ll, torch.Tensor) else 1
                raw = [out[3] / n, d_ll.sum() if n_ll == 1 else d_ll]  # a scalar lengthscale is broadcast over D
        kernel_grads = []
        for leaf, g in zip(kernel_leaves, raw):
            kernel_grads.append(g.detach().reshape(leaf.shape) if isinstance(leaf, torch.Tensor) else None)
        ctx.kernel_grads = kernel_grads
        ctx.mean_leaves = mean_leaves
        ctx.m_q = m_q
        ctx.save_for_backward((dm / n).detach())
        return ((out[0] + out[1]) / n).detach()

    @staticmethod
    def backward(ctx, grad_out):
        (d_m,) = ctx.saved_tensors
        kernel_grads = [None if g is None else grad_out * g for g in ctx.kernel_grads]
        mean_grads = [None] * len(ctx.mean_leaves)
        if ctx.mean_leaves and ctx.m_q.requires_grad:
            mean_grads = torch.autograd.grad(ctx.m_q, ctx.mean_leaves, grad_outputs=grad_out * d_m,
                                             allow_unused=True)
        return (None, None, None, None, *kernel_grads, *mean_grads)
