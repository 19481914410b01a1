"""Torch-tensor front end of the C ABI (include/gvigp.h).  Every function enqueues on torch's current stream and
never synchronises.  No CPU fallback: a missing library or a failed call raises."""
from __future__ import annotations

import torch

from .. import _lib
from .utils.device import device, empty

REG_KINDS = _lib.REG_KINDS


def _L():
    return _lib.load()


def ard_gram(x1, x2, log_scaling, log_ls, out, tensor_pipe: bool = True):
    nbytes = _L().gvi_ard_gram_workspace_bytes(x1.shape[0], x2.shape[0], x1.shape[1]) if tensor_pipe else 0
    ws = _workspace("gram", nbytes) if nbytes else None
    rc = _L().gvi_ard_gram_f64(_lib.ptr(x1), x1.shape[0], _lib.ptr(x2), x2.shape[0], x1.shape[1],
                               _lib.ptr(log_scaling), _lib.ptr(log_ls), log_ls.numel(), _lib.ptr(out),
                               out.stride(0) if out.ndim == 2 else x2.shape[0], _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_ard_gram_f64")


def ard_gram_diag(n, log_scaling, out):
    _lib.check(_L().gvi_ard_gram_diag_f64(n, _lib.ptr(log_scaling), _lib.ptr(out), _lib.current_stream()),
               "gvi_ard_gram_diag_f64")


KERNEL_ARD, KERNEL_POLYNOMIAL = 0, 1


def dot_gram(x1, x2, log_scaling, log_constant, degree, out):
    """out = (exp(log_scaling) x1 x2^T + exp(log_constant))^degree (log_constant None: inner-product kernel)."""
    rc = _L().gvi_dot_gram_f64(_lib.ptr(x1), x1.shape[0], _lib.ptr(x2), x2.shape[0], x1.shape[1],
                               _lib.ptr(log_scaling), _lib.ptr(log_constant), float(degree), _lib.ptr(out),
                               out.stride(0), _lib.current_stream())
    _lib.check(rc, "gvi_dot_gram_f64")


def kernel_pairs(kind, x1, x2, p0, p1, degree, out):
    """out[i] = k(x1_i, x2_i): the diagonal mode of KernelBase.calculate_gram for two different inputs."""
    rc = _L().gvi_kernel_pairs_f64(int(kind), _lib.ptr(x1), _lib.ptr(x2), x1.shape[0], x1.shape[1], _lib.ptr(p0),
                                   _lib.ptr(p1), p1.numel() if p1 is not None else 0, float(degree), _lib.ptr(out),
                                   _lib.current_stream())
    _lib.check(rc, "gvi_kernel_pairs_f64")


def add_diagonal_regulariser(a, diagonal_regularisation, is_absolute):
    _lib.check(_L().gvi_add_diagonal_regulariser_f64(_lib.ptr(a), a.shape[0], a.stride(0),
                                                     float(diagonal_regularisation), int(bool(is_absolute)),
                                                     _lib.current_stream()), "gvi_add_diagonal_regulariser_f64")


def chol_factor(a):
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    _lib.check(_L().gvi_chol_factor_f64(_lib.ptr(a), a.shape[0], a.stride(0), _lib.ptr(info),
                                        _lib.current_stream()), "gvi_chol_factor_f64")
    return info


class GpFactor:
    """Device-resident factorisation state of one GP over an inducing set (see gvi_gp_factor_f64)."""

    def __init__(self, m: int, d: int):
        self.m, self.d = m, d
        self.mp = (m + 31) // 32 * 32
        self.buf = empty(_L().gvi_gp_factor_bytes(m, d) // 8)
        self.ws = empty(_L().gvi_gp_factor_workspace_bytes(m) // 8)
        self.info = torch.zeros(1, dtype=torch.int32, device=device())
        self.version = 0  # bumped by every (re)factorisation; the autograd wrappers use it to detect stale state

    def factor(self, z, log_scaling, log_ls, diagonal_regularisation, is_absolute, log_noise=None, y_z=None):
        rc = _L().gvi_gp_factor_f64(_lib.ptr(z), self.m, self.d, _lib.ptr(log_scaling), _lib.ptr(log_ls),
                                    log_ls.numel(), float(diagonal_regularisation), int(bool(is_absolute)),
                                    _lib.ptr(log_noise), _lib.ptr(y_z), _lib.ptr(self.buf), _lib.ptr(self.ws),
                                    _lib.ptr(self.info), _lib.current_stream())
        _lib.check(rc, "gvi_gp_factor_f64")
        self.version += 1
        return self

    def factor_from_gram(self, k_zz, diagonal_regularisation, is_absolute, log_noise=None, y_z=None):
        """State from a caller-supplied K_ZZ (any base kernel); the buffer must have been sized with d = 1."""
        assert self.d == 1 and k_zz.shape == (self.m, self.m)
        rc = _L().gvi_gp_factor_gram_f64(_lib.ptr(k_zz), k_zz.stride(0), self.m, float(diagonal_regularisation),
                                         int(bool(is_absolute)), _lib.ptr(log_noise), _lib.ptr(y_z),
                                         _lib.ptr(self.buf), _lib.ptr(self.ws), _lib.ptr(self.info),
                                         _lib.current_stream())
        _lib.check(rc, "gvi_gp_factor_gram_f64")
        self.version += 1
        return self

    def factor_frozen(self, z, log_scaling, log_ls, chol_log_scaling, chol_log_ls, diagonal_regularisation,
                      is_absolute):
        rc = _L().gvi_gp_factor_frozen_f64(_lib.ptr(z), self.m, self.d, _lib.ptr(log_scaling), _lib.ptr(log_ls),
                                           log_ls.numel(), _lib.ptr(chol_log_scaling), _lib.ptr(chol_log_ls),
                                           chol_log_ls.numel(), float(diagonal_regularisation),
                                           int(bool(is_absolute)), _lib.ptr(self.buf), _lib.ptr(self.ws),
                                           _lib.ptr(self.info), _lib.current_stream())
        _lib.check(rc, "gvi_gp_factor_frozen_f64")
        self.version += 1
        return self

    def set_extra(self, e_matrix):
        """Attach (or detach with None) the lower-triangular E of the SVGP family: var += ||E k||^2."""
        rc = _L().gvi_gp_factor_set_extra_f64(_lib.ptr(self.buf), self.m, self.d, _lib.ptr(e_matrix),
                                              e_matrix.stride(0) if e_matrix is not None else 0,
                                              _lib.current_stream())
        _lib.check(rc, "gvi_gp_factor_set_extra_f64")
        self.version += 1
        return self

    # views into the workspace (valid until the next factor() call) -- used by the full-covariance API only
    def cholesky(self):
        mw = (self.m + 63) // 64 * 64  # the factorisation works on 64 x 64 blocks (csrc/chol_coop.cu)
        return self.ws[: mw * mw].view(mw, mw)[: self.m, : self.m]

    def inverse_cholesky(self):
        mw = (self.m + 63) // 64 * 64
        return self.ws[mw * mw: 2 * mw * mw].view(mw, mw)[: self.m, : self.m]


def gp_predict(factor: GpFactor, x, prior_mean=None, log_noise_add=None, want_mean=True):
    n = x.shape[0]
    mean = empty(n) if want_mean else None
    var = empty(n)
    nbytes = _L().gvi_gp_predict_workspace_bytes_n(factor.m, factor.d, n)
    ws = _workspace("predict", nbytes) if nbytes else None
    rc = _L().gvi_gp_predict_f64(_lib.ptr(factor.buf), factor.m, factor.d, _lib.ptr(x), n, _lib.ptr(prior_mean),
                                 _lib.ptr(log_noise_add), _lib.ptr(mean), _lib.ptr(var), _lib.ptr(ws),
                                 _lib.current_stream())
    _lib.check(rc, "gvi_gp_predict_f64")
    return mean, var


def gp_predict_gram(factor: GpFactor, k_xz, k_diag, prior_mean=None, log_noise_add=None, mean_out=None,
                    var_out=None):
    """Predictive from caller-supplied Gram rows K(x, Z) [n, M] and k(x_i, x_i) [n] (any base kernel)."""
    n = k_xz.shape[0]
    nbytes = _L().gvi_gp_predict_workspace_bytes(factor.m, 1)
    ws = _workspace("predict", nbytes) if nbytes else None
    rc = _L().gvi_gp_predict_gram_f64(_lib.ptr(factor.buf), factor.m, _lib.ptr(k_xz), k_xz.stride(0),
                                      _lib.ptr(k_diag), n, _lib.ptr(prior_mean), _lib.ptr(log_noise_add),
                                      _lib.ptr(mean_out), _lib.ptr(var_out), _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_gp_predict_gram_f64")


GRAM_CHUNK_BYTES = 256 << 20  # K(x_chunk, Z) rows materialised at a time on the generic-kernel path


def gp_predict_any_kernel(factor: GpFactor, kernel, kernel_parameters, z, x, prior_mean=None, want_mean=True):
    """Predictive mean / variance over ANY base kernel: K(x_chunk, Z) and k(x_i, x_i) from the kernel's own Gram routine,
    <= GRAM_CHUNK_BYTES at a time, the rest in the tile kernel."""
    n = x.shape[0]
    mean = empty(n) if want_mean else None
    var = empty(n)
    rows = max(24, GRAM_CHUNK_BYTES // (8 * factor.m) // 24 * 24)
    for r0 in range(0, n, rows):
        r1 = min(n, r0 + rows)
        xc = x[r0:r1]
        k_xz = kernel.calculate_gram(kernel_parameters, xc, z).contiguous()
        k_diag = kernel.calculate_gram(kernel_parameters, xc, xc, full_covariance=False).reshape(-1).contiguous()
        gp_predict_gram(factor, k_xz, k_diag, prior_mean[r0:r1] if prior_mean is not None else None, None,
                        mean[r0:r1] if want_mean else None, var[r0:r1])
    return mean, var


def gp_predict_any_kernel_grad(factor: GpFactor, kernel, kernel_parameters, z, x, diagonal_regularisation, is_absolute,
                               prior_mean=None, log_noise=None, y_z=None, want_mean=True, prefactored=False):
    """The differentiable variant of gp_predict_any_kernel (hyper-parameters of a non-ARD kernel, feature-map weights,
    observation noise, prior mean): K_zz is built with the autograd graph attached, factorised once by the CUDA
    factorisation, and every chunk goes through _autograd.PredictGram."""
    from ._autograd import PredictGram

    if prefactored:
        # the factor state was built from OTHER, frozen kernel parameters (FixedSparsePosteriorKernel): K_zz gets no
        # gradient, only the cross-Grams of the trainable kernel do
        k_zz = torch.zeros(factor.m, factor.m, dtype=torch.float64, device=x.device)
    else:
        k_zz = kernel.calculate_gram(kernel_parameters, z, z)
        noise = None if log_noise is None else log_noise.detach().reshape(-1)[:1].contiguous()
        factor.factor_from_gram(k_zz.detach().contiguous(), diagonal_regularisation, is_absolute, log_noise=noise,
                                y_z=y_z)
    linv = factor.inverse_cholesky().clone()
    alpha = None if y_z is None else linv.T @ (linv @ y_z)
    n = x.shape[0]
    rows = max(24, GRAM_CHUNK_BYTES // (8 * factor.m) // 24 * 24)
    means, variances = [], []
    for r0 in range(0, n, rows):
        xc = x[r0:min(n, r0 + rows)]
        k_xz = kernel.calculate_gram(kernel_parameters, xc, z)
        k_diag = kernel.calculate_gram(kernel_parameters, xc, xc, full_covariance=False).reshape(-1)
        pm = prior_mean[r0:r0 + xc.shape[0]] if prior_mean is not None else None
        mean, var = PredictGram.apply(factor, linv, alpha, diagonal_regularisation, is_absolute, k_zz, k_xz, k_diag, pm,
                                      log_noise)
        means.append(mean)
        variances.append(var)
    return (torch.cat(means) if want_mean else None), torch.cat(variances)


_risk_ws = {}


def _workspace(key, nbytes):
    """Scratch buffer cached per (purpose, device, stream): launches on different streams never share scratch."""
    dev = device()
    slot = (key, dev, torch.cuda.current_stream().cuda_stream)
    buf = _risk_ws.get(slot)
    if buf is None or buf.numel() * 8 < nbytes:
        buf = empty((nbytes + 7) // 8)
        _risk_ws[slot] = buf
    return buf


_side_streams = {}


def run_concurrently(jobs, n_streams: int = 4):
    """Run independent callables round-robin on a small pool of side streams and join them into the current stream.
    Used for the K per-class predictives of a multi-output GP: each is a latency-bound factorisation (a chain of small
    launches that fills a fraction of the GPU) followed by wide kernels, so neighbours overlap."""
    cur = torch.cuda.current_stream()
    dev = device()
    pool = _side_streams.setdefault(dev, [])
    while len(pool) < n_streams:
        pool.append(torch.cuda.Stream(device=dev))
    outs = []
    for i, job in enumerate(jobs):
        s = pool[i % n_streams]
        if i < n_streams:
            s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(job())
    for s in pool[: min(n_streams, len(outs))]:
        cur.wait_stream(s)
    for o in outs:
        for t in (o if isinstance(o, (tuple, list)) else (o,)):
            if isinstance(t, torch.Tensor):
                t.record_stream(cur)
    return outs


def risk(reg_kind, alpha, y, m_q, v_q, m_p, v_p):
    """Returns out3 = [sum nll, sum reg, N] (device tensor)."""
    n = m_q.numel()
    out3 = empty(3)
    ws = _workspace("risk", _L().gvi_risk_workspace_bytes(n))
    rc = _L().gvi_risk_f64(int(reg_kind), float(alpha), _lib.ptr(y), _lib.ptr(m_q), _lib.ptr(v_q), _lib.ptr(m_p),
                           _lib.ptr(v_p), n, _lib.ptr(out3), _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_risk_f64")
    return out3


def fused_step(factor_q: GpFactor, factor_p, x, y, m_q, prior_mean_p, m_p_in, v_p_in, reg_kind, alpha,
               out3=None, var_q_out=None, mean_p_out=None, var_p_out=None):
    n = x.shape[0]
    if out3 is None:
        out3 = empty(3)
    ws = _workspace("step", _L().gvi_fused_step_workspace_bytes_d(n, factor_q.m, factor_q.d))
    rc = _L().gvi_fused_step_f64(_lib.ptr(factor_q.buf), _lib.ptr(factor_p.buf) if factor_p is not None else None,
                                 factor_q.m, factor_q.d, _lib.ptr(x), _lib.ptr(y), _lib.ptr(m_q),
                                 _lib.ptr(prior_mean_p), _lib.ptr(m_p_in), _lib.ptr(v_p_in), n, int(reg_kind),
                                 float(alpha), _lib.ptr(out3), _lib.ptr(var_q_out), _lib.ptr(mean_p_out),
                                 _lib.ptr(var_p_out), _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_fused_step_f64")
    return out3


def fused_step_grad(factor_q: GpFactor, factor_p, x, y, m_q, prior_mean_p, m_p_in, v_p_in, reg_kind, alpha,
                    frozen_cholesky: bool = False):
    """Returns (out[4+D] = [sum nll, sum reg, N, d/dlog_scaling, d/dlog_ls[0..D)], dm[N]) in sum form."""
    n, d = x.shape
    out = empty(4 + d)
    dm = empty(n)
    ws = _workspace("grad", _L().gvi_fused_step_grad_workspace_bytes(n, factor_q.m, d))
    rc = _L().gvi_fused_step_grad_f64(_lib.ptr(factor_q.buf), _lib.ptr(factor_p.buf) if factor_p is not None else None,
                                      factor_q.m, d, _lib.ptr(x), _lib.ptr(y), _lib.ptr(m_q), _lib.ptr(prior_mean_p),
                                      _lib.ptr(m_p_in), _lib.ptr(v_p_in), n, int(reg_kind), float(alpha),
                                      int(bool(frozen_cholesky)), _lib.ptr(out), _lib.ptr(dm), None, None, None,
                                      _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_fused_step_grad_f64")
    return out, dm


def svgp_step_grad(factor_q: GpFactor, factor_p, x, y, m_q, prior_mean_p, m_p_in, v_p_in, reg_kind, alpha, e_matrix):
    """SVGP family: returns (out3 = [sum nll, sum reg, N], dm[N], dE[M,M]) in sum form (kernel parameters frozen)."""
    n, d = x.shape
    out3, dm = empty(3), empty(n)
    d_e = empty(factor_q.m, factor_q.m)
    ws = _workspace("grad", _L().gvi_fused_step_grad_workspace_bytes(n, factor_q.m, d))
    rc = _L().gvi_svgp_step_grad_f64(_lib.ptr(factor_q.buf), _lib.ptr(factor_p.buf) if factor_p is not None else None,
                                     factor_q.m, d, _lib.ptr(x), _lib.ptr(y), _lib.ptr(m_q), _lib.ptr(prior_mean_p),
                                     _lib.ptr(m_p_in), _lib.ptr(v_p_in), n, int(reg_kind), float(alpha),
                                     _lib.ptr(e_matrix), e_matrix.stride(0), _lib.ptr(out3), _lib.ptr(dm),
                                     _lib.ptr(d_e), _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_svgp_step_grad_f64")
    return out3, dm, d_e


def cond_var_select(x_perm, m_sel, log_scaling, log_ls, jitter):
    n, d = x_perm.shape
    idx = torch.empty(m_sel, dtype=torch.int64, device=x_perm.device)
    trace = empty(m_sel - 1)
    nbytes = _L().gvi_select_workspace_bytes(n, m_sel, d)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x_perm.device)
    rc = _L().gvi_cond_var_select_f64(_lib.ptr(x_perm), n, d, m_sel, _lib.ptr(log_scaling), _lib.ptr(log_ls),
                                      log_ls.numel(), float(jitter), _lib.ptr(idx), _lib.ptr(trace), _lib.ptr(ws),
                                      _lib.current_stream())
    _lib.check(rc, "gvi_cond_var_select_f64")
    return idx, trace


def cond_var_select_virtual_shards(x_perm, m_sel, log_scaling, log_ls, jitter, bounds):
    """The sharded selector's kernels and exchange protocol over len(bounds)-1 VIRTUAL shards of one device's points
    (gvi_cond_var_select_virtual_shards_f64).  Returns idx [R, m_sel] (every shard's view of the pivots) and
    trace [R, m_sel - 1]."""
    import numpy as np

    n, d = x_perm.shape
    r = len(bounds) - 1
    b = np.ascontiguousarray(bounds, dtype=np.int64)
    idx = torch.empty((r, m_sel), dtype=torch.int64, device=x_perm.device)
    trace = torch.empty((r, m_sel - 1), dtype=torch.float64, device=x_perm.device)
    nbytes = _L().gvi_select_virtual_shards_workspace_bytes(n, m_sel, d, r)
    assert nbytes > 0, "unsupported sizes"
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x_perm.device)
    rc = _L().gvi_cond_var_select_virtual_shards_f64(_lib.ptr(x_perm), n, d, m_sel, _lib.ptr(log_scaling),
                                                     _lib.ptr(log_ls), log_ls.numel(), float(jitter), r,
                                                     b.ctypes.data, _lib.ptr(idx), _lib.ptr(trace), _lib.ptr(ws),
                                                     _lib.current_stream())
    _lib.check(rc, "gvi_cond_var_select_virtual_shards_f64")
    return idx, trace


def threefry_random_bits(key, n, device):
    """jax.random.bits(key, (n,), uint32) as an int64 device tensor (gvi_threefry_random_bits)."""
    out = torch.empty(n, dtype=torch.int64, device=device)
    rc = _L().gvi_threefry_random_bits(int(key[0]), int(key[1]), n, _lib.ptr(out), _lib.current_stream())
    _lib.check(rc, "gvi_threefry_random_bits")
    return out


def cond_var_select_any_kernel(x_perm, m_sel, kernel, kernel_parameters, jitter):
    """The greedy selector over ANY kernel: k(x_i, x_i) and, per pivot, the column k(X, x_j) come from the kernel's own
    Gram routine (the pivot position stays on the device: x_j is gathered with a device index, no host sync); the
    pivoted-Cholesky update, rounding, clipping and the tie rules run in the same CUDA kernels as the ARD selector."""
    n = x_perm.shape[0]
    x_flat = x_perm.reshape(n, -1).contiguous()
    d = x_flat.shape[1]
    lib = _L()
    idx = torch.empty(m_sel, dtype=torch.int64, device=x_perm.device)
    trace = empty(m_sel - 1)
    ws = torch.empty(lib.gvi_select_workspace_bytes(n, m_sel, d), dtype=torch.uint8, device=x_perm.device)
    rl = lib.gvi_select_record_len(d, m_sel)
    record, winner = empty(rl), empty(rl)
    st = _lib.current_stream
    k_diag = kernel.calculate_gram(kernel_parameters, x_perm, x_perm, full_covariance=False).reshape(-1).contiguous()
    _lib.check(lib.gvi_select_init_gram_f64(_lib.ptr(x_flat), n, 0, d, m_sel, _lib.ptr(k_diag), float(jitter),
                                            _lib.ptr(record), _lib.ptr(ws), st()), "gvi_select_init_gram_f64")
    for m in range(m_sel - 1):
        _lib.check(lib.gvi_select_resolve_f64(_lib.ptr(record), 1, d, m_sel, int(m == 0), _lib.ptr(winner),
                                              idx[m:].data_ptr(), st()), "gvi_select_resolve_f64")
        x_j = x_perm.index_select(0, idx[m:m + 1])
        k_col = kernel.calculate_gram(kernel_parameters, x_j, x_perm).reshape(-1).contiguous()  # k is symmetric
        _lib.check(lib.gvi_select_update_gram_f64(_lib.ptr(x_flat), n, 0, d, m_sel, m, _lib.ptr(winner),
                                                  _lib.ptr(k_col), float(jitter), _lib.ptr(record),
                                                  trace[m:].data_ptr(), _lib.ptr(ws), st()),
                   "gvi_select_update_gram_f64")
    _lib.check(lib.gvi_select_resolve_f64(_lib.ptr(record), 1, d, m_sel, 0, _lib.ptr(winner),
                                          idx[m_sel - 1:].data_ptr(), st()), "gvi_select_resolve_f64")
    return idx, trace


# ---- vector-Jacobian products and the classification epilogue ---------------------------
def gp_predict_grad(factor: GpFactor, x, g_var, frozen_cholesky: bool = False):
    """out[4+D] with out[3] = dLoss/dlog_scaling and out[4:] = dLoss/dlog_lengthscales, given g_var = dLoss/dvar."""
    n, d = x.shape
    out = empty(4 + d)
    ws = _workspace("grad", _L().gvi_fused_step_grad_workspace_bytes(n, factor.m, d))
    rc = _L().gvi_gp_predict_grad_f64(_lib.ptr(factor.buf), factor.m, d, _lib.ptr(x), n, _lib.ptr(g_var),
                                      int(bool(frozen_cholesky)), _lib.ptr(out), _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_gp_predict_grad_f64")
    return out


def risk_grad(reg_kind, alpha, y, m_q, v_q, m_p, v_p, weights2):
    """(dm, dv): cotangents of the sums of `risk` pulled back to the approximate GP's mean and variance."""
    n = m_q.numel()
    dm, dv = empty(n), empty(n)
    rc = _L().gvi_risk_grad_f64(int(reg_kind), float(alpha), _lib.ptr(y), _lib.ptr(m_q), _lib.ptr(v_q),
                                _lib.ptr(m_p), _lib.ptr(v_p), n, _lib.ptr(weights2), _lib.ptr(dm), _lib.ptr(dv),
                                _lib.current_stream())
    _lib.check(rc, "gvi_risk_grad_f64")
    return dm, dv


def class_probabilities(mean, var, roots, weights, epsilon, cdf_lower_bound):
    """mean, var [K,N] -> probabilities [N,K] (gvi_class_probabilities_f64)."""
    k, n = mean.shape
    prob = empty(n, k)
    rc = _L().gvi_class_probabilities_f64(_lib.ptr(mean), _lib.ptr(var), k, n, _lib.ptr(roots), _lib.ptr(weights),
                                          roots.numel(), float(epsilon), float(cdf_lower_bound), _lib.ptr(prob),
                                          _lib.current_stream())
    _lib.check(rc, "gvi_class_probabilities_f64")
    return prob


def class_probabilities_grad(mean, var, roots, weights, epsilon, cdf_lower_bound, dprob):
    k, n = mean.shape
    dmean, dvar = empty(k, n), empty(k, n)
    rc = _L().gvi_class_probabilities_grad_f64(_lib.ptr(mean), _lib.ptr(var), k, n, _lib.ptr(roots),
                                               _lib.ptr(weights), roots.numel(), float(epsilon),
                                               float(cdf_lower_bound), _lib.ptr(dprob), _lib.ptr(dmean),
                                               _lib.ptr(dvar), _lib.current_stream())
    _lib.check(rc, "gvi_class_probabilities_grad_f64")
    return dmean, dvar


def multinomial_risk(prob_q, y, prob_p, power):
    """out4 = [sum NLL, sum Wasserstein, N, sum cross-entropy] (device tensor)."""
    n, k = prob_q.shape
    out4 = empty(4)
    ws = _workspace("mrisk", _L().gvi_multinomial_risk_workspace_bytes(n))
    rc = _L().gvi_multinomial_risk_f64(_lib.ptr(prob_q), _lib.ptr(y), _lib.ptr(prob_p), k, n, int(power),
                                       _lib.ptr(out4), _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_multinomial_risk_f64")
    return out4


def multinomial_risk_grad(prob_q, y, prob_p, power, weights4):
    n, k = prob_q.shape
    dprob = empty(n, k)
    rc = _L().gvi_multinomial_risk_grad_f64(_lib.ptr(prob_q), _lib.ptr(y), _lib.ptr(prob_p), k, n, int(power),
                                            _lib.ptr(weights4), _lib.ptr(dprob), _lib.current_stream())
    _lib.check(rc, "gvi_multinomial_risk_grad_f64")
    return dprob


def exact_gp_predict_grad(factor: GpFactor, x, g_var, h_mean, log_noise):
    """out[5+D]: out[3] = d/dlog_scaling, out[4:4+D] = d/dlog_lengthscales, out[4+D] = d/dlog_observation_noise of an
    exact-GP predictive (gvi_exact_gp_predict_grad_f64), given the cotangents of its variance and mean."""
    n, d = x.shape
    out = empty(5 + d)
    ws = _workspace("exact_grad", _L().gvi_exact_gp_predict_grad_workspace_bytes(n, factor.m, d))
    rc = _L().gvi_exact_gp_predict_grad_f64(_lib.ptr(factor.buf), factor.m, d, _lib.ptr(x), n, _lib.ptr(g_var),
                                            _lib.ptr(h_mean), _lib.ptr(log_noise), _lib.ptr(out), _lib.ptr(ws),
                                            _lib.current_stream())
    _lib.check(rc, "gvi_exact_gp_predict_grad_f64")
    return out


def weighted_gram(factor: GpFactor, x, g):
    """C[M,M] = sum_i g_i k(Z,x_i) k(x_i,Z) (gvi_weighted_gram_f64) - dLoss/dSigma of the SVGP family."""
    n, d = x.shape
    c = empty(factor.m, factor.m)
    ws = _workspace("grad", _L().gvi_fused_step_grad_workspace_bytes(n, factor.m, d))
    rc = _L().gvi_weighted_gram_f64(_lib.ptr(factor.buf), factor.m, d, _lib.ptr(x), n, _lib.ptr(g), _lib.ptr(c),
                                    _lib.ptr(ws), _lib.current_stream())
    _lib.check(rc, "gvi_weighted_gram_f64")
    return c
