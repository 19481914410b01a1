"""NumPy/SciPy float64 restatement of the reference's GP hot path (TEST INFRASTRUCTURE, not product).

Every function cites the file:line of /root/reference it follows.  The reference is 100% Python on
jax==0.4.13 (x64), which is not installable here, so this restatement is the oracle; it is pinned against the
reference's own golden vectors in tests/test_oracle_goldens.py.  scipy's
cho_factor/cho_solve call the same LAPACK potrf/potrs that jaxlib-CPU calls.

Parity status: pinned for the ARD scalar, add_diagonal_regulariser, sparse-posterior Gram (identity kernel),
exact/approximate GP predictive, Gaussian NLL, all regularisers and the selector's tie/permutation behaviour.
UNPINNED (no reference test exists): ARD Grams with N,M>1 on real data, sparse-posterior variance with the ARD
kernel, gradients, the selector with a non-degenerate kernel.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_factor, cho_solve

LOG_2PI = float(np.log(2.0 * np.pi))


# --------------------------------------------------------------------------------------------------------------
# a1-a3: ARD kernel  (src/kernels/standard/ard_kernel.py:58-66, src/kernels/standard/base.py:95-103,
#                     src/kernels/base.py:104-147)
# --------------------------------------------------------------------------------------------------------------
def ard_kernel_scalar(log_scaling, log_lengthscales, x1, x2) -> float:
    """k(x1,x2) for single points, exactly as ard_kernel.py:58-66: exp(ls)**2 * exp(-0.5 * exp(ll) @ (x1-x2)^2)."""
    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = np.atleast_2d(np.asarray(x2, dtype=np.float64))
    scaling = np.exp(np.atleast_1d(np.float64(log_scaling))) ** 2
    w = np.atleast_1d(np.exp(np.asarray(log_lengthscales, dtype=np.float64)))
    return float(np.sum(scaling * np.exp(-0.5 * w @ np.square(x1 - x2).T)))


def ard_gram(log_scaling, log_lengthscales, x1, x2=None, chunk: int = 4096) -> np.ndarray:
    """Full Gram [m1,m2]; the double vmap of standard/base.py:95-103 evaluated with direct differences."""
    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = x1 if x2 is None else np.atleast_2d(np.asarray(x2, dtype=np.float64))
    x1 = x1.reshape(x1.shape[0], -1)
    x2 = x2.reshape(x2.shape[0], -1)
    assert x1.shape[1] == x2.shape[1]
    d = x1.shape[1]
    w = np.broadcast_to(np.atleast_1d(np.exp(np.asarray(log_lengthscales, dtype=np.float64))), (d,))
    scaling = np.exp(np.float64(log_scaling)) ** 2
    out = np.empty((x1.shape[0], x2.shape[0]), dtype=np.float64)
    for s in range(0, x1.shape[0], chunk):
        acc = np.zeros((min(chunk, x1.shape[0] - s), x2.shape[0]), dtype=np.float64)
        for k in range(d):  # same summation order as the w @ square(diff) dot for small d
            diff = x1[s : s + chunk, k : k + 1] - x2[None, :, k]
            acc += w[k] * (diff * diff)
        out[s : s + chunk] = scaling * np.exp(-0.5 * acc)
    return out


def ard_gram_blas(log_scaling, log_lengthscales, x1, x2=None, threads: int = 1, chunk: int = 8192) -> np.ndarray:
    """The same Gram through the dot-product expansion ||a||^2 + ||b||^2 - 2 a.b on sqrt(w)-scaled inputs (BLAS GEMM,
    exp over row chunks on `threads` threads).  NOT the form the reference evaluates (it takes direct differences, which
    `ard_gram` follows): this variant exists so that bench.py's CPU arm is not handicapped by a Python loop over D - a
    jit-compiled reference would fuse that loop.  Agrees with `ard_gram` to ~1e-13 on the bench's standardised inputs
    (tests/test_oracle_goldens.py); parity tests never use it."""
    from concurrent.futures import ThreadPoolExecutor

    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = x1 if x2 is None else np.atleast_2d(np.asarray(x2, dtype=np.float64))
    x1 = x1.reshape(x1.shape[0], -1)
    x2 = x2.reshape(x2.shape[0], -1)
    d = x1.shape[1]
    sw = np.sqrt(np.broadcast_to(np.atleast_1d(np.exp(np.asarray(log_lengthscales, dtype=np.float64))), (d,)))
    a, b = x1 * sw, x2 * sw
    na, nb = np.einsum("ij,ij->i", a, a), np.einsum("ij,ij->i", b, b)
    scaling = np.exp(np.float64(log_scaling)) ** 2
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)

    def rows(s):
        blk = out[s : s + chunk]
        np.matmul(a[s : s + chunk], b.T, out=blk)
        blk *= -2.0
        blk += na[s : s + chunk, None]
        blk += nb[None, :]
        np.maximum(blk, 0.0, out=blk)
        blk *= -0.5
        np.exp(blk, out=blk)
        blk *= scaling

    starts = range(0, a.shape[0], chunk)
    if threads > 1:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(rows, starts))
    else:
        for s in starts:
            rows(s)
    return out


def ard_gram_diag(log_scaling, log_lengthscales, x) -> np.ndarray:
    """full_covariance=False path of src/kernels/base.py:132-147: k(x_i,x_i) = exp(ls)**2 for every i."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    return np.full((x.shape[0],), np.exp(np.float64(log_scaling)) ** 2 * np.exp(-0.0), dtype=np.float64)


# --------------------------------------------------------------------------------------------------------------
# a4: src/utils/matrix_operations.py:7-28
# --------------------------------------------------------------------------------------------------------------
def inner_product_gram(log_scaling, x1, x2=None) -> np.ndarray:
    """InnerProductKernel through StandardKernelBase._calculate_gram: exp(log_scaling) * <x1_i, x2_j>
    (src/kernels/non_stationary/inner_product_kernel.py:36-42, src/kernels/standard/base.py:95-103)."""
    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = x1 if x2 is None else np.atleast_2d(np.asarray(x2, dtype=np.float64))
    return np.exp(log_scaling) * (x1.reshape(x1.shape[0], -1) @ x2.reshape(x2.shape[0], -1).T)


def polynomial_gram(log_constant, log_scaling, degree, x1, x2=None) -> np.ndarray:
    """PolynomialKernel: (exp(log_scaling) * <x1_i, x2_j> + exp(log_constant)) ** degree
    (src/kernels/non_stationary/polynomial_kernel.py:40-52).  jnp.power with a Python-int degree is
    lax.integer_pow (repeated multiplication), which is what numpy's ** does for an int exponent."""
    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = x1 if x2 is None else np.atleast_2d(np.asarray(x2, dtype=np.float64))
    base = np.exp(log_scaling) * (x1.reshape(x1.shape[0], -1) @ x2.reshape(x2.shape[0], -1).T) + np.exp(log_constant)
    return np.power(base, degree)


def gram_pairs(gram_fn, x1, x2) -> np.ndarray:
    """Diagonal mode of KernelBase.calculate_gram: k(x1_i, x2_i) by a vmap over the pairs (src/kernels/base.py:136-147)."""
    x1, x2 = np.atleast_2d(x1), np.atleast_2d(x2)
    return np.array([gram_fn(x1[i:i + 1], x2[i:i + 1])[0, 0] for i in range(x1.shape[0])])


def add_diagonal_regulariser(matrix, diagonal_regularisation: float, is_absolute: bool) -> np.ndarray:
    matrix = np.asarray(matrix, dtype=np.float64)
    m = matrix.shape[0]
    if is_absolute:
        return matrix + diagonal_regularisation * np.eye(m)
    return matrix + (diagonal_regularisation * np.trace(matrix) / m) * np.eye(m)


# --------------------------------------------------------------------------------------------------------------
# a5: SparsePosteriorKernel  (src/kernels/approximate/sparse_posterior_kernel.py:54-96)
# --------------------------------------------------------------------------------------------------------------
def sparse_posterior_gram(gram_fn, inducing_points, x1, x2=None, diagonal_regularisation=1e-5,
                          is_absolute=False) -> np.ndarray:
    """r(x1,x2) = k(x1,x2) - k(x1,Z) cho_solve(cho_factor(K_ZZ + jitter), k(Z,x2)); gram_fn(a,b) is the base Gram."""
    x2 = x1 if x2 is None else x2
    kzz = gram_fn(inducing_points, inducing_points)
    c_and_lower = cho_factor(add_diagonal_regulariser(kzz, diagonal_regularisation, is_absolute))
    k1z = gram_fn(x1, inducing_points)
    k2z = gram_fn(x2, inducing_points)
    k12 = gram_fn(x1, x2)
    return k12 - k1z @ cho_solve(c_and_lower, k2z.T)


def sparse_posterior_diag(gram_fn, diag_fn, inducing_points, x, diagonal_regularisation=1e-5,
                          is_absolute=False, chunk: int = 50000) -> np.ndarray:
    """Diagonal mode (src/kernels/base.py:132-147 vmaps 1x1 Grams): v_i = k(x_i,x_i) - k_i^T A^-1 k_i."""
    kzz = gram_fn(inducing_points, inducing_points)
    c_and_lower = cho_factor(add_diagonal_regulariser(kzz, diagonal_regularisation, is_absolute))
    n = x.shape[0]
    out = np.empty((n,), dtype=np.float64)
    kxx = diag_fn(x)
    for s in range(0, n, chunk):
        kxz = gram_fn(x[s : s + chunk], inducing_points)  # [c, M]
        sol = cho_solve(c_and_lower, kxz.T)  # [M, c]
        out[s : s + chunk] = kxx[s : s + chunk] - np.einsum("ij,ji->i", kxz, sol)
    return out


# --------------------------------------------------------------------------------------------------------------
# a6: approximate-GP predictive (src/gps/base/base.py:173-212,246-276; approximate_base.py:16,31-39;
#     regression_base.py:34-43).  The observation noise of an approximate GP is exp(log 0) = 0 on this path.
# --------------------------------------------------------------------------------------------------------------
def approximate_gp_predict(mean_values, kernel_diag, log_observation_noise=-np.inf):
    mean = np.asarray(mean_values, dtype=np.float64).reshape(-1)
    cov = np.asarray(kernel_diag, dtype=np.float64).reshape(-1) + np.exp(log_observation_noise)
    return mean, cov


# --------------------------------------------------------------------------------------------------------------
# a7: exact-GP posterior (src/gps/base/base.py:384-412, 414-493, 495-526)
# --------------------------------------------------------------------------------------------------------------
def exact_gp_posterior(gram_fn, diag_fn, prior_mean_values, x_train, y_train, log_observation_noise, x,
                       chunk: int = 50000):
    """mean = k(x,Z) C^-1 y + m(x)   (y NOT centred by the prior mean); var = k(x,x) - k(x,Z) C^-1 k(Z,x)
    (no observation noise added), C = K_ZZ + exp(log_noise) I.  Returns flat [N] arrays."""
    x_train = np.atleast_2d(x_train)
    y = np.asarray(y_train, dtype=np.float64).reshape(-1)
    ktt = gram_fn(x_train, x_train)
    c_and_lower = cho_factor(ktt + np.exp(log_observation_noise) * np.eye(ktt.shape[0]))
    alpha = cho_solve(c_and_lower, y)
    n = x.shape[0]
    mean = np.empty((n,), dtype=np.float64)
    var = np.empty((n,), dtype=np.float64)
    kxx = diag_fn(x)
    pm = np.broadcast_to(np.asarray(prior_mean_values, dtype=np.float64).reshape(-1), (n,))
    for s in range(0, n, chunk):
        ktx = gram_fn(x_train, x[s : s + chunk])  # [M, c]
        mean[s : s + chunk] = ktx.T @ alpha + pm[s : s + chunk]
        var[s : s + chunk] = kxx[s : s + chunk] - np.einsum("ji,ji->i", ktx, cho_solve(c_and_lower, ktx))
    return mean, var


def exact_gp_prior(diag_fn, prior_mean_values, log_observation_noise, x):
    """mode="prior" of src/regularisations/base.py:124-130 -> base.py:197-212: var = k(x,x) + exp(log_noise)."""
    n = x.shape[0]
    mean = np.broadcast_to(np.asarray(prior_mean_values, dtype=np.float64).reshape(-1), (n,)).copy()
    return mean, diag_fn(x) + np.exp(log_observation_noise)


# --------------------------------------------------------------------------------------------------------------
# a9: Gaussian NLL (src/empirical_risks/negative_log_likelihood.py:41-55)
# --------------------------------------------------------------------------------------------------------------
def gaussian_nll(y, mean, variance) -> float:
    y = np.atleast_2d(np.asarray(y, dtype=np.float64).reshape(np.atleast_2d(mean).shape[0], -1))
    m = np.atleast_2d(mean)
    v = np.atleast_2d(variance)
    scale = np.sqrt(v)
    logpdf = -0.5 * ((y - m) / scale) ** 2 - np.log(scale) - 0.5 * LOG_2PI  # jax.scipy.stats.norm.logpdf
    return float(np.mean(np.mean(-logpdf, axis=1)))


# --------------------------------------------------------------------------------------------------------------
# a10: projected regularisers (src/regularisations/projected/*.py)
# --------------------------------------------------------------------------------------------------------------
def d0_bhattacharyya(m_p, c_p, m_q, c_q):  # projected_bhattacharyya_regularisation.py:34-36
    return 0.125 * (m_p - m_q) ** 2 / (c_p + c_q) + 0.5 * np.log(((c_p + c_q) / 2) / np.sqrt(c_p * c_q))


def d0_gaussian_wasserstein(m_p, c_p, m_q, c_q):  # projected_gaussian_wasserstein_regularisation.py:34
    return (m_p - m_q) ** 2 + c_p + c_q - 2 * np.sqrt(c_p * c_q)


def d0_hellinger(m_p, c_p, m_q, c_q):  # projected_hellinger_regularisation.py:34-45
    return 1 - np.sqrt((2 * np.sqrt(c_p) * np.sqrt(c_q)) / (c_p + c_q)) * np.exp(
        -0.25 * (m_p - m_q) ** 2 / (c_p + c_q))


def d0_kl(m_p, c_p, m_q, c_q):  # projected_kl_regularisation.py:34-38
    return np.log(np.sqrt(c_q) / np.sqrt(c_p)) + (c_p + (m_p - m_q) ** 2) / (2 * c_q) - 0.5


def d0_renyi(m_p, c_p, m_q, c_q, alpha=0.5):  # projected_renyi_regularisation.py:36-45
    mix = alpha * c_p + (1 - alpha) * c_q
    return (np.log(np.sqrt(c_p) / np.sqrt(c_q)) + np.log(c_p / mix) / (2 * (alpha - 1))
            + 0.5 * alpha * (m_p - m_q) ** 2 / mix)


D0 = {
    "bhattacharyya": d0_bhattacharyya,
    "gaussian_wasserstein": d0_gaussian_wasserstein,
    "hellinger": d0_hellinger,
    "kl": d0_kl,
    "renyi": d0_renyi,
}


def projected_regularisation(kind, m_p, c_p, m_q, c_q, alpha=0.5) -> float:
    """projected/base.py:44-92: mean over outputs of mean over points of D0 on [K,N] arrays."""
    m_p, c_p, m_q, c_q = (np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in (m_p, c_p, m_q, c_q))
    if kind == "renyi":
        vals = d0_renyi(m_p, c_p, m_q, c_q, alpha)
    else:
        vals = D0[kind](m_p, c_p, m_q, c_q)
    return float(np.mean(np.mean(vals, axis=1)))


# --------------------------------------------------------------------------------------------------------------
# a11: GWI without eigendecomposition (gaussian_wasserstein_regularisation.py:143-160,230-251) and
#      squared difference with diagonal covariances (gaussian_squared_difference_regularisation.py:48-86)
# --------------------------------------------------------------------------------------------------------------
def gwi_no_eig(m_p, c_p, m_q, c_q) -> float:
    m_p, c_p, m_q, c_q = (np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in (m_p, c_p, m_q, c_q))
    per_output = np.mean((m_p - m_q) ** 2, axis=1) + np.mean(c_p, axis=1) + np.mean(c_q, axis=1)
    return float(np.mean(per_output))


def gwi_with_eig(m_p, cov_p, m_q, cov_q, gram_batch_train_p, gram_batch_train_q, eigenvalue_regularisation=0.0,
                 is_absolute=False) -> float:
    """Single-output full form (gaussian_wasserstein_regularisation.py:106-160, 55-104) - only used to pin the
    reference golden 0.87307723; the O(N^3) eigendecomposition variant is out of scope for the CUDA path."""
    n_b, n_t = gram_batch_train_p.shape
    # symmetric branch (lines 80-92): eigenvalues of (q + reg I)(p + reg I) via src/utils/matrix_operations.py:69-102
    prod = (add_diagonal_regulariser(gram_batch_train_q, eigenvalue_regularisation, is_absolute)
            @ add_diagonal_regulariser(gram_batch_train_p, eigenvalue_regularisation, is_absolute))
    ev = np.linalg.eigvals(prod)
    ev = np.where(np.real(ev) > 0, np.real(ev), 0.0)
    return float(np.mean((m_p - m_q) ** 2) + np.trace(cov_p) / n_b + np.trace(cov_q) / n_b
                 - 2.0 / np.sqrt(n_b * n_t) * np.sum(np.sqrt(ev)))


def gaussian_squared_difference(m_p, c_p, m_q, c_q) -> float:
    m_p, c_p, m_q, c_q = (np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in (m_p, c_p, m_q, c_q))
    per_output = np.mean((m_p - m_q) ** 2, axis=1) + np.mean(
        (c_p - c_q) ** 2, axis=tuple(range(1, c_p.ndim)))
    return float(np.mean(per_output))


# --------------------------------------------------------------------------------------------------------------
# a12: src/generalised_variational_inference.py:89-94
# --------------------------------------------------------------------------------------------------------------
def gvi_loss(empirical_risk: float, regularisation: float) -> float:
    return empirical_risk + regularisation


# --------------------------------------------------------------------------------------------------------------
# a13: greedy conditional-variance selector (src/inducing_points_selection.py:113-176)
# --------------------------------------------------------------------------------------------------------------
def conditional_variance_select(x_perm, number_of_inducing_points, gram_fn, diag_fn, jitter=1e-12,
                                threshold=0.0, use_argsort=True):
    """Runs on the ALREADY PERMUTED inputs and returns pivot positions into x_perm (the caller maps them back
    through perm, inducing_points_selection.py:173-176).  use_argsort=True follows the reference's
    reversed-stable-argsort scan literally; False uses an equivalent O(N) max/last-index rule (same result)."""
    m_total = number_of_inducing_points
    assert m_total > 1, "Must have at least 2 inducing points"
    n = x_perm.shape[0]
    indices = np.zeros(m_total, dtype=int) + n
    di = diag_fn(x_perm) + jitter
    indices[0] = int(np.argmax(di))
    ci = np.zeros((m_total - 1, n))
    for m in range(m_total - 1):
        j = int(indices[m])
        dj = np.sqrt(di[j])
        cj = ci[:m, j]
        g = np.round(np.squeeze(np.array(gram_fn(x_perm, x_perm[j : j + 1]))), 20).reshape(n)
        g[j] += jitter
        ei = (g - np.dot(cj, ci[:m])) / dj
        ci[m, :] = ei
        di = di - ei**2
        di = np.clip(di, 0, None)
        if use_argsort:
            chosen = set(int(i) for i in indices[: m + 1])
            for next_idx in reversed(np.argsort(di, kind="stable")):
                if int(next_idx) not in chosen:
                    indices[m + 1] = int(next_idx)
                    break
        else:
            masked = di.copy()
            masked[indices[: m + 1]] = -1.0
            mx = masked.max()
            indices[m + 1] = int(np.flatnonzero(masked == mx)[-1])
        if np.sum(np.clip(di, 0, None)) < threshold:
            indices = indices[:m]
            break
    return indices.astype(int)


# --------------------------------------------------------------------------------------------------------------
# section 8(f)-1: fixed sparse posterior and SVGP-family kernels
# --------------------------------------------------------------------------------------------------------------
def fixed_sparse_posterior_gram(base_gram_fn, regulariser_gram_fn, inducing_points, x1, x2=None,
                                diagonal_regularisation=1e-5, is_absolute=False) -> np.ndarray:
    """src/kernels/approximate/fixed_sparse_posterior_kernel.py:41-52, 78-99: Cholesky from the regulariser kernel
    (fixed), cross-Grams from the base kernel."""
    x2 = x1 if x2 is None else x2
    c_and_lower = cho_factor(add_diagonal_regulariser(regulariser_gram_fn(inducing_points, inducing_points),
                                                      diagonal_regularisation, is_absolute))
    k1z, k2z = base_gram_fn(x1, inducing_points), base_gram_fn(x2, inducing_points)
    return base_gram_fn(x1, x2) - k1z @ cho_solve(c_and_lower, k2z.T)


def svgp_initial_inverse_cholesky(gram_fn, inducing_points, training_points, log_observation_noise) -> np.ndarray:
    """svgp/base.py:57-73 + cholesky_svgp_kernel.py:86-95: inv(chol(K_ZZ + exp(-log_noise) K_ZX K_XZ))."""
    kzz = gram_fn(inducing_points, inducing_points)
    kzx = gram_fn(inducing_points, training_points)
    return np.linalg.inv(np.linalg.cholesky(kzz + (1.0 / np.exp(log_observation_noise)) * kzx @ kzx.T))


def svgp_cholesky_initial_parameters(gram_fn, inducing_points, training_points, log_observation_noise,
                                     diagonal_regularisation=1e-5):
    """cholesky_svgp_kernel.py:75-100."""
    inv_chol = svgp_initial_inverse_cholesky(gram_fn, inducing_points, training_points, log_observation_noise)
    return np.tril(inv_chol, k=-1), np.log(np.clip(np.diag(inv_chol), diagonal_regularisation, None))


def svgp_diagonal_initial_parameters(gram_fn, inducing_points, training_points, log_observation_noise,
                                     diagonal_regularisation=1e-5):
    """diagonal_svgp_kernel.py:63-87."""
    inv_chol = svgp_initial_inverse_cholesky(gram_fn, inducing_points, training_points, log_observation_noise)
    return np.diag(np.log(np.clip(inv_chol @ inv_chol.T, diagonal_regularisation, None)))


def svgp_gram(gram_fn, inducing_points, sigma_matrix, x1, x2=None, diagonal_regularisation=1e-5,
              is_absolute=False) -> np.ndarray:
    """cholesky_svgp_kernel.py:133-150 / diagonal_svgp_kernel.py:117-129 with sigma = el^T el or diag(exp(d))."""
    x2 = x1 if x2 is None else x2
    c_and_lower = cho_factor(add_diagonal_regulariser(gram_fn(inducing_points, inducing_points),
                                                      diagonal_regularisation, is_absolute))
    k1z, k2z = gram_fn(x1, inducing_points), gram_fn(x2, inducing_points)
    return gram_fn(x1, x2) - k1z @ cho_solve(c_and_lower, k2z.T) + k1z @ sigma_matrix @ k2z.T


def cholesky_sigma(el_matrix_lower_triangle, el_matrix_log_diagonal) -> np.ndarray:
    el = np.tril(el_matrix_lower_triangle, k=-1) + np.diag(np.exp(el_matrix_log_diagonal))
    return el.T @ el


# --------------------------------------------------------------------------------------------------------------
# section 8(f)-2: classification (multi-output GPs + Gauss-Hermite class probabilities)
# --------------------------------------------------------------------------------------------------------------
def multi_output_exact_gp_posterior(gram_fns, diag_fns, prior_mean_values, x_train, y_train, log_observation_noise,
                                    x):
    """K independent exact posteriors, one per output (src/gps/base/base.py:414-493 vmapped over outputs, 384-412 in
    diagonal mode): gram_fns[k] / diag_fns[k] are the k-th kernel of a MultiOutputKernel
    (src/kernels/multi_output_kernel.py:38-70), y_train is [n, K] and is transposed to [K, n] (base.py:480),
    log_observation_noise is [K] (construct_observation_noise_matrix, base.py:153-171).  Returns [K, N] arrays."""
    k_out = len(gram_fns)
    y_t = np.atleast_2d(np.asarray(y_train, dtype=np.float64).T)
    noise = np.broadcast_to(np.atleast_1d(np.asarray(log_observation_noise, dtype=np.float64)), (k_out,))
    pm = np.asarray(prior_mean_values, dtype=np.float64).reshape(k_out, -1)
    means, variances = [], []
    for k in range(k_out):
        m, v = exact_gp_posterior(gram_fns[k], diag_fns[k], pm[k], x_train, y_t[k], noise[k], x)
        means.append(m)
        variances.append(v)
    return np.stack(means), np.stack(variances)


def multi_output_prior(diag_fns, prior_mean_values, log_observation_noise, x):
    """ApproximateGPBase._calculate_prediction_gaussian -> calculate_prior (approximate_base.py:31-39,
    base.py:173-212): covariance [K, N] = diag k_k(x, x) + exp(log_observation_noise)[:, None]."""
    k_out = len(diag_fns)
    noise = np.broadcast_to(np.atleast_1d(np.exp(np.asarray(log_observation_noise, dtype=np.float64))), (k_out,))
    pm = np.asarray(prior_mean_values, dtype=np.float64).reshape(k_out, -1)
    cov = np.stack([diag_fns[k](x) for k in range(k_out)]) + noise[:, None]
    return pm.copy(), cov


def classification_s_matrix(means, covariance_diagonals, hermite_roots, hermite_weights, cdf_lower_bound):
    """GPClassificationBase._calculate_s_matrix (src/gps/base/classification_base.py:76-153): [K,N],[K,N] -> [N,K]."""
    from scipy.stats import norm

    means = np.asarray(means, dtype=np.float64)
    sd = np.sqrt(np.asarray(covariance_diagonals, dtype=np.float64))
    r = np.asarray(hermite_roots, dtype=np.float64)
    w = np.asarray(hermite_weights, dtype=np.float64)
    # (h, k_j, k_l, n)
    cdf_input = (np.sqrt(2.0) * r[:, None, None, None] * sd[None, :, None, :] + means[None, :, None, :]
                 - means[None, None, :, :]) / sd[None, None, :, :]
    log_cdf = np.log(np.clip(norm.cdf(cdf_input), cdf_lower_bound, None))
    diag = np.swapaxes(np.diagonal(log_cdf, axis1=1, axis2=2), 1, 2)  # (h, n, k) -> (h, k, n)
    comp = w[:, None, None] * np.exp(np.sum(log_cdf, axis=2) - diag)
    return ((1.0 / np.sqrt(np.pi)) * np.sum(comp, axis=0)).T


def classification_probabilities(means, covariance_diagonals, epsilon=0.01, hermite_polynomial_order=50,
                                 cdf_lower_bound=1e-10):
    """GPClassificationBase._predict_probability (classification_base.py:53-74) with the default constructor
    arguments of gp_classification.py:38-45 / approximate_gp_classification.py:30-36."""
    from scipy.special import roots_hermite

    roots, weights = roots_hermite(hermite_polynomial_order)
    s = classification_s_matrix(means, covariance_diagonals, roots, weights, cdf_lower_bound)
    k_out = np.asarray(means).shape[0]
    return (1.0 - epsilon) * s + (epsilon / (k_out - 1)) * (1.0 - s)


def multinomial_nll(probabilities, y) -> float:
    """negative_log_likelihood.py:57-78: -sum_i log sum_k p_ik y_ik  (a SUM over points, not a mean)."""
    return float(-np.sum(np.log(np.sum(np.asarray(probabilities) * np.asarray(y, dtype=np.float64), axis=1))))


def multinomial_wasserstein(probabilities_p, probabilities_q, power=2) -> float:
    """multinomial_wasserstein_regularisation.py:46-76: mean_i (sum_k |p_ik - q_ik|^power)^(1/power)."""
    d = np.abs(np.asarray(probabilities_p) - np.asarray(probabilities_q))
    return float(np.mean(np.power(np.sum(np.power(d, power), axis=1), 1.0 / power)))


def tempered_diag(diag_fn, log_tempering_factor):
    """src/kernels/tempered_kernel.py:72-81: exp(log_tempering_factor) * base Gram (per output)."""
    return lambda x: np.exp(log_tempering_factor) * diag_fn(x)


def cross_entropy_from_probabilities(probabilities, y) -> float:
    """cross_entropy.py:11-30: jax_metrics.losses.Crossentropy() (third-party, jax_metrics pinned in poetry.lock) is
    called with preds = the class probabilities and its default from_logits=True, i.e. the probabilities go through a
    log-softmax; pinned by the reference golden 1.306689 (tests/test_empirical_risks.py:215-269)."""
    p = np.asarray(probabilities, dtype=np.float64)
    logp = p - np.log(np.sum(np.exp(p), axis=1, keepdims=True))
    return float(np.mean(-np.sum(np.asarray(y, dtype=np.float64) * logp, axis=1)))


def log_svgp_initial_parameters(gram_fn, inducing_points, training_points, log_observation_noise,
                                diagonal_regularisation=1e-5):
    """log_svgp_kernel.py:60-79: log(clip(inv(chol(K_ZZ + K_ZX K_XZ / noise)), diagonal_regularisation))."""
    inv_chol = svgp_initial_inverse_cholesky(gram_fn, inducing_points, training_points, log_observation_noise)
    return np.log(np.clip(inv_chol, diagonal_regularisation, None))


def log_svgp_sigma(log_el_matrix) -> np.ndarray:
    """log_svgp_kernel.py:108-112: El = exp(P) exp(P)^T (elementwise exp), Sigma = El^T El."""
    ex = np.exp(np.asarray(log_el_matrix, dtype=np.float64))
    el = ex @ ex.T
    return el.T @ el


# --------------------------------------------------------------------------------------------------------------
# section 8(f)-4: optimiser rules of the reference's trainer (experiments/shared/resolvers/optimiser.py:6-16).
# optax 0.1.7 (third-party, pinned in poetry.lock:1916-1917, absent from /root/reference): published update rules of
# optax.adam / optax.adabelief / optax.rmsprop with their default hyper-parameters.  PARITY UNPINNED: the reference
# holds no golden for an optimiser step.
# --------------------------------------------------------------------------------------------------------------
def optimiser_step(kind, p, g, mu, nu, count, lr):
    """One step; returns (p, mu, nu).  count is the 1-based step index."""
    if kind == "adam":
        b1, b2, eps, eps_root = 0.9, 0.999, 1e-8, 0.0
        mu = b1 * mu + (1 - b1) * g
        nu = b2 * nu + (1 - b2) * g * g
        upd = (mu / (1 - b1**count)) / (np.sqrt(nu / (1 - b2**count) + eps_root) + eps)
    elif kind == "adabelief":
        b1, b2, eps, eps_root = 0.9, 0.999, 1e-16, 1e-16
        err = g - mu  # prediction error against the PREVIOUS first moment (optax.scale_by_belief)
        mu = b1 * mu + (1 - b1) * g
        nu = b2 * nu + (1 - b2) * err * err + eps_root
        upd = (mu / (1 - b1**count)) / (np.sqrt(nu / (1 - b2**count)) + eps)
    elif kind == "rmsprop":
        decay, eps = 0.9, 1e-8
        nu = decay * nu + (1 - decay) * g * g
        upd = g / np.sqrt(nu + eps)
    else:
        raise ValueError(kind)
    return p - lr * upd, mu, nu


def conformal_gp_coverage(mean_calibration, variance_calibration, y_calibration, mean, variance, coverage):
    """ConformalGPRegression.predict_coverage for one coverage percentage, from the predictives at the calibration and
    the query points (src/conformal_calibration/regression/conformal_gp_regression.py:27-71 for the Gaussian intervals,
    regression/base.py:76-137 for the calibration).  Parity unpinned: the reference has no test for this module."""
    from scipy.special import ndtri

    scale = ndtri((coverage + 1.0) / 2.0)
    lo_c = mean_calibration - scale * np.sqrt(variance_calibration)
    up_c = mean_calibration + scale * np.sqrt(variance_calibration)
    scores = np.maximum(lo_c - y_calibration, y_calibration - up_c)
    n = len(y_calibration)
    calibration = np.quantile(scores, np.clip((n + 1) * coverage / n, 0.0, 1.0))
    lower = mean - scale * np.sqrt(variance) - calibration
    upper = mean + scale * np.sqrt(variance) + calibration
    return np.minimum(lower, mean), np.maximum(upper, mean)
