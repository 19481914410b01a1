// grad_wide.cu - the variance vector-Jacobian product (gvi_gp_predict_grad_f64) for WIDE inputs, D > 32 (BASELINE config
// #4: D = 784, one log-lengthscale per pixel).  The tile kernel's gradient phase keeps a [T x D] x tile and per-lane
// D-vectors in registers, which stops at D = 32; for wide inputs the contraction with dK/dlog_ls[d] is rewritten as
// GEMMs so that it runs on the tensor pipe for every d at once.  With a_i = A^-1 k_i and
//   W_ij = g_i a_ij k_ij (N x M),  r = W 1,  c = W^T 1,  Q = W^T X^,   S = sum_i g_i a_i a_i^T,  V = S o K_ZZ,
//   T2_d    = sum_ij W_ij (x^_id - z^_jd)^2 = sum_i x^_id^2 r_i - 2 sum_j z^_jd Q_jd + sum_j z^_jd^2 c_j
//   term3_d = -1/2 sum_jl V_jl (z^_jd - z^_ld)^2 = -sum_j z^_jd^2 (V 1)_j + sum_j z^_jd (V Z^)_jd
//   dLoss/dlog_ls[d] = T2_d + term3_d,   dLoss/dlog_scaling = 2 sum g v - 2 jitter_abs sum g |a|^2
// (x^ = x~ - c~, z^ = z~ - c~: scaled coordinates shifted by the first inducing point, so the expansion of the square
// does not cancel for data far from the origin).  Per 2048-row chunk: K(x,Z) from the tensor-pipe Gram kernel
// (gram_dmma.cu), A^-1 K^T by two library GEMMs against the dense L^-1 of the factor state, one fused elementwise pass
// for W, diag(g) Aa and the row sums, Q (+ c as an extra column of ones) and S by library GEMMs accumulating across
// chunks.  All reductions have a fixed order.  Reference: jax.grad through sparse_posterior_kernel.py:54-96.
#include "dgemm.cuh"

#include "common.cuh"

int gvi_launch_wide_gram(const double* F, const FactorLayout& f, const double* x, int64_t nr, double* out,
                         double* workspace, cudaStream_t st);
int gvi_launch_wide_gram_zz(const double* F, const FactorLayout& f, double* out, double* workspace, cudaStream_t st);
size_t gvi_gram_dmma_workspace_doubles(int64_t m1, int64_t m2, int D);

namespace {
// rows of X per pass: 37 x 128, so that a chunk x M = 2048 product is 16 x 37 = 592 = 4 x 148 output tiles of the in-tree GEMM -
// whole waves of CTAs (2048 rows gave 256 tiles: 1.73 waves)
constexpr int WCHUNK = 4736;

// out[i][d] = rows[i][d] * (sqrtw ? sqrtw[d] : 1) - shift[d]  (d < D),  out[i][D] = 1;  leading dimension Da = D + 1
__global__ void __launch_bounds__(256) shift_aug_kernel(const double* __restrict__ rows, int64_t nr, int D,
                                                        const double* __restrict__ sqrtw,
                                                        const double* __restrict__ shift, double* __restrict__ out) {
  const int Da = D + 1;
  const int64_t total = nr * Da;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
    const int64_t i = idx / Da;
    const int d = (int)(idx % Da);
    out[idx] = (d == D) ? 1.0 : rows[i * D + d] * (sqrtw ? sqrtw[d] : 1.0) - shift[d];
  }
}

// one warp per row: W = g a k (over K), B = g a (over T), rowstats[i] = {sum_j W_ij, sum_j k_ij a_ij, sum_j a_ij^2}
__global__ void __launch_bounds__(256) weights_kernel(double* __restrict__ K, const double* __restrict__ Aa,
                                                      double* __restrict__ B, const double* __restrict__ g, int nr,
                                                      int Mp, double* __restrict__ rowstats) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < nr; i += gridDim.x * 8) {
    const double gi = g[i];
    double r = 0.0, ka = 0.0, na = 0.0;
    for (int j = lane; j < Mp; j += 32) {
      const int64_t idx = (int64_t)i * Mp + j;
      const double k = K[idx], a = Aa[idx];
      const double b = gi * a;
      const double w = b * k;
      K[idx] = w;
      B[idx] = b;
      r += w;
      ka = fma(k, a, ka);
      na = fma(a, a, na);
    }
    r = warp_sum(r);
    ka = warp_sum(ka);
    na = warp_sum(na);
    if (lane == 0) {
      rowstats[3 * i] = r;
      rowstats[3 * i + 1] = ka;
      rowstats[3 * i + 2] = na;
    }
  }
}

// t2x[d] (+)= sum_i x^_id^2 r_i: 32 features per CTA (coalesced across lanes), the rows split over the CTA's 16 warps
// in contiguous segments with two independent accumulators each, segments combined in a fixed order.  (The first version
// walked all rows with ONE thread per feature: 0.4 ms per 2048-row chunk of dependent loads - 240 ms of the 770 ms
// value-and-gradient step of config #4.)
constexpr int RSQ_WARPS = 16;
__global__ void __launch_bounds__(32 * RSQ_WARPS) row_square_kernel(const double* __restrict__ Xs, int nr, int D,
                                                                   const double* __restrict__ rowstats, int first,
                                                                   double* __restrict__ t2x) {
  __shared__ double part[RSQ_WARPS][32];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  const int per = (nr + RSQ_WARPS - 1) / RSQ_WARPS;
  const int i0 = seg * per, i1 = min(nr, i0 + per);
  double a0 = 0.0, a1 = 0.0;
  if (d < D) {
    int i = i0;
    for (; i + 1 < i1; i += 2) {
      const double x0 = Xs[(int64_t)i * (D + 1) + d], x1 = Xs[(int64_t)(i + 1) * (D + 1) + d];
      a0 = fma(x0 * x0, rowstats[3 * i], a0);
      a1 = fma(x1 * x1, rowstats[3 * (i + 1)], a1);
    }
    if (i < i1) {
      const double x0 = Xs[(int64_t)i * (D + 1) + d];
      a0 = fma(x0 * x0, rowstats[3 * i], a0);
    }
  }
  part[seg][lane] = a0 + a1;
  __syncthreads();
  if (seg == 0 && d < D) {
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < RSQ_WARPS; s++) acc += part[s][lane];
    t2x[d] = first ? acc : t2x[d] + acc;
  }
}

// scal[0] (+)= sum g v, scal[1] (+)= sum g |a|^2, scal[2] (+)= sum g;  v_i = k(x,x) - sum_j k_ij a_ij
__global__ void __launch_bounds__(256) row_scalars_kernel(const double* __restrict__ rowstats,
                                                          const double* __restrict__ g, int nr,
                                                          const double* __restrict__ F, int first,
                                                          double* __restrict__ scal) {
  __shared__ double red[3][256];
  const int tid = threadIdx.x;
  const double kxx = F[1];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = tid; i < nr; i += 256) {
    a = fma(g[i], kxx - rowstats[3 * i + 1], a);
    b = fma(g[i], rowstats[3 * i + 2], b);
    c += g[i];
  }
  red[0][tid] = a;
  red[1][tid] = b;
  red[2][tid] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s)
      for (int q = 0; q < 3; q++) red[q][tid] += red[q][tid + s];
    __syncthreads();
  }
  if (tid < 3) scal[tid] = first ? red[tid][0] : scal[tid] + red[tid][0];
}

__global__ void __launch_bounds__(256) hadamard_kernel(double* __restrict__ S, const double* __restrict__ K,
                                                       int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256)
    S[idx] *= K[idx];
}

// out[4+d] = t2x[d] + sum_j (z^_jd^2 c_j - 2 z^_jd Q_jd) + (R3 ? sum_j (z^_jd R3_jd - z^_jd^2 (V1)_j) : 0)
// `xz_scale` multiplies the x-z part: 1 when W already carries dLoss/dk_ij * (-1/2) * (-2) (variance VJP), -1/2 when W
// is dLoss/dk_ij o k_ij (exact-GP VJP)
__global__ void __launch_bounds__(128) wide_finish_kernel(const double* __restrict__ Zh, const double* __restrict__ Q,
                                                          const double* __restrict__ R3,
                                                          const double* __restrict__ t2x, int M, int D, double xz_scale,
                                                          double* __restrict__ out) {
  const int d = blockIdx.x * 128 + threadIdx.x;
  if (d >= D) return;
  const int Da = D + 1;
  double acc = t2x[d], acc3 = 0.0;
  for (int j = 0; j < M; j++) {
    const double z = Zh[(int64_t)j * Da + d];
    acc += z * (z * Q[(int64_t)j * Da + D] - 2.0 * Q[(int64_t)j * Da + d]);
    if (R3) acc3 += z * (R3[(int64_t)j * Da + d] - z * R3[(int64_t)j * Da + D]);
  }
  out[4 + d] = fma(xz_scale, acc, acc3);
}

__global__ void wide_scalars_kernel(const double* __restrict__ scal, const double* __restrict__ F, int64_t N,
                                    double* __restrict__ out) {
  out[0] = out[1] = 0.0;
  out[2] = (double)N;
  // same closed forms as finish_partials_kernel (predict.cu): A built from the trained parameters, or frozen
  out[3] = (F[5] != 0.0) ? 4.0 * scal[0] - 2.0 * F[1] * scal[2] : 2.0 * scal[0] - 2.0 * F[4] * scal[1];
}

struct WideWs {
  double *K, *T, *Aa, *Xs, *rowstats, *Q, *S, *Kzz, *Zh, *R3, *t2x, *scal, *gram;
};
size_t align256(size_t x) { return (x + 255) / 256 * 256; }
WideWs carve_wide(void* ws, int64_t N, int M, int D, size_t* total) {
  const size_t Mp = (size_t)(M + 31) / 32 * 32, Da = (size_t)D + 1;
  const size_t R = (size_t)(N < WCHUNK ? N : WCHUNK);
  unsigned char* p = (unsigned char*)ws;
  size_t o = 0;
  auto take = [&](size_t doubles) { double* r = (double*)(p + o); o += align256(doubles * 8); return r; };
  WideWs w;
  w.K = take(R * Mp);
  w.T = take(R * Mp);
  w.Aa = take(R * Mp);
  w.Xs = take(R * Da);
  w.rowstats = take(R * 3);
  w.Q = take(Mp * Da);
  w.S = take(Mp * Mp);
  w.Kzz = take(Mp * Mp);
  w.Zh = take(Mp * Da);
  w.R3 = take(Mp * Da);
  w.t2x = take(D);
  w.scal = take(4);
  w.gram = take(gvi_gram_dmma_workspace_doubles((int64_t)(R > Mp ? R : Mp), M, D));
  if (total) *total = o;
  return w;
}
}  // namespace

size_t gvi_gp_predict_grad_wide_workspace_bytes(int64_t N, int M, int D) {
  size_t total = 0;
  carve_wide(nullptr, N, M, D, &total);
  return total;
}

int gvi_gp_predict_grad_wide(const double* F, int M, int D, const double* X, int64_t N, const double* g, int frozen,
                             double* out, void* workspace, cudaStream_t st) {
  const FactorLayout f = factor_layout(M, D);
  const int Mp = f.Mp, Da = D + 1;
  const double* Linv = F + f.linv_off;
  const double* Zs = F + f.zs_off;  // row 0 = the common shift c~
  WideWs w = carve_wide(workspace, N, M, D, nullptr);
  int rc = GVI_OK;
  for (int64_t r0 = 0; r0 < N; r0 += WCHUNK) {
    const int nr = (int)(N - r0 < WCHUNK ? N - r0 : WCHUNK);
    const int first = (r0 == 0);
    const double beta = first ? 0.0 : 1.0;
    // the Gram kernel writes columns < M only: the padding columns of the GEMM operands must be zero
    if (Mp != M) GVI_CUDA(cudaMemsetAsync(w.K, 0, (size_t)nr * Mp * sizeof(double), st));
    if ((rc = gvi_launch_wide_gram(F, f, X + r0 * D, nr, w.K, w.gram, st))) return rc;
    shift_aug_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(X + r0 * D, nr, D, F + f.sqrtw_off, Zs, w.Xs);
    GVI_CHECK_LAUNCH("shift_aug_kernel");
    // in-tree tensor-pipe GEMM (dgemm.cu): triangular L^-1 -> half the k range; symmetric S -> half the tiles
    if ((rc = gvi_dgemm_rm(false, true, nr, Mp, Mp, w.K, Mp, Linv, Mp, 0.0, w.T, Mp, GVI_GEMM_K_UPTO_N, false, st)))
      return rc;  // K L^-T
    if ((rc = gvi_dgemm_rm(false, false, nr, Mp, Mp, w.T, Mp, Linv, Mp, 0.0, w.Aa, Mp, GVI_GEMM_K_FROM_N, false, st)))
      return rc;  // Aa = K A^-1
    const int wblocks = (nr + 7) / 8 < GVI_NUM_SMS * 8 ? (nr + 7) / 8 : GVI_NUM_SMS * 8;
    weights_kernel<<<wblocks, 256, 0, st>>>(w.K, w.Aa, w.T, g + r0, nr, Mp, w.rowstats);           // K <- W, T <- B
    GVI_CHECK_LAUNCH("weights_kernel");
    if ((rc = gvi_dgemm_rm(true, false, Mp, Da, nr, w.K, Mp, w.Xs, Da, beta, w.Q, Da, GVI_GEMM_K_FULL, false, st)))
      return rc;  // Q|c += W^T [X^|1]
    if (!frozen)
      if ((rc = gvi_dgemm_rm(true, false, Mp, Mp, nr, w.Aa, Mp, w.T, Mp, beta, w.S, Mp, GVI_GEMM_K_FULL, true, st)))
        return rc;  // S += Aa^T B (B = diag(g) Aa: symmetric)
    row_square_kernel<<<(D + 31) / 32, 32 * RSQ_WARPS, 0, st>>>(w.Xs, nr, D, w.rowstats, first, w.t2x);
    GVI_CHECK_LAUNCH("row_square_kernel");
    row_scalars_kernel<<<1, 256, 0, st>>>(w.rowstats, g + r0, nr, F, first, w.scal);
    GVI_CHECK_LAUNCH("row_scalars_kernel");
  }
  shift_aug_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(Zs, Mp, D, nullptr, Zs, w.Zh);
  GVI_CHECK_LAUNCH("shift_aug_kernel");
  const double* R3 = nullptr;
  if (!frozen) {
    GVI_CUDA(cudaMemsetAsync(w.Kzz, 0, (size_t)Mp * Mp * sizeof(double), st));
    if ((rc = gvi_launch_wide_gram_zz(F, f, w.Kzz, w.gram, st))) return rc;
    hadamard_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(w.S, w.Kzz, (int64_t)Mp * Mp);                // V = S o K_ZZ
    GVI_CHECK_LAUNCH("hadamard_kernel");
    if ((rc = gvi_dgemm_rm(false, false, Mp, Da, Mp, w.S, Mp, w.Zh, Da, 0.0, w.R3, Da, GVI_GEMM_K_FULL, false, st)))
      return rc;  // [V Z^ | V 1]
    R3 = w.R3;
  }
  wide_finish_kernel<<<(D + 127) / 128, 128, 0, st>>>(w.Zh, w.Q, R3, w.t2x, M, D, 1.0, out);
  GVI_CHECK_LAUNCH("wide_finish_kernel");
  wide_scalars_kernel<<<1, 1, 0, st>>>(w.scal, F, N, out);
  GVI_CHECK_LAUNCH("wide_scalars_kernel");
  return GVI_OK;
}


// ---- exact-GP predictive VJP (mean and variance) for wide inputs: same contraction machinery ------------------------------
// dLoss/dK_xZ = G1 = h alpha^T - 2 diag(g) Aa,  dLoss/dA = G2 = Aa^T diag(g) Aa - w alpha^T (w = Aa^T h);  see exact.cu.
//   dLoss/dlog_ls[d]  = -1/2 sum_ij (G1 o K)_ij (x^_id - z^_jd)^2 - 1/2 sum_jl (sym(G2) o K_ZZ)_jl (z^_jd - z^_ld)^2
//   dLoss/dlog_scaling = 2 (sum G1 o K + sum G2 o K_ZZ + amp2 sum g),   dLoss/dlog_noise = s2 tr(G2)
namespace {
// one warp per row: B = g a (over T), W1 = (h_i alpha_j - 2 g_i a_ij) k_ij (over K), rowstats[i] = {sum_j W1_ij, -, -}
__global__ void __launch_bounds__(256) exact_weights_kernel(double* __restrict__ K, const double* __restrict__ Aa,
                                                            double* __restrict__ B, const double* __restrict__ g,
                                                            const double* __restrict__ h,
                                                            const double* __restrict__ alpha, int nr, int Mp,
                                                            double* __restrict__ rowstats) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < nr; i += gridDim.x * 8) {
    const double gi = g[i], hi = h ? h[i] : 0.0;
    double r = 0.0;
    for (int j = lane; j < Mp; j += 32) {
      const int64_t idx = (int64_t)i * Mp + j;
      const double b = gi * Aa[idx];
      const double w1 = fma(hi, alpha[j], -2.0 * b) * K[idx];
      K[idx] = w1;
      B[idx] = b;
      r += w1;
    }
    r = warp_sum(r);
    if (lane == 0) {
      rowstats[3 * i] = r;
      rowstats[3 * i + 1] = 0.0;
      rowstats[3 * i + 2] = 0.0;
    }
  }
}

// scal[0] (+)= sum_i r_i, scal[2] (+)= sum_i g_i
__global__ void __launch_bounds__(256) exact_row_scalars_kernel(const double* __restrict__ rowstats,
                                                                const double* __restrict__ g, int nr, int first,
                                                                double* __restrict__ scal) {
  __shared__ double red[2][256];
  const int tid = threadIdx.x;
  double a = 0.0, c = 0.0;
  for (int i = tid; i < nr; i += 256) {
    a += rowstats[3 * i];
    c += g[i];
  }
  red[0][tid] = a;
  red[1][tid] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) {
      red[0][tid] += red[0][tid + s];
      red[1][tid] += red[1][tid + s];
    }
    __syncthreads();
  }
  if (tid == 0) {
    scal[0] = first ? red[0][0] : scal[0] + red[0][0];
    scal[2] = first ? red[1][0] : scal[2] + red[1][0];
  }
}

// G2 <- G2 - w alpha^T in place; tr_out[0] = tr(G2) (rows < M), one block
__global__ void __launch_bounds__(256) rank1_trace_kernel(double* __restrict__ G2, const double* __restrict__ wv,
                                                          const double* __restrict__ alpha, int M, int Mp,
                                                          double* __restrict__ tr_out) {
  __shared__ double red[256];
  const int64_t total = (int64_t)Mp * Mp;
  for (int64_t idx = threadIdx.x; idx < total; idx += 256) G2[idx] = fma(-wv[idx / Mp], alpha[idx % Mp], G2[idx]);
  __syncthreads();
  double a = 0.0;
  for (int j = threadIdx.x; j < M; j += 256) a += G2[(int64_t)j * Mp + j];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) tr_out[0] = red[0];
}

// V = 1/2 (G2 + G2^T) o K_ZZ into Vout; vsum[0] = sum V (fixed order: one block)
__global__ void __launch_bounds__(256) sym_hadamard_kernel(const double* __restrict__ G2, const double* __restrict__ Kzz,
                                                           int Mp, double* __restrict__ Vout,
                                                           double* __restrict__ vsum) {
  __shared__ double red[256];
  const int64_t total = (int64_t)Mp * Mp;
  double a = 0.0;
  for (int64_t idx = threadIdx.x; idx < total; idx += 256) {
    const int64_t r = idx / Mp, c = idx % Mp;
    const double v = 0.5 * (G2[idx] + G2[c * Mp + r]) * Kzz[idx];
    Vout[idx] = v;
    a += v;
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) vsum[0] = red[0];
}

__global__ void exact_wide_scalars_kernel(const double* __restrict__ scal, const double* __restrict__ F, int64_t N,
                                          int D, const double* __restrict__ log_noise, double* __restrict__ out) {
  out[0] = out[1] = 0.0;
  out[2] = (double)N;
  // scal: [0] sum G1 o K, [1] tr(G2), [2] sum g, [3] sum sym(G2) o K_ZZ
  out[3] = 2.0 * (scal[0] + scal[3] + F[1] * scal[2]);
  out[4 + D] = log_noise ? exp(log_noise[0]) * scal[1] : 0.0;
}
}  // namespace

size_t gvi_exact_gp_predict_grad_wide_workspace_bytes(int64_t N, int M, int D) {
  // carve_wide + one extra Mp x Mp matrix (V) + w[Mp]
  const size_t Mp = (size_t)(M + 31) / 32 * 32;
  return gvi_gp_predict_grad_wide_workspace_bytes(N, M, D) + align256(Mp * Mp * 8) + align256(Mp * 8);
}

int gvi_exact_gp_predict_grad_wide(const double* F, int M, int D, const double* X, int64_t N, const double* g,
                                   const double* h_mean, const double* log_noise, double* out, void* workspace,
                                   cudaStream_t st) {
  const FactorLayout f = factor_layout(M, D);
  const int Mp = f.Mp, Da = D + 1;
  const double* Linv = F + f.linv_off;
  const double* alpha = F + f.alpha_off;
  const double* Zs = F + f.zs_off;
  size_t base = 0;
  WideWs w = carve_wide(workspace, N, M, D, &base);
  double* V = (double*)((unsigned char*)workspace + base);
  double* wv = (double*)((unsigned char*)workspace + base + align256((size_t)Mp * Mp * 8));
  int rc = GVI_OK;
  GVI_CUDA(cudaMemsetAsync(wv, 0, (size_t)Mp * sizeof(double), st));
  for (int64_t r0 = 0; r0 < N; r0 += WCHUNK) {
    const int nr = (int)(N - r0 < WCHUNK ? N - r0 : WCHUNK);
    const int first = (r0 == 0);
    const double beta = first ? 0.0 : 1.0;
    if (Mp != M) GVI_CUDA(cudaMemsetAsync(w.K, 0, (size_t)nr * Mp * sizeof(double), st));
    if ((rc = gvi_launch_wide_gram(F, f, X + r0 * D, nr, w.K, w.gram, st))) return rc;
    shift_aug_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(X + r0 * D, nr, D, F + f.sqrtw_off, Zs, w.Xs);
    GVI_CHECK_LAUNCH("shift_aug_kernel");
    if ((rc = gvi_dgemm_rm(false, true, nr, Mp, Mp, w.K, Mp, Linv, Mp, 0.0, w.T, Mp, GVI_GEMM_K_UPTO_N, false, st)))
      return rc;
    if ((rc = gvi_dgemm_rm(false, false, nr, Mp, Mp, w.T, Mp, Linv, Mp, 0.0, w.Aa, Mp, GVI_GEMM_K_FROM_N, false, st)))
      return rc;
    const int wblocks = (nr + 7) / 8 < GVI_NUM_SMS * 8 ? (nr + 7) / 8 : GVI_NUM_SMS * 8;
    exact_weights_kernel<<<wblocks, 256, 0, st>>>(w.K, w.Aa, w.T, g + r0, h_mean ? h_mean + r0 : nullptr, alpha, nr, Mp,
                                                  w.rowstats);                                     // K <- W1, T <- B
    GVI_CHECK_LAUNCH("exact_weights_kernel");
    if ((rc = gvi_dgemm_rm(true, false, Mp, Da, nr, w.K, Mp, w.Xs, Da, beta, w.Q, Da, GVI_GEMM_K_FULL, false, st)))
      return rc;  // Q|c += W1^T [X^|1]
    if ((rc = gvi_dgemm_rm(true, false, Mp, Mp, nr, w.Aa, Mp, w.T, Mp, beta, w.S, Mp, GVI_GEMM_K_FULL, true, st)))
      return rc;  // G2 += Aa^T B
    if (h_mean && (rc = gvi_dgemv_rm(true, Mp, nr, w.Aa, Mp, h_mean + r0, 1.0, wv, st))) return rc;
    row_square_kernel<<<(D + 31) / 32, 32 * RSQ_WARPS, 0, st>>>(w.Xs, nr, D, w.rowstats, first, w.t2x);
    GVI_CHECK_LAUNCH("row_square_kernel");
    exact_row_scalars_kernel<<<1, 256, 0, st>>>(w.rowstats, g + r0, nr, first, w.scal);
    GVI_CHECK_LAUNCH("exact_row_scalars_kernel");
  }
  rank1_trace_kernel<<<1, 256, 0, st>>>(w.S, wv, alpha, M, Mp, w.scal + 1);
  GVI_CHECK_LAUNCH("rank1_trace_kernel");
  shift_aug_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(Zs, Mp, D, nullptr, Zs, w.Zh);
  GVI_CHECK_LAUNCH("shift_aug_kernel");
  GVI_CUDA(cudaMemsetAsync(w.Kzz, 0, (size_t)Mp * Mp * sizeof(double), st));
  if ((rc = gvi_launch_wide_gram_zz(F, f, w.Kzz, w.gram, st))) return rc;
  sym_hadamard_kernel<<<1, 256, 0, st>>>(w.S, w.Kzz, Mp, V, w.scal + 3);
  GVI_CHECK_LAUNCH("sym_hadamard_kernel");
  if ((rc = gvi_dgemm_rm(false, false, Mp, Da, Mp, V, Mp, w.Zh, Da, 0.0, w.R3, Da, GVI_GEMM_K_FULL, false, st)))
    return rc;  // [V Z^ | V 1]
  wide_finish_kernel<<<(D + 127) / 128, 128, 0, st>>>(w.Zh, w.Q, w.R3, w.t2x, M, D, -0.5, out);
  GVI_CHECK_LAUNCH("wide_finish_kernel");
  exact_wide_scalars_kernel<<<1, 1, 0, st>>>(w.scal, F, N, D, log_noise, out);
  GVI_CHECK_LAUNCH("exact_wide_scalars_kernel");
  return GVI_OK;
}
