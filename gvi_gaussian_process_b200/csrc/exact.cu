// exact.cu - SURVEY.md section 8(f)-3: the vector-Jacobian product of the EXACT GP predictive, i.e. the backward pass of
// the exact-GP NLL training step on the inducing set (experiments/regression/trainers.py:19-85: jax.grad of
// NegativeLogLikelihood over GPRegression.predict_probability, src/gps/base/base.py:384-526).
//
//   A = K_ZZ + s2 I (s2 = exp(log_noise)),  alpha = A^-1 y_Z,
//   mean_i = k_i^T alpha + m(x_i),   var_i = k(x_i,x_i) - k_i^T A^-1 k_i
// With h_i = dLoss/dmean_i, g_i = dLoss/dvar_i, a_i = A^-1 k_i, w = sum_i h_i a_i:
//   dLoss/dK_xZ = G1 = h alpha^T - 2 diag(g) Aa           (Aa rows = a_i^T)
//   dLoss/dA    = G2 = Aa^T diag(g) Aa - w alpha^T
//   dLoss/dtheta = <G1, dK_xZ/dtheta> + <G2, dK_ZZ/dtheta> + sum_i g_i dk_ii/dtheta,   dLoss/dlog s2 = s2 tr(G2)
// Unlike the approximate GP the mean depends on the hyper-parameters (through alpha and k_i), and the work is O(M^3)
// on at most a few thousand points (the inducing set itself), so this is the DENSE route: K(x_chunk, Z) is
// materialised chunk by chunk, the M^3-shaped products run on the in-tree tensor-pipe GEMM (dgemm.cu) against the dense
// L^-1 kept in the factor state, and hand-written kernels do the Gram chunk, the elementwise cotangents and the contractions
// with dK/dtheta (recomputing k_ij on the fly, fixed-order reductions).
#include "dgemm.cuh"

#include "common.cuh"

// wide inputs (D > 32): GEMM formulation of the contractions with dK/dlog_ls (grad_wide.cu)
size_t gvi_exact_gp_predict_grad_wide_workspace_bytes(int64_t N, int M, int D);
int gvi_exact_gp_predict_grad_wide(const double* F, int M, int D, const double* X, int64_t N, const double* g,
                                   const double* h_mean, const double* log_noise, double* out, void* workspace,
                                   cudaStream_t st);

namespace {
constexpr int ECHUNK = 4096;  // rows of X per pass

// K[i][j] = amp2 exp(-0.5 |x~_i - z~_j|^2) for j < M, 0 for the padding columns; x is scaled by sqrtw on the fly
__global__ void __launch_bounds__(256) kxz_kernel(const double* __restrict__ F, FactorLayout f,
                                                  const double* __restrict__ X, int nr, double* __restrict__ K) {
  const double amp2 = F[0];
  const double* __restrict__ sqrtw = F + f.sqrtw_off;
  const double* __restrict__ Zs = F + f.zs_off;
  const int64_t total = (int64_t)nr * f.Mp;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
    const int i = (int)(idx / f.Mp), j = (int)(idx % f.Mp);
    double s = 0.0;
    for (int d = 0; d < f.D; d++) {
      const double diff = X[(int64_t)i * f.D + d] * sqrtw[d] - Zs[(int64_t)j * f.D + d];
      s = fma(diff, diff, s);
    }
    K[idx] = (j < f.M) ? amp2 * gvi_exp(-0.5 * s) : 0.0;
  }
}

// B = diag(g) Aa,  G1 = h alpha^T - 2 B
__global__ void __launch_bounds__(256) cotangent_kernel(const double* __restrict__ Aa, const double* __restrict__ g,
                                                        const double* __restrict__ h,
                                                        const double* __restrict__ alpha, int nr, int Mp,
                                                        double* __restrict__ B, double* __restrict__ G1) {
  const int64_t total = (int64_t)nr * Mp;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
    const int i = (int)(idx / Mp), j = (int)(idx % Mp);
    const double b = g[i] * Aa[idx];
    B[idx] = b;
    G1[idx] = fma(h ? h[i] : 0.0, alpha[j], -2.0 * b);
  }
}

// G2 -= w alpha^T
__global__ void __launch_bounds__(256) rank1_sub_kernel(double* __restrict__ G2, const double* __restrict__ w,
                                                        const double* __restrict__ alpha, int Mp) {
  const int64_t total = (int64_t)Mp * Mp;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256)
    G2[idx] = fma(-w[idx / Mp], alpha[idx % Mp], G2[idx]);
}

// One CTA per row r of G [R x C, row stride ld]:  rowpart[r][d] = sum_c G_rc k_rc (-1/2)(row~_rd - col~_cd)^2 (d < D),
// rowpart[r][32] = sum_c G_rc k_rc (+ diag_w[r] amp2: the sum_i g_i k_ii term).  `rows` are raw inputs scaled by
// sqrtw on the fly when scale_rows != 0, otherwise already scaled (the Z~ of the factor state).
__global__ void __launch_bounds__(256) ard_contract_kernel(const double* __restrict__ G, int64_t ld, int C,
                                                           const double* __restrict__ rows, int scale_rows,
                                                           const double* __restrict__ F, FactorLayout f,
                                                           const double* __restrict__ diag_w,
                                                           double* __restrict__ rowpart) {
  __shared__ double red[8];
  __shared__ double xr[32];
  const double* __restrict__ Zs = F + f.zs_off;
  const int r = blockIdx.x, tid = threadIdx.x, D = f.D;
  if (tid < D) xr[tid] = rows[(int64_t)r * D + tid] * (scale_rows ? F[f.sqrtw_off + tid] : 1.0);
  __syncthreads();
  const double amp2 = F[0];
  double t[33];
#pragma unroll
  for (int d = 0; d < 33; d++) t[d] = 0.0;
  for (int c = tid; c < C; c += 256) {
    double s = 0.0, diff2[32];
#pragma unroll
    for (int d = 0; d < 32; d++) {
      if (d < D) {
        const double diff = xr[d] - Zs[(int64_t)c * D + d];
        diff2[d] = diff * diff;
        s += diff2[d];
      } else {
        diff2[d] = 0.0;
      }
    }
    const double gk = G[(int64_t)r * ld + c] * amp2 * gvi_exp(-0.5 * s);
    t[32] += gk;
    const double coef = -0.5 * gk;
#pragma unroll
    for (int d = 0; d < 32; d++) t[d] = fma(coef, diff2[d], t[d]);
  }
  for (int d = 0; d <= 32; d++) {
    if (d >= D && d < 32) continue;
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 33; q++) v = (q == d) ? t[q] : v;  // static indexing keeps t[] in registers
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double a = 0.0;
      for (int w = 0; w < 8; w++) a += red[w];
      if (d == 32 && diag_w) a = fma(diag_w[r], amp2, a);
      rowpart[(size_t)r * 33 + d] = a;
    }
  }
}

// fixed-order block sum (256 threads): thread t adds its strided share, then a shared-memory tree
__device__ __forceinline__ double block_sum_256(double v, double* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// acc[d] (+)= sum_r rowpart[r][d] in a fixed order; slot 32 is the <G, K> sum.  One CTA per slot d (33 CTAs).
__global__ void __launch_bounds__(256) contract_finish_kernel(const double* __restrict__ rowpart, int R, int D, int first,
                                                              double* __restrict__ acc) {
  __shared__ double red[256];
  const int d = blockIdx.x;
  if (d >= D && d < 32) return;
  double a = 0.0;
  for (int r = threadIdx.x; r < R; r += 256) a += rowpart[(size_t)r * 33 + d];
  a = block_sum_256(a, red);
  if (threadIdx.x == 0) acc[d] = first ? a : acc[d] + a;
}

// out[0..2] = {0, 0, N}; out[3] = 2 (<G1,K> + <G2,K_ZZ> + sum g k_ii); out[4+d] = both contractions;
// out[4+D] = s2 tr(G2)
__global__ void __launch_bounds__(256) exact_finish_kernel(const double* __restrict__ acc1, const double* __restrict__ acc2,
                                                           const double* __restrict__ G2, int M, int Mp, int D, int64_t N,
                                                           const double* __restrict__ log_noise, double* __restrict__ out) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  double tr = 0.0;
  for (int j = tid; j < M; j += 256) tr += G2[(size_t)j * Mp + j];
  tr = block_sum_256(tr, red);
  if (tid < D) out[4 + tid] = acc1[tid] + acc2[tid];
  if (tid == 32) {
    out[0] = out[1] = 0.0;
    out[2] = (double)N;
    out[3] = 2.0 * (acc1[32] + acc2[32]);
  }
  if (tid == 33) out[4 + D] = log_noise ? exp(log_noise[0]) * tr : 0.0;
}

struct ExactWs {
  double *K, *T, *Aa, *G2, *w, *rowpart, *acc1, *acc2;
};
size_t align256(size_t x) { return (x + 255) / 256 * 256; }
ExactWs carve_exact(void* ws, int64_t N, int M, size_t* total) {
  const size_t Mp = (size_t)(M + 31) / 32 * 32;
  const size_t nc = (size_t)(N < ECHUNK ? N : ECHUNK);
  unsigned char* p = (unsigned char*)ws;
  size_t o = 0;
  auto take = [&](size_t bytes) { double* r = (double*)(p + o); o += align256(bytes); return r; };
  ExactWs w;
  w.K = take(nc * Mp * 8);
  w.T = take(nc * Mp * 8);
  w.Aa = take(nc * Mp * 8);
  w.G2 = take(Mp * Mp * 8);
  w.w = take(Mp * 8);
  w.rowpart = take((nc > Mp ? nc : Mp) * 33 * 8);
  w.acc1 = take(33 * 8);
  w.acc2 = take(33 * 8);
  if (total) *total = o;
  return w;
}
}  // namespace

extern "C" size_t gvi_exact_gp_predict_grad_workspace_bytes(int64_t N, int M, int D) {
  if (N < 1 || M < 1 || D < 1) return 0;
  if (D > 32) return gvi_exact_gp_predict_grad_wide_workspace_bytes(N, M, D);
  size_t total = 0;
  carve_exact(nullptr, N, M, &total);
  return total;
}

extern "C" int gvi_exact_gp_predict_grad_f64(const void* factor, int M, int D, const double* X, int64_t N,
                                             const double* g_var, const double* h_mean, const double* log_noise,
                                             double* out, void* workspace, void* stream) {
  GVI_CHECK_ARG(M >= 1 && D >= 1 && N >= 1, "sizes");
  GVI_CHECK_ARG(factor && X && g_var && out && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (D > 32)
    return gvi_exact_gp_predict_grad_wide((const double*)factor, M, D, X, N, g_var, h_mean, log_noise, out, workspace,
                                          st);
  const FactorLayout f = factor_layout(M, D);
  const double* F = (const double*)factor;
  const double* Linv = F + f.linv_off;
  const double* alpha = F + f.alpha_off;
  const int Mp = f.Mp;
  ExactWs w = carve_exact(workspace, N, M, nullptr);
  int rc = GVI_OK;
  GVI_CUDA(cudaMemsetAsync(w.w, 0, (size_t)Mp * 8, st));
  for (int64_t r0 = 0; r0 < N; r0 += ECHUNK) {
    const int nr = (int)(N - r0 < ECHUNK ? N - r0 : ECHUNK);
    const int first = (r0 == 0);
    const int64_t elems = (int64_t)nr * Mp;
    const int blocks = (int)((elems + 255) / 256 < GVI_NUM_SMS * 8 ? (elems + 255) / 256 : GVI_NUM_SMS * 8);
    kxz_kernel<<<blocks, 256, 0, st>>>(F, f, X + r0 * D, nr, w.K);
    GVI_CHECK_LAUNCH("kxz_kernel");
    // Aa = K A^-1 = (K L^-T) L^-1
    // (in-tree tensor-pipe GEMM, dgemm.cu: L^-1 is lower triangular, so both products run over half the k range)
    if ((rc = gvi_dgemm_rm(false, true, nr, Mp, Mp, w.K, Mp, Linv, Mp, 0.0, w.T, Mp, GVI_GEMM_K_UPTO_N, false, st)))
      return rc;
    if ((rc = gvi_dgemm_rm(false, false, nr, Mp, Mp, w.T, Mp, Linv, Mp, 0.0, w.Aa, Mp, GVI_GEMM_K_FROM_N, false, st)))
      return rc;
    // B = diag(g) Aa -> T,  G1 = h alpha^T - 2 B -> K
    cotangent_kernel<<<blocks, 256, 0, st>>>(w.Aa, g_var + r0, h_mean ? h_mean + r0 : nullptr, alpha, nr, Mp, w.T, w.K);
    GVI_CHECK_LAUNCH("cotangent_kernel");
    // G2 (+)= Aa^T B ;  w += Aa^T h
    // (symmetric: B = diag(g) Aa - only the tiles on or below the diagonal are multiplied)
    if ((rc = gvi_dgemm_rm(true, false, Mp, Mp, nr, w.Aa, Mp, w.T, Mp, first ? 0.0 : 1.0, w.G2, Mp, GVI_GEMM_K_FULL, true,
                           st)))
      return rc;
    if (h_mean && (rc = gvi_dgemv_rm(true, Mp, nr, w.Aa, Mp, h_mean + r0, 1.0, w.w, st))) return rc;
    ard_contract_kernel<<<nr, 256, 0, st>>>(w.K, Mp, M, X + r0 * D, 1, F, f, g_var + r0, w.rowpart);
    GVI_CHECK_LAUNCH("ard_contract_kernel");
    contract_finish_kernel<<<33, 256, 0, st>>>(w.rowpart, nr, D, first, w.acc1);
    GVI_CHECK_LAUNCH("contract_finish_kernel");
  }
  rank1_sub_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(w.G2, w.w, alpha, Mp);
  GVI_CHECK_LAUNCH("rank1_sub_kernel");
  ard_contract_kernel<<<M, 256, 0, st>>>(w.G2, Mp, M, F + f.zs_off, 0, F, f, nullptr, w.rowpart);
  GVI_CHECK_LAUNCH("ard_contract_kernel");
  contract_finish_kernel<<<33, 256, 0, st>>>(w.rowpart, M, D, 1, w.acc2);
  GVI_CHECK_LAUNCH("contract_finish_kernel");
  exact_finish_kernel<<<1, 256, 0, st>>>(w.acc1, w.acc2, w.G2, M, Mp, D, N, log_noise, out);
  GVI_CHECK_LAUNCH("exact_finish_kernel");
  return GVI_OK;
}
