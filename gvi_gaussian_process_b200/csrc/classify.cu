// classify.cu - SURVEY.md section 8(f)-2: the classification epilogue on [K,N] predictive means / variances.
//   * class probabilities by Gauss-Hermite quadrature of prod_{l != j} Phi(.)
//     (GPClassificationBase._calculate_s_matrix / _predict_probability, src/gps/base/classification_base.py:53-153)
//     and the vector-Jacobian product of that map (what jax.grad does through it in the reference's trainer);
//   * the risks / regularisers that consume probabilities: multinomial NLL (negative_log_likelihood.py:57-78),
//     the discrete Wasserstein regulariser (multinomial_wasserstein_regularisation.py:46-76) and the cross entropy
//     (cross_entropy.py:11-30; jax_metrics' Crossentropy treats `preds` as logits by default) - value and gradient.
// The probability kernels are FP64-pipe bound (H * K * (K-1) normal-CDF evaluations per point, no reuse, O(K) doubles
// of HBM traffic per point); one warp owns a point, lanes own quadrature nodes, the per-class sums are reduced with
// shuffle trees in a fixed order (bit-reproducible).
#include "common.cuh"

namespace {
constexpr int CTHREADS = 128;           // 4 warps, one point per warp at a time
constexpr int CWARPS = CTHREADS / 32;
constexpr double INV_SQRT2 = 0.70710678118654752440;
constexpr double SQRT2 = 1.41421356237309504880;
constexpr double INV_SQRT_PI = 0.56418958354775628695;
constexpr double INV_SQRT_2PI = 0.39894228040143267794;

__device__ __forceinline__ double norm_cdf(double u) { return 0.5 * erfc(-u * INV_SQRT2); }

// s[n, j] = 1/sqrt(pi) sum_h w_h prod_{l != j} max(Phi((sqrt2 r_h sd_j + m_j - m_l) / sd_l), lb)
// prob    = (1 - eps) s + eps / (K - 1) (1 - s)
__global__ void __launch_bounds__(CTHREADS) class_prob_kernel(const double* __restrict__ mean,
                                                              const double* __restrict__ var, int K, int64_t N,
                                                              const double* __restrict__ roots,
                                                              const double* __restrict__ weights, int H, double eps,
                                                              double lb, double* __restrict__ prob) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* m = sm + (size_t)warp * 2 * K;
  double* sd = m + K;
  const int64_t wstride = (int64_t)gridDim.x * CWARPS;
  for (int64_t n = (int64_t)blockIdx.x * CWARPS + warp; n < N; n += wstride) {
    __syncwarp();
    for (int k = lane; k < K; k += 32) {
      m[k] = mean[(int64_t)k * N + n];
      sd[k] = sqrt(var[(int64_t)k * N + n]);
    }
    __syncwarp();
    for (int j = 0; j < K; j++) {
      double acc = 0.0;
      for (int h = lane; h < H; h += 32) {
        const double base = SQRT2 * roots[h] * sd[j] + m[j];
        double p = 1.0;
        for (int l = 0; l < K; l++) {
          if (l == j) continue;
          p *= fmax(norm_cdf((base - m[l]) / sd[l]), lb);
        }
        acc = fma(weights[h], p, acc);
      }
      const double s = warp_sum(acc) * INV_SQRT_PI;
      if (lane == 0) prob[n * K + j] = (1.0 - eps) * s + (eps / (K - 1)) * (1.0 - s);
    }
  }
}

// VJP of the map above: given dprob[N,K] returns dmean[K,N], dvar[K,N].
// With u_hjl = (sqrt2 r_h sd_j + m_j - m_l) / sd_l, P_hj = prod_l Phi_c(u_hjl), rho = phi / Phi_c (0 where clipped):
//   ds_j/dm_j  =  c sum_h w_h P_hj sum_l rho_hjl / sd_l          ds_j/dm_l  = -c sum_h w_h P_hj rho_hjl / sd_l
//   ds_j/dsd_j =  c sum_h w_h P_hj sqrt2 r_h sum_l rho_hjl/sd_l  ds_j/dsd_l = -c sum_h w_h P_hj rho_hjl u_hjl / sd_l
// Thread-private accumulators live in shared memory as [K][CTHREADS] (conflict-free, no dynamic register indexing).
__global__ void __launch_bounds__(CTHREADS) class_prob_grad_kernel(const double* __restrict__ mean,
                                                                   const double* __restrict__ var, int K, int64_t N,
                                                                   const double* __restrict__ roots,
                                                                   const double* __restrict__ weights, int H,
                                                                   double eps, double lb,
                                                                   const double* __restrict__ dprob,
                                                                   double* __restrict__ dmean,
                                                                   double* __restrict__ dvar) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* m = sm + (size_t)warp * 3 * K;
  double* sd = m + K;
  double* cj = sd + K;
  double* priv = sm + (size_t)CWARPS * 3 * K;  // 4 arrays [K][CTHREADS]
  double* gm = priv + tid;
  double* gs = gm + (size_t)K * CTHREADS;
  double* t1 = gs + (size_t)K * CTHREADS;
  double* t2 = t1 + (size_t)K * CTHREADS;
  const double dscale = (1.0 - eps - eps / (K - 1)) * INV_SQRT_PI;
  const int64_t wstride = (int64_t)gridDim.x * CWARPS;
  for (int64_t n = (int64_t)blockIdx.x * CWARPS + warp; n < N; n += wstride) {
    __syncwarp();
    for (int k = lane; k < K; k += 32) {
      m[k] = mean[(int64_t)k * N + n];
      sd[k] = sqrt(var[(int64_t)k * N + n]);
      cj[k] = dprob[n * K + k] * dscale;
    }
    for (int k = 0; k < K; k++) gm[k * CTHREADS] = gs[k * CTHREADS] = 0.0;
    __syncwarp();
    for (int h = lane; h < H; h += 32) {
      const double rh = SQRT2 * roots[h], wh = weights[h];
      for (int j = 0; j < K; j++) {
        const double coef = wh * cj[j];
        if (coef == 0.0) continue;
        const double base = rh * sd[j] + m[j];
        double p = 1.0, sum1 = 0.0;
        for (int l = 0; l < K; l++) {
          if (l == j) continue;
          const double u = (base - m[l]) / sd[l];
          const double cdf = norm_cdf(u);
          const bool clipped = cdf < lb;
          const double cc = clipped ? lb : cdf;
          const double rho = clipped ? 0.0 : INV_SQRT_2PI * exp(-0.5 * u * u) / (cc * sd[l]);
          p *= cc;
          sum1 += rho;
          t1[l * CTHREADS] = rho;
          t2[l * CTHREADS] = u;
        }
        const double cp = coef * p;
        gm[j * CTHREADS] = fma(cp, sum1, gm[j * CTHREADS]);
        gs[j * CTHREADS] = fma(cp * rh, sum1, gs[j * CTHREADS]);
        for (int l = 0; l < K; l++) {
          if (l == j) continue;
          const double r1 = cp * t1[l * CTHREADS];
          gm[l * CTHREADS] -= r1;
          gs[l * CTHREADS] = fma(-r1, t2[l * CTHREADS], gs[l * CTHREADS]);
        }
      }
    }
    for (int k = 0; k < K; k++) {
      const double a = warp_sum(gm[k * CTHREADS]);
      const double b = warp_sum(gs[k * CTHREADS]);
      if (lane == 0) {
        dmean[(int64_t)k * N + n] = a;
        dvar[(int64_t)k * N + n] = b / (2.0 * sd[k]);
      }
    }
  }
}

// ---- risks on probabilities -------------------------------------------------------------------------------------
constexpr int MTHREADS = 256;
constexpr int MBLOCKS = GVI_NUM_SMS * 4;

__device__ __forceinline__ double powi(double x, int p) {
  double r = 1.0;
  for (int i = 0; i < p; i++) r *= x;
  return r;
}

// per point: nll_i = -log sum_k q_ik y_ik;  w_i = (sum_k |p_ik - q_ik|^pw)^(1/pw);
//            ce_i = -sum_k y_ik log_softmax(q_i)_k   (probabilities used as logits, like the reference)
__global__ void __launch_bounds__(MTHREADS) multinomial_risk_kernel(const double* __restrict__ q,
                                                                    const double* __restrict__ y,
                                                                    const double* __restrict__ p, int K, int64_t N,
                                                                    int power, double* __restrict__ partials) {
  __shared__ double red[3][MTHREADS / 32];
  const int tid = threadIdx.x;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * MTHREADS + tid; i < N; i += (int64_t)gridDim.x * MTHREADS) {
    const double* qi = q + i * K;
    if (y) {
      const double* yi = y + i * K;
      double dot = 0.0, mx = qi[0], ysum = 0.0, yq = 0.0;
      for (int k = 0; k < K; k++) {
        dot = fma(qi[k], yi[k], dot);
        mx = fmax(mx, qi[k]);
        ysum += yi[k];
        yq = fma(yi[k], qi[k], yq);
      }
      double se = 0.0;
      for (int k = 0; k < K; k++) se += exp(qi[k] - mx);
      a -= log(dot);
      c += ysum * (mx + log(se)) - yq;
    }
    if (p) {
      const double* pi = p + i * K;
      double s = 0.0;
      for (int k = 0; k < K; k++) s += powi(fabs(pi[k] - qi[k]), power);
      b += (power == 2) ? sqrt(s) : (power == 1 ? s : pow(s, 1.0 / power));
    }
  }
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = a;
    red[1][tid >> 5] = b;
    red[2][tid >> 5] = c;
  }
  __syncthreads();
  if (tid < 3) {
    double t = 0.0;
    for (int w = 0; w < MTHREADS / 32; w++) t += red[tid][w];
    partials[3 * blockIdx.x + tid] = t;
  }
}

__global__ void __launch_bounds__(96) multinomial_finish_kernel(const double* __restrict__ partials, int nblocks,
                                                                int64_t N, double* __restrict__ out4) {
  // one warp per quantity, fixed order
  const int which = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double t = 0.0;
  for (int i = lane; i < nblocks; i += 32) t += partials[3 * i + which];
  t = warp_sum(t);
  if (lane == 0) {
    out4[which == 2 ? 3 : which] = t;
    if (which == 0) out4[2] = (double)N;
  }
}

// dq[i,k] = w4[0] d nll_i/dq_ik + w4[1] d w_i/dq_ik + w4[3] d ce_i/dq_ik  (w4 = the cotangent of out4, on the device)
__global__ void __launch_bounds__(MTHREADS) multinomial_risk_grad_kernel(const double* __restrict__ q,
                                                                         const double* __restrict__ y,
                                                                         const double* __restrict__ p, int K,
                                                                         int64_t N, int power,
                                                                         const double* __restrict__ w4,
                                                                         double* __restrict__ dq) {
  const double w_nll = w4[0], w_reg = w4[1], w_ce = w4[3];
  for (int64_t i = (int64_t)blockIdx.x * MTHREADS + threadIdx.x; i < N; i += (int64_t)gridDim.x * MTHREADS) {
    const double* qi = q + i * K;
    double* di = dq + i * K;
    double dot = 0.0, mx = qi[0], ysum = 0.0, se = 0.0, s = 0.0;
    if (y) {
      const double* yi = y + i * K;
      for (int k = 0; k < K; k++) {
        dot = fma(qi[k], yi[k], dot);
        mx = fmax(mx, qi[k]);
        ysum += yi[k];
      }
      for (int k = 0; k < K; k++) se += exp(qi[k] - mx);
    }
    if (p) {
      const double* pi = p + i * K;
      for (int k = 0; k < K; k++) s += powi(fabs(pi[k] - qi[k]), power);
    }
    const double outer = (p && s > 0.0) ? pow(s, 1.0 / power - 1.0) : 0.0;
    for (int k = 0; k < K; k++) {
      double g = 0.0;
      if (y) {
        const double yk = y[i * K + k];
        g += w_nll * (-yk / dot);
        g += w_ce * (ysum * exp(qi[k] - mx) / se - yk);
      }
      if (p) {
        const double diff = qi[k] - p[i * K + k];
        const double ad = fabs(diff);
        if (ad > 0.0) g += w_reg * outer * powi(ad, power - 1) * (diff > 0.0 ? 1.0 : -1.0);
      }
      di[k] = g;
    }
  }
}

int check_class_args(const double* mean, const double* var, int K, int64_t N, const double* roots,
                     const double* weights, int H) {
  GVI_CHECK_ARG(K >= 2 && K <= 64, "2 <= K <= 64 classes");
  GVI_CHECK_ARG(H >= 1, "H");
  GVI_CHECK_ARG(N >= 0, "N");
  if (N == 0) return GVI_OK;
  GVI_CHECK_ARG(mean && var && roots && weights, "null pointer");
  return GVI_OK;
}
}  // namespace

extern "C" int gvi_class_probabilities_f64(const double* mean, const double* var, int K, int64_t N,
                                           const double* roots, const double* weights, int H, double epsilon,
                                           double cdf_lower_bound, double* prob_out, void* stream) {
  int rc = check_class_args(mean, var, K, N, roots, weights, H);
  if (rc) return rc;
  if (N == 0) return GVI_OK;
  GVI_CHECK_ARG(prob_out, "null pointer");
  const int64_t want = (N + CWARPS - 1) / CWARPS;
  const int blocks = (int)(want < GVI_NUM_SMS * 16 ? want : GVI_NUM_SMS * 16);
  class_prob_kernel<<<blocks, CTHREADS, (size_t)CWARPS * 2 * K * sizeof(double), (cudaStream_t)stream>>>(
      mean, var, K, N, roots, weights, H, epsilon, cdf_lower_bound, prob_out);
  GVI_CHECK_LAUNCH("class_prob_kernel");
  return GVI_OK;
}

extern "C" int gvi_class_probabilities_grad_f64(const double* mean, const double* var, int K, int64_t N,
                                                const double* roots, const double* weights, int H, double epsilon,
                                                double cdf_lower_bound, const double* dprob, double* dmean_out,
                                                double* dvar_out, void* stream) {
  int rc = check_class_args(mean, var, K, N, roots, weights, H);
  if (rc) return rc;
  if (N == 0) return GVI_OK;
  GVI_CHECK_ARG(dprob && dmean_out && dvar_out, "null pointer");
  const size_t smem = ((size_t)CWARPS * 3 * K + (size_t)4 * K * CTHREADS) * sizeof(double);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    GVI_CUDA(cudaFuncSetAttribute(class_prob_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t want = (N + CWARPS - 1) / CWARPS;
  const int blocks = (int)(want < GVI_NUM_SMS * 8 ? want : GVI_NUM_SMS * 8);
  class_prob_grad_kernel<<<blocks, CTHREADS, smem, (cudaStream_t)stream>>>(
      mean, var, K, N, roots, weights, H, epsilon, cdf_lower_bound, dprob, dmean_out, dvar_out);
  GVI_CHECK_LAUNCH("class_prob_grad_kernel");
  return GVI_OK;
}

extern "C" size_t gvi_multinomial_risk_workspace_bytes(int64_t N) {
  (void)N;
  return (size_t)MBLOCKS * 3 * sizeof(double);
}

extern "C" int gvi_multinomial_risk_f64(const double* prob_q, const double* y, const double* prob_p, int K,
                                        int64_t N, int power, double* out4, void* workspace, void* stream) {
  GVI_CHECK_ARG(K >= 2 && N >= 1 && power >= 1, "sizes");
  GVI_CHECK_ARG(prob_q && out4 && workspace, "null pointer");
  const int64_t want = (N + MTHREADS - 1) / MTHREADS;
  const int blocks = (int)(want > MBLOCKS ? MBLOCKS : want);
  cudaStream_t st = (cudaStream_t)stream;
  multinomial_risk_kernel<<<blocks, MTHREADS, 0, st>>>(prob_q, y, prob_p, K, N, power, (double*)workspace);
  GVI_CHECK_LAUNCH("multinomial_risk_kernel");
  multinomial_finish_kernel<<<1, 96, 0, st>>>((const double*)workspace, blocks, N, out4);
  GVI_CHECK_LAUNCH("multinomial_finish_kernel");
  return GVI_OK;
}

extern "C" int gvi_multinomial_risk_grad_f64(const double* prob_q, const double* y, const double* prob_p, int K,
                                             int64_t N, int power, const double* weights4, double* dprob_out,
                                             void* stream) {
  GVI_CHECK_ARG(K >= 2 && N >= 1 && power >= 1, "sizes");
  GVI_CHECK_ARG(prob_q && dprob_out && weights4, "null pointer");
  const int64_t want = (N + MTHREADS - 1) / MTHREADS;
  const int blocks = (int)(want > MBLOCKS ? MBLOCKS : want);
  multinomial_risk_grad_kernel<<<blocks, MTHREADS, 0, (cudaStream_t)stream>>>(prob_q, y, prob_p, K, N, power,
                                                                             weights4, dprob_out);
  GVI_CHECK_LAUNCH("multinomial_risk_grad_kernel");
  return GVI_OK;
}
