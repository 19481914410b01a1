// gram_mma.cu - materialised ARD Gram for 5 <= D < 24 with the x~ . z~ part of the exponent on the FP64 tensor pipe,
// behind a guard.  (ARDKernel.calculate_gram: src/kernels/standard/base.py:95-103, ard_kernel.py:58-66.)
//
// The direct-difference tile kernel (gram.cu) spends 2 FP64 instructions per element and dimension (subtract, fma) plus
// 12 for the exp: 28 per element at D = 8, which bounds it at 52 % of the HBM store roofline.  Expanding
//   s_ij = |u_i - v_j|^2 = |u_i|^2 + |v_j|^2 - 2 u_i . v_j        (u, v: inputs centred on x2[0], scaled by sqrt(w_d / 2))
// moves the D-dependent part onto DMMA m8n8k4: ONE datapath slot per element and dimension (the accumulator starts at
// |u_i|^2 + |v_j|^2 - 2 log_scaling, the A operand carries -2 u), i.e. 8 + 1 + 12 = 21 slots per element at D = 8.
// Measured (400 000 x 1024, tools/gram_time.py, guard and fall-back launches included): D = 8: 0.773 ms = 4.24 TB/s =
// 64.6 % of the measured store roofline (direct kernel: 0.962 ms = 52 %); D = 16: 47 % (33 %); D = 20: 36 % (15 %).
// The expansion cancels: its absolute error is ~ (D + 3) eps (|u_i| + |v_j|)^2, which is the relative error of k.  The
// GUARD makes that a bound instead of a hope: gram_guard_kernel reduces R1 = max_i |u_i|, R2 = max_j |v_j| (one pass
// over the inputs, 8 D bytes per row against 8 M bytes of output per row) and the tensor-pipe kernel runs only if
// (R1 + R2)^2 <= 4096, i.e. error <= 27 x 1.1e-16 x 4096 = 1.2e-11 relative, inside the 1e-10 parity tolerance; otherwise
// it exits at once and the direct-difference kernel (launched right behind it, exiting at once in the other case) does
// the work - short lengthscales (the reference's identity-like golden: exp(log_ls) = e^20), far-apart clusters, inf / NaN
// inputs (a NaN norm fails the test) all take the exact route.  No host synchronisation: both kernels read the
// guard word themselves.
// Not clamped: a rounding-negative distance (>= -1.2e-11 under the guard) gives k <= amp^2 (1 + 1.2e-11).
#include "common.cuh"

namespace {
// slot = max(slot, max_rows sum_d (sc_d (x_d - c_d))^2) as the bit pattern of a non-negative double (monotonic as an
// unsigned integer; NaN patterns compare above every finite value, so a NaN row fails the guard)
__global__ void __launch_bounds__(256) gram_guard_kernel(const double* __restrict__ x, int64_t m, int D,
                                                         const double* __restrict__ centre,
                                                         const double* __restrict__ log_ls, int ls_len,
                                                         unsigned long long* __restrict__ slot) {
  __shared__ double sc2[32], ctr[32];
  __shared__ unsigned long long wmax[8];
  const int tid = threadIdx.x;
  if (tid < D) {
    sc2[tid] = 0.5 * exp(log_ls[ls_len == 1 ? 0 : tid]);
    ctr[tid] = centre[tid];
  }
  __syncthreads();
  unsigned long long best = 0ull;
  for (int64_t r = (int64_t)blockIdx.x * 256 + tid; r < m; r += (int64_t)gridDim.x * 256) {
    double s = 0.0;
    for (int d = 0; d < D; d++) {
      const double diff = x[r * D + d] - ctr[d];
      s = fma(sc2[d] * diff, diff, s);
    }
    const unsigned long long b = (unsigned long long)__double_as_longlong(s);
    best = b > best ? b : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((tid & 31) == 0) wmax[tid >> 5] = best;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; w++) best = wmax[w] > best ? wmax[w] : best;
    atomicMax(slot, best);
  }
}

// One warp: 8 RTW rows x 8 NTW columns per step; its columns' B fragments (v) and norms stay in registers over the CTA's
// row range.  8 warps per CTA = 64 NTW columns.  Fragment roles (mma.m8n8k4): lane = 4 r + q holds A[r][q], B[q][r] and
// C[r][2 q], C[r][2 q + 1].
template <int KP, int RTW, int NTW>
__global__ void __launch_bounds__(256, 2) ard_gram_mma_kernel(const double* __restrict__ x1, int64_t m1,
                                                             const double* __restrict__ x2, int64_t m2, int D,
                                                             const double* __restrict__ log_scaling,
                                                             const double* __restrict__ log_ls, int ls_len,
                                                             double* __restrict__ out, int64_t ld, int rows_per_cta,
                                                             const unsigned long long* __restrict__ guard) {
  constexpr int KS = 2 * KP;     // k-steps of 4 features
  constexpr int WC = 8 * NTW;    // columns per warp (NTW n-tiles of 8)
  __shared__ double2 tab[32][8];
  if (!gvi_gram_fast_ok(guard)) return;  // uniform: the direct-difference kernel behind this launch does the work
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane & 3, r = lane >> 2;
  tab[tid >> 3][tid & 7] = GVI_EXP2_TABLE[tid >> 3];
  __syncthreads();
  const double2* tb = &tab[0][tid & 7];
  double sc[KS], ct[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    const int k = q + 4 * ks;
    sc[ks] = k < D ? sqrt(0.5 * exp(log_ls[ls_len == 1 ? 0 : k])) : 0.0;
    ct[ks] = k < D ? x2[k] : 0.0;
  }
  double scm2[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ks++) scm2[ks] = -2.0 * sc[ks];
  const double acc0 = -2.0 * log_scaling[0];  // exp(-(acc0 + s)) = exp(log_scaling)^2 exp(-s)
  const int64_t n0 = (int64_t)blockIdx.x * (8 * WC) + warp * WC;
  if (n0 >= m2) return;
  double b[NTW][KS], nv[NTW][2];
#pragma unroll
  for (int j = 0; j < NTW; j++) {
    const int64_t col = n0 + 8 * j + r;
    double part = 0.0;
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const int k = q + 4 * ks;
      const double v = (col < m2 && k < D) ? sc[ks] * (x2[col * D + k] - ct[ks]) : 0.0;
      b[j][ks] = v;
      part = fma(v, v, part);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);  // |v|^2 of column n0 + 8 j + r in the four lanes of row r
#pragma unroll
    for (int e = 0; e < 2; e++) nv[j][e] = __shfl_sync(0xffffffffu, part, (2 * q + e) * 4) + acc0;
  }
  // the warp's columns all exist and every lane's column pair is 16-byte aligned: straight-line 16-byte stores
  const bool full_cols = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && n0 + WC <= m2;
  const int64_t row_lo = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t row_hi = row_lo + rows_per_cta < m1 ? row_lo + rows_per_cta : m1;
  // raw x1 values of the next step travel while this step's MMAs and exponentials run
  double xr[RTW][KS];
  auto load_rows = [&](int64_t rs) {
#pragma unroll
    for (int t = 0; t < RTW; t++) {
      const int64_t row = rs + 8 * t + r;
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        const int k = q + 4 * ks;
        xr[t][ks] = (row < row_hi && k < D) ? x1[row * D + k] : ct[ks];
      }
    }
  };
  load_rows(row_lo);
  for (int64_t rs = row_lo; rs < row_hi; rs += 8 * RTW) {
    double a[RTW][KS], acc[RTW][NTW][2];
#pragma unroll
    for (int t = 0; t < RTW; t++) {
      double part = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        a[t][ks] = scm2[ks] * (xr[t][ks] - ct[ks]);  // -2 u
        part = fma(a[t][ks], a[t][ks], part);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part *= 0.25;  // |u|^2 of row rs + 8 t + r
#pragma unroll
      for (int j = 0; j < NTW; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) acc[t][j][e] = part + nv[j][e];
    }
    if (rs + 8 * RTW < row_hi) load_rows(rs + 8 * RTW);
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
      for (int t = 0; t < RTW; t++)
#pragma unroll
        for (int j = 0; j < NTW; j++) dmma884(acc[t][j][0], acc[t][j][1], a[t][ks], b[j][ks]);
    // one range test for the warp-step's exponents (high words as integers; under the guard they are finite, but
    // s - 2 log_scaling may exceed 708 -> exactly 0 like gvi_exp, and a NaN log_scaling must come out as NaN): the
    // common case runs all RTW x 8 exponentials as one straight-line block of independent chains
    int hmax = __double2hiint(acc[0][0][0]);
#pragma unroll
    for (int t = 0; t < RTW; t++)
#pragma unroll
      for (int j = 0; j < NTW; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) hmax = max(hmax, __double2hiint(acc[t][j][e]));
    if (hmax < 0x40862000) {
#pragma unroll
      for (int t = 0; t < RTW; t++)
#pragma unroll
        for (int j = 0; j < NTW; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) acc[t][j][e] = gvi_exp_neg_t32<false>(acc[t][j][e], tb, 8);
    } else {
#pragma unroll
      for (int t = 0; t < RTW; t++)
#pragma unroll
        for (int j = 0; j < NTW; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) acc[t][j][e] = gvi_exp_neg_t32<true>(acc[t][j][e], tb, 8);
    }
    if (full_cols && rs + 8 * RTW <= row_hi) {
#pragma unroll
      for (int t = 0; t < RTW; t++) {
        double* dst = out + (rs + 8 * t + r) * ld + n0 + 2 * q;
#pragma unroll
        for (int j = 0; j < NTW; j++)
          __stcs(reinterpret_cast<double2*>(dst + 8 * j), make_double2(acc[t][j][0], acc[t][j][1]));
      }
    } else {
#pragma unroll
      for (int t = 0; t < RTW; t++) {
        const int64_t row = rs + 8 * t + r;
#pragma unroll
        for (int j = 0; j < NTW; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int64_t col = n0 + 8 * j + 2 * q + e;
            if (row < row_hi && col < m2) __stcs(out + row * ld + col, acc[t][j][e]);
          }
      }
    }
  }
}

template <int KP, int RTW, int NTW>
int launch_mma(const double* x1, int64_t m1, const double* x2, int64_t m2, int D, const double* log_scaling,
               const double* log_ls, int ls_len, double* out, int64_t ld, const unsigned long long* guard,
               cudaStream_t st) {
  const int64_t gx = (m2 + 64 * NTW - 1) / (64 * NTW);
  // about 6 CTAs per SM worth of work, at least 64 rows per CTA (the column fragments are set up once per CTA)
  int64_t rows = (m1 * gx + (int64_t)GVI_NUM_SMS * 6 - 1) / ((int64_t)GVI_NUM_SMS * 6);
  rows = (rows + 8 * RTW - 1) / (8 * RTW) * (8 * RTW);
  if (rows < 64) rows = 64;
  if (rows > 4096) rows = 4096;
  const int64_t gy = (m1 + rows - 1) / rows;
  const int64_t max_gy = 65535;
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    const int64_t ny = gy - y0 < max_gy ? gy - y0 : max_gy;
    ard_gram_mma_kernel<KP, RTW, NTW><<<dim3((unsigned)gx, (unsigned)ny), 256, 0, st>>>(
        x1 + y0 * rows * D, m1 - y0 * rows, x2, m2, D, log_scaling, log_ls, ls_len, out + y0 * rows * ld, ld, (int)rows,
        guard);
    GVI_CHECK_LAUNCH("ard_gram_mma_kernel");
  }
  return GVI_OK;
}
}  // namespace

size_t gvi_gram_mma_workspace_bytes() { return 256; }

// guard words (R1^2, R2^2 bit patterns) into `guard`, then the tensor-pipe kernel; the caller launches the
// direct-difference kernel with the same guard pointer right behind (it exits at once when the fast path ran)
int gvi_launch_ard_gram_mma(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                            const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                            unsigned long long* guard, cudaStream_t st) {
  GVI_CUDA(cudaMemsetAsync(guard, 0, 2 * sizeof(unsigned long long), st));
  const int b1 = (int)((m1 + 255) / 256 < GVI_NUM_SMS * 8 ? (m1 + 255) / 256 : GVI_NUM_SMS * 8);
  const int b2 = (int)((m2 + 255) / 256 < GVI_NUM_SMS * 8 ? (m2 + 255) / 256 : GVI_NUM_SMS * 8);
  gram_guard_kernel<<<b1, 256, 0, st>>>(x1, m1, D, x2, log_ls, ls_len, guard);
  GVI_CHECK_LAUNCH("gram_guard_kernel");
  gram_guard_kernel<<<b2, 256, 0, st>>>(x2, m2, D, x2, log_ls, ls_len, guard + 1);
  GVI_CHECK_LAUNCH("gram_guard_kernel");
  // RTW x NTW x 2 = 8 independent exp chains per thread at 128 registers (two CTAs = 16 warps per SM): measured best -
  // 16 chains at 128 registers serialise in ptxas' schedule, at 255 registers only 8 warps fit (tools/gram_time.py)
  if (D <= 8) return launch_mma<1, 1, 4>(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld, guard, st);
  if (D <= 16) return launch_mma<2, 1, 4>(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld, guard, st);
  return launch_mma<3, 1, 4>(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld, guard, st);
}

