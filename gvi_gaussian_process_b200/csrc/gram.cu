// gram.cu - materialised ARD Gram (a1-a3).  Store-bound: 8 bytes per element written once, coalesced.
// Reference: src/kernels/standard/ard_kernel.py:58-66, src/kernels/standard/base.py:95-103, src/kernels/base.py:104-147.
#include "common.cuh"

namespace {
constexpr int TR = 32;    // rows of x1 per CTA
constexpr int TC = 256;   // columns (rows of x2) per CTA = threads per CTA
constexpr int DC = 16;    // feature chunk staged in shared memory

__global__ void __launch_bounds__(TC) ard_gram_kernel(const double* __restrict__ x1, int64_t m1,
                                                      const double* __restrict__ x2, int64_t m2, int D,
                                                      const double* __restrict__ log_scaling,
                                                      const double* __restrict__ log_ls, int ls_len,
                                                      double* __restrict__ out, int64_t ld,
                                                      const unsigned long long* __restrict__ guard) {
  __shared__ double xs[TR][DC];
  __shared__ double zs[DC][TC + 1];
  __shared__ double w[DC];
  if (guard && gvi_gram_fast_ok(guard)) return;  // the tensor-pipe kernel (gram_mma.cu) produced this Gram
  const int tid = threadIdx.x;
  const int64_t col0 = (int64_t)blockIdx.x * TC;
  const int64_t row0 = (int64_t)blockIdx.y * TR;
  const double amp = exp(log_scaling[0]);
  const double amp2 = amp * amp;
  double acc[TR];
#pragma unroll
  for (int r = 0; r < TR; r++) acc[r] = 0.0;
  for (int d0 = 0; d0 < D; d0 += DC) {
    const int dc = min(DC, D - d0);
    __syncthreads();
    if (tid < dc) w[tid] = exp(log_ls[ls_len == 1 ? 0 : d0 + tid]);
    for (int idx = tid; idx < TR * dc; idx += TC) {
      int r = idx / dc, d = idx % dc;
      xs[r][d] = (row0 + r < m1) ? x1[(row0 + r) * D + d0 + d] : 0.0;
    }
    for (int idx = tid; idx < TC * dc; idx += TC) {
      int c = idx / dc, d = idx % dc;
      zs[d][c] = (col0 + c < m2) ? x2[(col0 + c) * D + d0 + d] : 0.0;
    }
    __syncthreads();
    for (int d = 0; d < dc; d++) {
      const double z = zs[d][tid];
      const double wd = w[d];
#pragma unroll
      for (int r = 0; r < TR; r++) {
        double diff = xs[r][d] - z;
        acc[r] = fma(wd, diff * diff, acc[r]);
      }
    }
  }
  if (col0 + tid < m2) {
#pragma unroll
    for (int r = 0; r < TR; r++) {
      if (row0 + r < m1) out[(row0 + r) * ld + col0 + tid] = amp2 * gvi_exp(-0.5 * acc[r]);
    }
  }
}

constexpr int GRAM_RB = 64;      // rows of x1 staged per block of the tile kernel
constexpr int GRAM_RBLOCKS = 4;  // row blocks per CTA (large Grams; small ones take fewer so that the grid covers the SMs)

// Materialised Gram for D <= 16 - the API call ARDKernel.calculate_gram (src/kernels/standard/base.py:95-103,
// ard_kernel.py:58-66).  The output is written once (8 B per element: the HBM store roofline), so the kernel is shaped
// around the FP64 instruction count per element, which is what actually bounds it (28 per element at D = 8):
//   - inputs are centred on x2[0] and pre-scaled by sqrt(exp(log_ls_d) / 2) when they are staged (x rows in shared
//     memory, the thread's own columns in registers), so a dimension costs one subtract and one fma per element and the
//     exponent 0.5 sum_d w_d (x_d - z_d)^2 needs no further scaling; still DIRECT differences (no ||x||^2 + ||z||^2 -
//     2 x.z cancellation), the centring keeps the pre-scaled operands as small as the data spread;
//   - the amplitude enters through the accumulator's start value -2 log_scaling (exp(-s) exp(2 ls));
//   - the table-driven exp (12 FP64 instructions instead of 20), its tables replicated per half-warp lane;
//   - RT x CT independent accumulator / exp chains per thread, x rows read as broadcast LDS.128 (two dimensions each);
//   - adjacent column pairs per thread -> 16-byte streaming stores, 512 contiguous bytes per warp and row;
//   - a CTA keeps its columns in registers over RBLOCKS row blocks (the column loads and table fill amortise).
template <int DP, int CT>
#ifndef GRAM_MIN_CTAS
#define GRAM_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(256, GRAM_MIN_CTAS) ard_gram_tile_kernel(const double* __restrict__ x1, int64_t m1,
                                                               const double* __restrict__ x2, int64_t m2, int D,
                                                               const double* __restrict__ log_scaling,
                                                               const double* __restrict__ log_ls, int ls_len,
                                                               double* __restrict__ out, int64_t ld,
                                                               const unsigned long long* __restrict__ guard,
                                                               int rblocks) {
  if (guard && gvi_gram_fast_ok(guard)) return;  // the tensor-pipe kernel (gram_mma.cu) produced this Gram
  constexpr int RB = GRAM_RB;  // rows of x1 per staged block
  constexpr int RT = DP <= 8 ? 8 : 4;  // rows per inner step (lab: tools/lab/gram_lab.cu - 8 x 2 chains per thread at 3 CTAs/SM was best)
  __shared__ __align__(16) double xs[RB][DP];
  __shared__ double2 tab[32][8];
  __shared__ double sc[DP], ctr[DP];
  const int tid = threadIdx.x;
  const int64_t col0 = (int64_t)blockIdx.x * (256 * CT) + (int64_t)tid * CT;
  tab[tid >> 3][tid & 7] = GVI_EXP2_TABLE[tid >> 3];
  if (tid < DP) {
    sc[tid] = tid < D ? sqrt(0.5 * exp(log_ls[ls_len == 1 ? 0 : tid])) : 0.0;
    ctr[tid] = tid < D ? x2[tid] : 0.0;
  }
  __syncthreads();
  double z[CT][DP];
#pragma unroll
  for (int c = 0; c < CT; c++)
#pragma unroll
    for (int d = 0; d < DP; d++) z[c][d] = (col0 + c < m2 && d < D) ? sc[d] * (x2[(col0 + c) * D + d] - ctr[d]) : 0.0;
  const double acc0 = -2.0 * log_scaling[0];  // exp(-(acc0 + s)) = exp(log_scaling)^2 exp(-s)
  const double2* tb = &tab[0][tid & 7];
  const bool vec2 = CT == 2 && ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (col0 + 1 < m2);
  for (int blk = 0; blk < rblocks; blk++) {
    const int64_t row0 = ((int64_t)blockIdx.y * rblocks + blk) * RB;
    if (row0 >= m1) break;
    __syncthreads();  // the previous block's readers are done with xs
    for (int idx = tid; idx < RB * DP; idx += 256) {
      const int r = idx / DP, d = idx % DP;
      xs[r][d] = (row0 + r < m1 && d < D) ? sc[d] * (x1[(row0 + r) * D + d] - ctr[d]) : 0.0;
    }
    __syncthreads();
    const int rows = (int)((m1 - row0) < RB ? (m1 - row0) : RB);
    for (int r0 = 0; r0 < rows; r0 += RT) {
      double acc[RT][CT];
#pragma unroll
      for (int rr = 0; rr < RT; rr++)
#pragma unroll
        for (int c = 0; c < CT; c++) acc[rr][c] = acc0;
#pragma unroll
      for (int d = 0; d < DP; d += 2) {
#pragma unroll
        for (int rr = 0; rr < RT; rr++) {
          const double2 xv = *reinterpret_cast<const double2*>(&xs[r0 + rr][d]);
#pragma unroll
          for (int c = 0; c < CT; c++) {
            const double d0 = xv.x - z[c][d];
            acc[rr][c] = fma(d0, d0, acc[rr][c]);
            const double d1 = xv.y - z[c][d + 1];
            acc[rr][c] = fma(d1, d1, acc[rr][c]);
          }
        }
      }
      // one range test for the thread's RT x CT exponents (high words as integers: s > 708, +inf and NaN read as large);
      // the rare out-of-range ones are parked at 0 and patched afterwards, so that ALL RT x CT exponentials run as one
      // straight-line block of independent chains without per-element range logic
      int hmax = __double2hiint(acc[0][0]);
#pragma unroll
      for (int rr = 0; rr < RT; rr++)
#pragma unroll
        for (int c = 0; c < CT; c++) hmax = max(hmax, __double2hiint(acc[rr][c]));
      unsigned zero_mask = 0u, nan_mask = 0u;
      if (hmax >= 0x40862000) {
#pragma unroll
        for (int rr = 0; rr < RT; rr++)
#pragma unroll
          for (int c = 0; c < CT; c++) {
            const double a = acc[rr][c];
            if (a != a) nan_mask |= 1u << (rr * CT + c);
            else if (a > 708.0) zero_mask |= 1u << (rr * CT + c);
            if (!(a <= 708.0)) acc[rr][c] = 0.0;
          }
      }
      double v[RT][CT];
#pragma unroll
      for (int rr = 0; rr < RT; rr++)
#pragma unroll
        for (int c = 0; c < CT; c++) v[rr][c] = gvi_exp_neg_t32<false>(acc[rr][c], tb, 8);
      if (zero_mask | nan_mask) {
#pragma unroll
        for (int rr = 0; rr < RT; rr++)
#pragma unroll
          for (int c = 0; c < CT; c++) {
            if (zero_mask >> (rr * CT + c) & 1u) v[rr][c] = 0.0;  // exp(-s) < 1e-307: as gvi_exp
            if (nan_mask >> (rr * CT + c) & 1u) v[rr][c] = __longlong_as_double(0x7ff8000000000000ll);
          }
      }
      double* dst = out + (row0 + r0) * ld + col0;
      if (vec2 && r0 + RT <= rows) {
#pragma unroll
        for (int rr = 0; rr < RT; rr++)
          __stcs(reinterpret_cast<double2*>(dst + rr * ld), make_double2(v[rr][0], v[rr][CT - 1]));
      } else {
#pragma unroll
        for (int rr = 0; rr < RT; rr++)
#pragma unroll
          for (int c = 0; c < CT; c++)
            if (r0 + rr < rows && col0 + c < m2) __stcs(dst + rr * ld + c, v[rr][c]);
      }
    }
  }
}

template <int DP, int CT>
int launch_ard_gram_tile(const double* x1, int64_t m1, const double* x2, int64_t m2, int D, const double* log_scaling,
                         const double* log_ls, int ls_len, double* out, int64_t ld, const unsigned long long* guard,
                         cudaStream_t st) {
  const int64_t gx = (m2 + 256 * CT - 1) / (256 * CT);
  // K_ZZ of a factorisation (1024 x 1024: 2 column blocks) ran on 8 CTAs with four row blocks each - 74 us of every
  // factor chain; one row block per CTA until the grid reaches two CTAs per SM
  int rblocks = GRAM_RBLOCKS;
  while (rblocks > 1 && gx * ((m1 + (int64_t)GRAM_RB * rblocks - 1) / ((int64_t)GRAM_RB * rblocks)) < 2 * GVI_NUM_SMS) rblocks /= 2;
  const int64_t rows_per_cta = (int64_t)GRAM_RB * rblocks;
  const int64_t gy = (m1 + rows_per_cta - 1) / rows_per_cta;
  const int64_t max_gy = 65535;  // grid.y limit: loop over row super-blocks
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    const int64_t ny = gy - y0 < max_gy ? gy - y0 : max_gy;
    ard_gram_tile_kernel<DP, CT><<<dim3((unsigned)gx, (unsigned)ny), 256, 0, st>>>(
        x1 + y0 * rows_per_cta * D, m1 - y0 * rows_per_cta, x2, m2, D, log_scaling, log_ls, ls_len,
        out + y0 * rows_per_cta * ld, ld, guard, rblocks);
    GVI_CHECK_LAUNCH("ard_gram_tile_kernel");
  }
  return GVI_OK;
}

__global__ void ard_gram_diag_kernel(int64_t n, const double* __restrict__ log_scaling, double* __restrict__ out) {
  const double amp = exp(log_scaling[0]);
  const double v = amp * amp;  // * exp(-0.5 * 0)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = v;
}

// ---- inner-product and polynomial kernels ------------------------------------------------------------------------
// k(x,z) = (exp(log_scaling) <x,z> + exp(log_constant))^degree   (src/kernels/non_stationary/polynomial_kernel.py:40-52;
// inner_product_kernel.py:36-42 is the case without constant and degree 1).  jnp.power with a Python-int exponent lowers
// to lax.integer_pow (repeated squaring), with a float exponent to pow: integral degrees take the multiplication chain.
__device__ __forceinline__ double gvi_poly_epilogue(double dot, double scale, double constant, double degree) {
  const double b = fma(scale, dot, constant);
  if (degree == 1.0) return b;
  const double r = rint(degree);
  if (r == degree && r >= 0.0 && r <= 64.0) {
    int n = (int)r;
    double acc = 1.0, p = b;
    while (n) {
      if (n & 1) acc *= p;
      p *= p;
      n >>= 1;
    }
    return acc;
  }
  return pow(b, degree);
}

// same tiling as ard_gram_kernel: 32 x 256 output tile per CTA, features staged 16 at a time, coalesced stores
__global__ void __launch_bounds__(TC) dot_gram_kernel(const double* __restrict__ x1, int64_t m1,
                                                      const double* __restrict__ x2, int64_t m2, int D,
                                                      const double* __restrict__ log_scaling,
                                                      const double* __restrict__ log_constant, double degree,
                                                      double* __restrict__ out, int64_t ld) {
  __shared__ double xs[TR][DC];
  __shared__ double zs[DC][TC + 1];
  const int tid = threadIdx.x;
  const int64_t col0 = (int64_t)blockIdx.x * TC;
  const int64_t row0 = (int64_t)blockIdx.y * TR;
  const double scale = exp(log_scaling[0]);
  const double constant = log_constant ? exp(log_constant[0]) : 0.0;
  double acc[TR];
#pragma unroll
  for (int r = 0; r < TR; r++) acc[r] = 0.0;
  for (int d0 = 0; d0 < D; d0 += DC) {
    const int dc = min(DC, D - d0);
    __syncthreads();
    for (int idx = tid; idx < TR * dc; idx += TC) {
      int r = idx / dc, d = idx % dc;
      xs[r][d] = (row0 + r < m1) ? x1[(row0 + r) * D + d0 + d] : 0.0;
    }
    for (int idx = tid; idx < TC * dc; idx += TC) {
      int c = idx / dc, d = idx % dc;
      zs[d][c] = (col0 + c < m2) ? x2[(col0 + c) * D + d0 + d] : 0.0;
    }
    __syncthreads();
    for (int d = 0; d < dc; d++) {
      const double z = zs[d][tid];
#pragma unroll
      for (int r = 0; r < TR; r++) acc[r] = fma(xs[r][d], z, acc[r]);
    }
  }
  if (col0 + tid < m2) {
#pragma unroll
    for (int r = 0; r < TR; r++) {
      if (row0 + r < m1) out[(row0 + r) * ld + col0 + tid] = gvi_poly_epilogue(acc[r], scale, constant, degree);
    }
  }
}

// diagonal mode of KernelBase.calculate_gram (src/kernels/base.py:132-147): out[i] = k(x1_i, x2_i), one thread per pair
__global__ void kernel_pairs_kernel(int kind, const double* __restrict__ x1, const double* __restrict__ x2, int64_t n,
                                    int D, const double* __restrict__ p0, const double* __restrict__ p1, int p1_len,
                                    double degree, double* __restrict__ out) {
  const double e0 = exp(p0[0]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double* a = x1 + i * D;
    const double* b = x2 + i * D;
    double s = 0.0;
    if (kind == GVI_KERNEL_ARD) {
      for (int d = 0; d < D; d++) {
        const double diff = a[d] - b[d];
        s = fma(exp(p1[p1_len == 1 ? 0 : d]), diff * diff, s);
      }
      out[i] = e0 * e0 * gvi_exp(-0.5 * s);
    } else {
      for (int d = 0; d < D; d++) s = fma(a[d], b[d], s);
      out[i] = gvi_poly_epilogue(s, e0, p1 ? exp(p1[0]) : 0.0, degree);
    }
  }
}
}  // namespace

extern "C" int gvi_dot_gram_f64(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                                const double* log_scaling, const double* log_constant, double degree, double* out,
                                int64_t ld_out, void* stream) {
  GVI_CHECK_ARG(m1 >= 0 && m2 >= 0 && D >= 1, "sizes");
  GVI_CHECK_ARG(ld_out >= m2, "ld_out < m2");
  if (m1 == 0 || m2 == 0) return GVI_OK;
  GVI_CHECK_ARG(x1 && x2 && log_scaling && out, "null pointer");
  const int64_t gy = (m1 + TR - 1) / TR, gx = (m2 + TC - 1) / TC, max_gy = 65535;
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    const int64_t ny = gy - y0 < max_gy ? gy - y0 : max_gy;
    dot_gram_kernel<<<dim3((unsigned)gx, (unsigned)ny), TC, 0, (cudaStream_t)stream>>>(
        x1 + y0 * TR * D, m1 - y0 * TR, x2, m2, D, log_scaling, log_constant, degree, out + y0 * TR * ld_out, ld_out);
    GVI_CHECK_LAUNCH("dot_gram_kernel");
  }
  return GVI_OK;
}

extern "C" int gvi_kernel_pairs_f64(int kind, const double* x1, const double* x2, int64_t n, int D, const double* p0,
                                    const double* p1, int p1_len, double degree, double* out, void* stream) {
  GVI_CHECK_ARG(kind == GVI_KERNEL_ARD || kind == GVI_KERNEL_POLYNOMIAL, "kind");
  GVI_CHECK_ARG(n >= 0 && D >= 1, "sizes");
  if (n == 0) return GVI_OK;
  GVI_CHECK_ARG(x1 && x2 && p0 && out, "null pointer");
  if (kind == GVI_KERNEL_ARD) GVI_CHECK_ARG(p1 && (p1_len == 1 || p1_len == D), "log_lengthscales must have 1 or D entries");
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  kernel_pairs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kind, x1, x2, n, D, p0, p1, p1_len, degree, out);
  GVI_CHECK_LAUNCH("kernel_pairs_kernel");
  return GVI_OK;
}

// guard != null: the tensor-pipe kernel of gram_mma.cu was launched ahead; these kernels exit at once if it ran
static int launch_ard_gram_direct(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                                  const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                                  const unsigned long long* guard, cudaStream_t st) {
  if (D <= 2) return launch_ard_gram_tile<2, 2>(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld, guard, st);
  if (D <= 4) return launch_ard_gram_tile<4, 2>(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld, guard, st);
  if (D <= 8) return launch_ard_gram_tile<8, 2>(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld, guard, st);
  if (D <= 16) return launch_ard_gram_tile<16, 1>(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld, guard, st);
  // 16 < D < 24 (or no workspace for the tensor-pipe kernel): one column per thread, features staged 16 at a time
  int64_t gy = (m1 + TR - 1) / TR;
  int64_t gx = (m2 + TC - 1) / TC;
  // grid.y is limited to 65535: loop over row super-blocks
  const int64_t max_gy = 65535;
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    int64_t ny = gy - y0 < max_gy ? gy - y0 : max_gy;
    dim3 grid((unsigned)gx, (unsigned)ny);
    ard_gram_kernel<<<grid, TC, 0, st>>>(x1 + y0 * TR * D, m1 - y0 * TR, x2, m2, D, log_scaling, log_ls, ls_len,
                                         out + y0 * TR * ld, ld, guard);
    GVI_CHECK_LAUNCH("ard_gram_kernel");
  }
  return GVI_OK;
}

int gvi_launch_ard_gram(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                        const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                        cudaStream_t st) {
  if (m1 == 0 || m2 == 0) return GVI_OK;
  return launch_ard_gram_direct(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld, nullptr, st);
}

size_t gvi_gram_mma_workspace_bytes();
int gvi_launch_ard_gram_mma(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                            const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                            unsigned long long* guard, cudaStream_t st);
// 5 <= D < 24 and an output large enough to pay for three extra launches: guarded tensor-pipe kernel, then the
// direct-difference kernel as its fall-back (exactly one of the two does the work; decided on the device)
constexpr int GVI_GRAM_MMA_MIN_D = 5;
constexpr int64_t GVI_GRAM_MMA_MIN_ELEMS = 1 << 18;

size_t gvi_gram_dmma_workspace_doubles(int64_t m1, int64_t m2, int D);
int gvi_launch_ard_gram_dmma(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                             const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                             double* workspace, cudaStream_t st);

// wide inputs go through the tensor-pipe contraction (gram_dmma.cu) when the caller supplies the workspace
#ifndef GVI_GRAM_DMMA_MIN_D_VALUE
#define GVI_GRAM_DMMA_MIN_D_VALUE 24  // (A/B switch: at D = 8 / 16 / 20 that kernel runs at 27 / 25 / 22 % of the store roofline)
#endif
constexpr int GVI_GRAM_DMMA_MIN_D = GVI_GRAM_DMMA_MIN_D_VALUE;

extern "C" size_t gvi_ard_gram_workspace_bytes(int64_t m1, int64_t m2, int D) {
  if (m1 < 1 || m2 < 1) return 0;
  if (D >= GVI_GRAM_DMMA_MIN_D) return gvi_gram_dmma_workspace_doubles(m1, m2, D) * sizeof(double);
  if (D >= GVI_GRAM_MMA_MIN_D && m1 * m2 >= GVI_GRAM_MMA_MIN_ELEMS) return gvi_gram_mma_workspace_bytes();
  return 0;
}

extern "C" int gvi_ard_gram_f64(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                                const double* log_scaling, const double* log_ls, int ls_len, double* out,
                                int64_t ld_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(m1 >= 0 && m2 >= 0 && D >= 1, "sizes");
  GVI_CHECK_ARG(ls_len == 1 || ls_len == D, "log_lengthscales must have 1 or D entries");
  GVI_CHECK_ARG(ld_out >= m2, "ld_out < m2");
  if (m1 == 0 || m2 == 0) return GVI_OK;
  GVI_CHECK_ARG(x1 && x2 && log_scaling && log_ls && out, "null pointer");
  if (D >= GVI_GRAM_DMMA_MIN_D && workspace)
    return gvi_launch_ard_gram_dmma(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld_out,
                                    (double*)workspace, (cudaStream_t)stream);
  if (workspace && D >= GVI_GRAM_MMA_MIN_D && D < GVI_GRAM_DMMA_MIN_D && m1 * m2 >= GVI_GRAM_MMA_MIN_ELEMS &&
      (reinterpret_cast<uintptr_t>(workspace) & 7) == 0) {
    unsigned long long* guard = (unsigned long long*)workspace;
    int rc = gvi_launch_ard_gram_mma(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld_out, guard,
                                     (cudaStream_t)stream);
    if (rc) return rc;
    return launch_ard_gram_direct(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld_out, guard,
                                  (cudaStream_t)stream);
  }
  return gvi_launch_ard_gram(x1, m1, x2, m2, D, log_scaling, log_ls, ls_len, out, ld_out, (cudaStream_t)stream);
}

extern "C" int gvi_ard_gram_diag_f64(int64_t n, const double* log_scaling, double* out, void* stream) {
  GVI_CHECK_ARG(n >= 0, "n");
  if (n == 0) return GVI_OK;
  GVI_CHECK_ARG(log_scaling && out, "null pointer");
  int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  ard_gram_diag_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, log_scaling, out);
  GVI_CHECK_LAUNCH("ard_gram_diag_kernel");
  return GVI_OK;
}
