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
                                                      double* __restrict__ out, int64_t ld) {
  __shared__ double xs[TR][DC];
  __shared__ double zs[DC][TC + 1];
  __shared__ double w[DC];
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

int gvi_launch_ard_gram(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                        const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                        cudaStream_t st) {
  if (m1 == 0 || m2 == 0) return GVI_OK;
  int64_t gy = (m1 + TR - 1) / TR;
  int64_t gx = (m2 + TC - 1) / TC;
  // grid.y is limited to 65535: loop over row super-blocks
  const int64_t max_gy = 65535;
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    int64_t ny = gy - y0 < max_gy ? gy - y0 : max_gy;
    dim3 grid((unsigned)gx, (unsigned)ny);
    ard_gram_kernel<<<grid, TC, 0, st>>>(x1 + y0 * TR * D, m1 - y0 * TR, x2, m2, D, log_scaling, log_ls, ls_len,
                                         out + y0 * TR * ld, ld);
    GVI_CHECK_LAUNCH("ard_gram_kernel");
  }
  return GVI_OK;
}

size_t gvi_gram_dmma_workspace_doubles(int64_t m1, int64_t m2, int D);
int gvi_launch_ard_gram_dmma(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                             const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                             double* workspace, cudaStream_t st);

// wide inputs go through the tensor-pipe contraction (gram_dmma.cu) when the caller supplies the workspace
constexpr int GVI_GRAM_DMMA_MIN_D = 24;

extern "C" size_t gvi_ard_gram_workspace_bytes(int64_t m1, int64_t m2, int D) {
  if (m1 < 1 || m2 < 1 || D < GVI_GRAM_DMMA_MIN_D) return 0;
  return gvi_gram_dmma_workspace_doubles(m1, m2, D) * sizeof(double);
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
