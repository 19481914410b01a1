// common.cuh - shared helpers for libgvigp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../include/gvigp.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgvigp is written for sm_100a (B200) only"
#endif

void gvi_set_error(const char* fmt, ...);
void gvi_count_launch();

#define GVI_CHECK_ARG(cond, msg)                                  \
  do {                                                            \
    if (!(cond)) {                                                \
      gvi_set_error("invalid argument: %s (%s)", msg, #cond);     \
      return GVI_EINVAL;                                          \
    }                                                             \
  } while (0)

#define GVI_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    gvi_count_launch();                                                               \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      gvi_set_error("CUDA error launching %s: %s", name, cudaGetErrorString(e__));    \
      return GVI_ECUDA;                                                               \
    }                                                                                 \
  } while (0)

#define GVI_CUDA(call)                                                                \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      gvi_set_error("CUDA error in %s: %s", #call, cudaGetErrorString(e__));          \
      return GVI_ECUDA;                                                               \
    }                                                                                 \
  } while (0)

constexpr int GVI_NUM_SMS = 148;  // B200: 2 dies x 74 SMs

// ---- factor buffer layout (doubles) ---------------------------------------------------------------------------
// [0] amp2 = exp(log_scaling)^2   [1] k(x,x)   [2] jitter actually added   [3] chol info (as double)  [4..7] rsvd
// [8, 8+D)            sqrtw[d] = sqrt(exp(log_ls[d]))
// [zs_off, +Mp*D)     Z pre-scaled by sqrtw (rows >= M are zero)
// [alpha_off, +Mp)    alpha = A^-1 y_Z (zero padded)
// [wpack_off, ...)    L^-1 packed per 32-row group in DMMA m8n8k4 fragment order (see pack_linv_kernel)
struct FactorLayout {
  int M, Mp, D, G;
  int64_t sqrtw_off, zs_off, alpha_off, wpack_off, total;
};
__host__ __device__ inline FactorLayout factor_layout(int M, int D) {
  FactorLayout f;
  f.M = M;
  f.Mp = (M + 31) / 32 * 32;
  f.D = D;
  f.G = f.Mp / 32;
  f.sqrtw_off = 8;
  int64_t o = 8 + D;
  o = (o + 1) / 2 * 2;
  f.zs_off = o;
  o += (int64_t)f.Mp * D;
  o = (o + 1) / 2 * 2;
  f.alpha_off = o;
  o += f.Mp;
  o = (o + 1) / 2 * 2;
  f.wpack_off = o;
  o += 512LL * f.G * (f.G + 1);
  f.total = o;
  return f;
}
// offset (doubles) of 32-row group g inside wpack: sum_{g'<g} 1024*(g'+1)
__host__ __device__ inline int64_t wpack_group_off(int g) { return 512LL * g * (g + 1); }

// ---- FP64 tensor-core MMA (legacy warp-level path: SASS DMMA.8x8x4; there is no tcgen05 FP64 kind) -------------
// A 8x4 row-major: lane holds A[lane/4][lane%4]; B 4x8 col: lane holds B[lane%4][lane/4];
// C 8x8: lane holds C[lane/4][(lane%4)*2 + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// pointwise loss terms, shared by risk.cu and predict.cu ---------------------------------------------------------
__device__ __forceinline__ double gvi_nll_term(double y, double m, double v) {
  // -norm.logpdf(y, loc=m, scale=sqrt(v))  (negative_log_likelihood.py:44-48)
  const double half_log_2pi = 0.91893853320467274178;
  double sd = sqrt(v);
  double z = (y - m) / sd;
  return 0.5 * z * z + log(sd) + half_log_2pi;
}
__device__ __forceinline__ double gvi_reg_term(int kind, double alpha, double mp, double cp, double mq, double cq) {
  double dm = mp - mq;
  switch (kind) {
    case GVI_REG_PROJ_BHATTACHARYYA:
      return 0.125 * dm * dm / (cp + cq) + 0.5 * log(((cp + cq) / 2.0) / sqrt(cp * cq));
    case GVI_REG_PROJ_GAUSSIAN_WASSERSTEIN:
      return dm * dm + cp + cq - 2.0 * sqrt(cp * cq);
    case GVI_REG_PROJ_HELLINGER:
      return 1.0 - sqrt((2.0 * sqrt(cp) * sqrt(cq)) / (cp + cq)) * exp(-0.25 * dm * dm / (cp + cq));
    case GVI_REG_PROJ_KL:
      return log(sqrt(cq) / sqrt(cp)) + (cp + dm * dm) / (2.0 * cq) - 0.5;
    case GVI_REG_PROJ_RENYI: {
      double mix = alpha * cp + (1.0 - alpha) * cq;
      return log(sqrt(cp) / sqrt(cq)) + log(cp / mix) / (2.0 * (alpha - 1.0)) + 0.5 * alpha * dm * dm / mix;
    }
    case GVI_REG_GWI_NO_EIG:
      return dm * dm + cp + cq;
    case GVI_REG_SQUARED_DIFFERENCE: {
      double dc = cp - cq;
      return dm * dm + dc * dc;
    }
    default:
      return 0.0;
  }
}
