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

// cudaFuncSetAttribute is per device: `mask` (one static per kernel instantiation) remembers which devices are done, so
// a host that drives several GPUs from one process configures each of them.  True on the first call per device.
#include <atomic>
inline bool gvi_first_use_on_this_device(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  return !(mask.fetch_or(bit, std::memory_order_relaxed) & bit);
}

// profiling switches (GVI_DEBUG_SKIP, GVI_DEBUG_GRID, GVI_ONE_CTA, GVI_STATIC_TILES) exist only in libgvigp_test.so, the
// build with -DGVI_TEST_HOOKS that tests and tools load; the product library never reads the environment
#ifdef GVI_TEST_HOOKS
#include <cstdlib>
inline int gvi_debug_env(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}
#else
constexpr int gvi_debug_env(const char*) { return 0; }
#endif

// ---- factor buffer layout (doubles) ---------------------------------------------------------------------------
// [0] amp2 = exp(log_scaling)^2   [1] k(x,x)   [2] jitter actually added   [3] chol info (as double)
// [4] absolute jitter (0 when the regulariser is trace-relative; used by d var / d log_scaling)
// [5] 1 when the Cholesky factor was built from OTHER (frozen) kernel parameters than the K-tile parameters
// [6] 1 when the extra packed matrix E is set: var += ||E k||^2 (SVGP family)   [7] rsvd
// [8, 8+D)            sqrtw[d] = sqrt(exp(log_ls[d]))
// [zs_off, +Mp*D)     Z pre-scaled by sqrtw (rows >= M are zero)
// [alpha_off, +Mp)    alpha = A^-1 y_Z (zero padded)
// [wpack_off, ...)    L^-1 packed per 32-row group in DMMA m8n8k4 fragment order (see pack_linv_kernel)
// [wpackT_off, ...)   L^-T packed the same way (rows j, k-range c >= 32g): a = L^-T b in the backward pass
// [linv_off, +Mp*Mp)  dense row-major L^-1 (S = A^-1 C A^-1 in the backward pass; full-covariance API)
// [wextra_off, ...)   optional extra lower-triangular matrix E packed like L^-1 (gvi_gp_factor_set_extra_f64)
// rows of L^-1 handled by one warp-level group in the tile kernel = 8 * GVI_MT (m-tiles of the DMMA m8n8k4 shape)
constexpr int GVI_MT = 2;
constexpr int GVI_GROUP_ROWS = 8 * GVI_MT;
struct FactorLayout {
  int M, Mp, D, G;  // G = number of GVI_GROUP_ROWS-row groups
  int64_t sqrtw_off, zs_off, alpha_off, wpack_off, wpackT_off, linv_off, wextra_off, total;
};
__host__ __device__ inline int64_t wpack_total(int G) {
  return (int64_t)GVI_MT * 64 * (GVI_GROUP_ROWS / 8) * ((int64_t)G * (G + 1) / 2);
}
__host__ __device__ inline FactorLayout factor_layout(int M, int D) {
  FactorLayout f;
  f.M = M;
  f.Mp = (M + 31) / 32 * 32;
  f.D = D;
  f.G = f.Mp / GVI_GROUP_ROWS;
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
  o += wpack_total(f.G);
  f.wpackT_off = o;
  o += wpack_total(f.G);
  f.linv_off = o;
  o += (int64_t)f.Mp * f.Mp;
  f.wextra_off = o;
  o += wpack_total(f.G);
  f.total = o;
  return f;
}
// transposed pack: group g needs k-pairs kp0(g) = R g / 8 .. Mp/8 - 1, i.e. (R/8) (G - g) of them
__host__ __device__ inline int64_t wpackT_group_off(int g, int G) {
  return (int64_t)GVI_MT * 64 * (GVI_GROUP_ROWS / 8) * ((int64_t)g * G - (int64_t)g * (g - 1) / 2);
}
// group g (rows [R g, R g + R), R = GVI_GROUP_ROWS) needs k-pairs (8 columns each) 0 .. nkp(g)-1, nkp = R (g+1) / 8;
// one k-pair of one group = GVI_MT * 64 doubles.
__host__ __device__ inline int wpack_nkp(int g) { return GVI_GROUP_ROWS * (g + 1) / 8; }
__host__ __device__ inline int64_t wpack_group_off(int g) {
  return (int64_t)GVI_MT * 64 * (GVI_GROUP_ROWS / 8) * ((int64_t)g * (g + 1) / 2);
}

// ---- FP64 tensor-core MMA (legacy warp-level path: SASS DMMA.8x8x4; there is no tcgen05 FP64 kind) -------------
// A 8x4 row-major: lane holds A[lane/4][lane%4]; B 4x8 col: lane holds B[lane%4][lane/4];
// C 8x8: lane holds C[lane/4][(lane%4)*2 + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  // not volatile: the data dependences through the accumulators order the MMAs; everything else may be scheduled
  // between them (FP64-pipe work fills the issue slots the DMMA cadence leaves free)
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Branch-free exp for the kernel exponents (x = -0.5 * squared distance <= 0; valid for any x < 709): Cody-Waite
// reduction x = n ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (remainder < 5e-18 relative), scaling by an
// integer add into the exponent field; exactly 0 below -708 (k < 1e-307).  <= 1 ulp like CUDA's exp(), but without
// the slow-path branches, so independent evaluations interleave on the FP64 pipe and between DMMA issue slots.
__device__ __forceinline__ double gvi_exp(double x) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: the low word of (y + SHIFT) is rint(y)
  const bool tiny = x < -708.0;
  const double xc = tiny ? -708.0 : x;  // not fmax(): a NaN argument has to come out as NaN, as jnp.exp gives it
  const double t = fma(xc, 1.4426950408889634074, SHIFT);
  const int n = __double2loint(t);
  const double nf = t - SHIFT;
  double r = fma(nf, -6.93147180369123816490e-01, xc);
  r = fma(nf, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;             // 1/13!
  p = fma(p, r, 2.08767569878681e-09);           // 1/12!
  p = fma(p, r, 2.505210838544172e-08);          // 1/11!
  p = fma(p, r, 2.755731922398589e-07);          // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);         // 1/9!
  p = fma(p, r, 2.48015873015873e-05);           // 1/8!
  p = fma(p, r, 1.984126984126984e-04);          // 1/7!
  p = fma(p, r, 1.3888888888888889e-03);         // 1/6!
  p = fma(p, r, 8.333333333333333e-03);          // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);         // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);         // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  return tiny ? 0.0 : res;
}

// 2^(j/32), j = 0 .. 31, as correctly rounded double + the remainder (hi, lo): gvi_exp_neg_t32
static __device__ const double2 GVI_EXP2_TABLE[32] = {
    {1.0, 0.0},
    {1.0218971486541166, 5.109225028973444e-17},
    {1.0442737824274138, 8.551889705537965e-17},
    {1.0671404006768237, -7.899853966841582e-17},
    {1.0905077326652577, -3.046782079812471e-17},
    {1.1143867425958924, 1.0410278456845571e-16},
    {1.1387886347566916, 8.912812676025408e-17},
    {1.1637248587775775, 3.8292048369240935e-17},
    {1.189207115002721, 3.982015231465646e-17},
    {1.215247359980469, -7.712630692681488e-17},
    {1.241857812073484, 4.658027591836937e-17},
    {1.2690509571917332, 2.667932131342186e-18},
    {1.2968395546510096, 2.5382502794888315e-17},
    {1.3252366431597413, -2.8587312100388614e-17},
    {1.3542555469368927, 7.70094837980299e-17},
    {1.383909881963832, -6.770511658794786e-17},
    {1.4142135623730951, -9.667293313452913e-17},
    {1.4451808069770467, -3.0237581349939873e-17},
    {1.4768261459394993, -3.483994556892796e-17},
    {1.5091644275934228, -1.016455327754295e-16},
    {1.5422108254079407, 7.949834809697621e-17},
    {1.5759808451078865, -1.0136916471278304e-17},
    {1.6104903319492543, 2.4707192569797888e-17},
    {1.645755478153965, -1.0125679913674773e-16},
    {1.681792830507429, 8.199010020581497e-17},
    {1.718619298122478, -1.851380418263111e-17},
    {1.7562521603732995, 2.960140695448873e-17},
    {1.7947090750031072, 1.8227458427912087e-17},
    {1.8340080864093424, 3.283107224245627e-17},
    {1.8741676341103, -6.122763413004143e-17},
    {1.9152065613971474, -1.0619946056195963e-16},
    {1.9571441241754002, 8.960767791036668e-17}};

// Table-driven variant for kernels that are bound by the FP64 pipe itself (the materialised Gram: one exp per 8 bytes
// stored).  Takes s >= 0 and returns exp(-s): -s = (32 k + j) ln2/32 + r, |r| <= ln2/64,
// exp(-s) = 2^k (T_j + T_j (e^r - 1)) with T_j = 2^(j/32) = hi_j + lo_j from 32-entry tables and
// e^r - 1 = r (1 + r/2 + ... + r^5/720) (remainder r^7/5040 < 4e-18): 12 FP64 instructions instead of 20; only the final
// add rounds (0.56 ulp over the host sweep in tests/test_host_logic.py).  Exactly 0 for s > 708 (incl. +inf; the range test
// is two integer compares on the bit pattern, not an FP64 instruction), NaN -> NaN.  tab: entry j = (hi, lo) at
// [j * stride] (shared-memory copies of GVI_EXP2_TABLE; the lookups are data dependent, so the Gram kernel keeps 8
// interleaved copies - one per lane of a quarter-warp, the unit a 16-byte shared load is served in - and every lookup
// is one bank-conflict-free LDS.128).
// CHECKED = false skips the range test: the caller has established s <= 708 (or NaN) for a whole group of arguments.
template <bool CHECKED = true>
__device__ __forceinline__ double gvi_exp_neg_t32(double s, const double2* tab, int stride) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: the low word of (y + SHIFT) is rint(y)
  const long long bs = __double_as_longlong(s);
  const bool tiny = CHECKED && bs > 0x4086200000000000ll && bs <= 0x7ff0000000000000ll;  // s > 708 (or +inf), not NaN
  const double sc = tiny ? 708.0 : s;
  const double t = fma(sc, -46.16624130844683, SHIFT);  // -32 / ln2
  const int n = __double2loint(t);
  const double nf = t - SHIFT;
  double r = fma(nf, -0.021660849392446835, -sc);  // ln2/32, leading 36 bits: nf * hi is exact
  r = fma(nf, -5.145609244655338e-14, r);
  double p = 1.3888888888888889e-03;             // 1/6!
  p = fma(p, r, 8.333333333333333e-03);          // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);         // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);         // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  const double q = p * r;                        // e^r - 1
  const double2 tj = tab[(n & 31) * stride];
  const double res = fma(tj.x, q, tj.y) + tj.x;
  const double scaled = __hiloint2double(__double2hiint(res) + ((n >> 5) << 20), __double2loint(res));
  return tiny ? 0.0 : scaled;
}

// guard of the tensor-pipe Gram for small D (gram_mma.cu): guard[0], guard[1] = bit patterns of R1^2 = max_i |u_i|^2 and
// R2^2 = max_j |v_j|^2 (scaled, centred inputs).  True = the expanded form |u|^2 + |v|^2 - 2 u.v is accurate to
// ~1e-11 relative in k for every pair; false (also for NaN / inf norms) = direct differences.
__device__ __forceinline__ bool gvi_gram_fast_ok(const unsigned long long* __restrict__ guard) {
  const double r1 = __longlong_as_double((long long)guard[0]), r2 = __longlong_as_double((long long)guard[1]);
  return r1 + r2 + 2.0 * sqrt(r1 * r2) <= 4096.0;
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
// d nll / d m and d nll / d v
__device__ __forceinline__ void gvi_nll_grad(double y, double m, double v, double& dm, double& dv) {
  double r = y - m;
  dm = -r / v;
  dv = 0.5 / v - 0.5 * r * r / (v * v);
}
// d D0 / d m_q and d D0 / d c_q (p is the fixed regulariser GP)
__device__ __forceinline__ void gvi_reg_grad(int kind, double alpha, double mp, double cp, double mq, double cq,
                                             double& dmq, double& dcq) {
  const double dm = mp - mq;
  const double s = cp + cq;
  switch (kind) {
    case GVI_REG_PROJ_BHATTACHARYYA:
      dmq = -0.25 * dm / s;
      dcq = -0.125 * dm * dm / (s * s) + 0.5 / s - 0.25 / cq;
      break;
    case GVI_REG_PROJ_GAUSSIAN_WASSERSTEIN:
      dmq = -2.0 * dm;
      dcq = 1.0 - sqrt(cp / cq);
      break;
    case GVI_REG_PROJ_HELLINGER: {
      double be = sqrt((2.0 * sqrt(cp) * sqrt(cq)) / s) * exp(-0.25 * dm * dm / s);
      dmq = -be * (0.5 * dm / s);
      dcq = -be * (0.25 / cq - 0.5 / s + 0.25 * dm * dm / (s * s));
      break;
    }
    case GVI_REG_PROJ_KL:
      dmq = -dm / cq;
      dcq = 0.5 / cq - (cp + dm * dm) / (2.0 * cq * cq);
      break;
    case GVI_REG_PROJ_RENYI: {
      double mix = alpha * cp + (1.0 - alpha) * cq;
      dmq = -alpha * dm / mix;
      dcq = -0.5 / cq + 0.5 / mix - 0.5 * alpha * (1.0 - alpha) * dm * dm / (mix * mix);
      break;
    }
    case GVI_REG_GWI_NO_EIG:
      dmq = -2.0 * dm;
      dcq = 1.0;
      break;
    case GVI_REG_SQUARED_DIFFERENCE:
      dmq = -2.0 * dm;
      dcq = -2.0 * (cp - cq);
      break;
    default:
      dmq = 0.0;
      dcq = 0.0;
  }
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
