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
// [6] 1 when the extra packed matrix E is set: var += ||E k||^2 (SVGP family)
// [7] max_j |z~_j - z~_0|^2 (guard of the tile kernel's tensor-pipe phase 1)
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

// ---- natural logarithm of the pointwise loss terms -----------------------------------------------------------------
// The risk reduction (risk.cu) is bound by its FP64 instructions, not by its 40 B / point of loads; a library log is
// ~29 FP64 instructions (division included).  Table-driven form: x = 2^k z with the bit pattern of z in
// [OFF, OFF + 2^52) (z in [0.69, 1.38)), interval i = top 7 bits, r = z invc_i - 1 (one fma: exact product, one rounding
// of a number <= 2^-8), log x = k ln2 + log c_i + log1p(r) with log1p(r) - r = r^2 (-1/2 + r/3 - ... + r^5/7) (remainder
// r^8/8 < 7e-21).  k ln2_hi + logc_hi is exact (both multiples of 2^-42, |sum| < 2^10): 14 FP64 instructions, <= 2 ulp
// over the host sweep of tests/test_host_logic.py (1.84 in the intervals next to the one around 1, where log c and r
// cancel by one bit; the loss terms need 1e-13); the interval around 1 has c = 1 exactly, so log(1) = 0 and the
// relative accuracy holds through x -> 1.  x <= 0, subnormal, inf and NaN take the library call (NaN / -inf as the
// reference's jnp.log).  Table: tools/gen_log_table.py.
// generated by tools/gen_log_table.py: {invc, logc_hi, logc_lo} per interval; max |r| = 0.003906
static __device__ const double GVI_LOG_TABLE[128][3] = {
    {0x1.724287f46debcp+0, -0x1.79e26687d0000p-2, 0x1.30a0168817444p-44},
    {0x1.702e05c0b8170p+0, -0x1.741d876c68000p-2, 0x1.13c7b5b11cfa7p-44},
    {0x1.6e1f76b4337c7p+0, -0x1.6e60ee6af2000p-2, 0x1.a3556a0f7749ep-44},
    {0x1.6c16c16c16c17p+0, -0x1.68ac83e9c7000p-2, 0x1.7acd66c548a30p-44},
    {0x1.6a13cd1537290p+0, -0x1.630030b3ab000p-2, 0x1.dbc23e731ae00p-45},
    {0x1.6816816816817p+0, -0x1.5d5bddf596000p-2, 0x1.9de2a08a465dcp-47},
    {0x1.661ec6a5122f9p+0, -0x1.57bf753c8d000p-2, -0x1.fadadee5d40efp-46},
    {0x1.642c8590b2164p+0, -0x1.522ae0738a000p-2, -0x1.eba708164c759p-45},
    {0x1.623fa77016240p+0, -0x1.4c9e09e173000p-2, 0x1.e18891b0ad8a4p-45},
    {0x1.6058160581606p+0, -0x1.4718dc271c000p-2, -0x1.071d8fb4c14c5p-44},
    {0x1.5e75bb8d015e7p+0, -0x1.419b423d5f000p-2, 0x1.ce7a9226de3ecp-44},
    {0x1.5c9882b931057p+0, -0x1.3c25277333000p-2, -0x1.83454b606bd5cp-46},
    {0x1.5ac056b015ac0p+0, -0x1.36b6776be1000p-2, -0x1.15ecdb0f177c8p-46},
    {0x1.58ed2308158edp+0, -0x1.314f1e1d36000p-2, 0x1.8e5bad3213cb8p-45},
    {0x1.571ed3c506b3ap+0, -0x1.2bef07cdc9000p-2, -0x1.aa5ba4a5004f4p-45},
    {0x1.5555555555555p+0, -0x1.269621134e000p-2, 0x1.1ba1f10522625p-44},
    {0x1.5390948f40febp+0, -0x1.214456d0ec000p-2, 0x1.cac5428b728a3p-44},
    {0x1.51d07eae2f815p+0, -0x1.1bf99635a7000p-2, 0x1.1ade9575c2125p-44},
    {0x1.5015015015015p+0, -0x1.16b5ccbad0000p-2, 0x1.232a9042d74bfp-44},
    {0x1.4e5e0a72f0539p+0, -0x1.1178e8227e000p-2, -0x1.1e9b8ce2d07f2p-44},
    {0x1.4cab88725af6ep+0, -0x1.0c42d67616000p-2, -0x1.70d4b163ceae9p-45},
    {0x1.4afd6a052bf5bp+0, -0x1.07138604d6000p-2, 0x1.e70124e912b17p-44},
    {0x1.49539e3b2d067p+0, -0x1.01eae5626c000p-2, -0x1.a44ecfade85aep-44},
    {0x1.47ae147ae147bp+0, -0x1.f991c6cb3c000p-3, 0x1.90b84cd7cc834p-44},
    {0x1.460cbc7f5cf9ap+0, -0x1.ef5ade4dd0000p-3, 0x1.ad11565bb8e11p-51},
    {0x1.446f86562d9fbp+0, -0x1.e530effe72000p-3, 0x1.fdafbb13f7c18p-44},
    {0x1.42d6625d51f87p+0, -0x1.db13db0d48000p-3, -0x1.2813a847527e6p-44},
    {0x1.4141414141414p+0, -0x1.d1037f2656000p-3, 0x1.8527e75b6f6e4p-47},
    {0x1.3fb013fb013fbp+0, -0x1.c6ffbc6f00000p-3, -0x1.ee128d3a69d43p-44},
    {0x1.3e22cbce4a902p+0, -0x1.bd087383be000p-3, 0x1.d5844595412b6p-45},
    {0x1.3c995a47babe7p+0, -0x1.b31d8575bc000p-3, -0x1.c75de562a63cbp-44},
    {0x1.3b13b13b13b14p+0, -0x1.a93ed3c8ae000p-3, 0x1.86a4350562169p-45},
    {0x1.3991c2c187f63p+0, -0x1.9f6c40708a000p-3, 0x1.33aa94bcd3f43p-44},
    {0x1.3813813813814p+0, -0x1.95a5adcf70000p-3, -0x1.8262858a0ff6fp-47},
    {0x1.3698df3de0748p+0, -0x1.8beafeb390000p-3, 0x1.71154aae92cd1p-47},
    {0x1.3521cfb2b78c1p+0, -0x1.823c16551a000p-3, -0x1.e02db9a631e83p-46},
    {0x1.33ae45b57bcb2p+0, -0x1.7898d85444000p-3, -0x1.8e81be3dbaf3fp-44},
    {0x1.323e34a2b10bfp+0, -0x1.6f0128b756000p-3, -0x1.571d90d31ef0fp-44},
    {0x1.30d190130d190p+0, -0x1.6574ebe8c2000p-3, 0x1.98d1d34f0f462p-44},
    {0x1.2f684bda12f68p+0, -0x1.5bf406b544000p-3, 0x1.28023eb68981cp-46},
    {0x1.2e025c04b8097p+0, -0x1.527e5e4a1c000p-3, 0x1.4e61b8d4b411dp-44},
    {0x1.2c9fb4d812ca0p+0, -0x1.4913d8333c000p-3, 0x1.53a43558124c4p-44},
    {0x1.2b404ad012b40p+0, -0x1.3fb45a5992000p-3, -0x1.19313c0cae559p-44},
    {0x1.29e4129e4129ep+0, -0x1.365fcb015a000p-3, 0x1.fd720afb9691bp-44},
    {0x1.288b01288b013p+0, -0x1.2d1610c868000p-3, -0x1.3d0eccb81b4a1p-47},
    {0x1.27350b8812735p+0, -0x1.23d712a49c000p-3, -0x1.00aa38fd3df5cp-46},
    {0x1.25e22708092f1p+0, -0x1.1aa2b7e240000p-3, 0x1.1ad48dde3b366p-44},
    {0x1.2492492492492p+0, -0x1.1178e8227e000p-3, -0x1.1e778ce2d07f2p-45},
    {0x1.23456789abcdfp+0, -0x1.08598b59e4000p-3, 0x1.7e5fd7009902cp-45},
    {0x1.21fb78121fb78p+0, -0x1.fe89139dbc000p-4, -0x1.56494d82f7a82p-44},
    {0x1.20b470c67c0d9p+0, -0x1.ec739830a0000p-4, -0x1.1267ba80cdd10p-44},
    {0x1.1f7047dc11f70p+0, -0x1.da72763844000p-4, -0x1.a79401fa71733p-46},
    {0x1.1e2ef3b3fb874p+0, -0x1.c885801bc4000p-4, -0x1.63f51c65aacd3p-45},
    {0x1.1cf06ada2811dp+0, -0x1.b6ac88dad4000p-4, -0x1.b1cbff50225c7p-44},
    {0x1.1bb4a4046ed29p+0, -0x1.a4e7640b1c000p-4, 0x1.e4336b94407c8p-47},
    {0x1.1a7b9611a7b96p+0, -0x1.9335e5d594000p-4, -0x1.30f5c3abd47dap-45},
    {0x1.19453808ca29cp+0, -0x1.8197e2f410000p-4, 0x1.c102460d20041p-44},
    {0x1.1811811811812p+0, -0x1.700d30aeac000p-4, -0x1.d068da99ded32p-49},
    {0x1.16e0689427379p+0, -0x1.5e95a4d978000p-4, -0x1.1ccace1d17171p-44},
    {0x1.15b1e5f75270dp+0, -0x1.4d3115d208000p-4, 0x1.53e2582f4e1efp-48},
    {0x1.1485f0e0acd3bp+0, -0x1.3bdf5a7d20000p-4, 0x1.1a1e0ad125895p-44},
    {0x1.135c81135c811p+0, -0x1.2aa04a4470000p-4, -0x1.7a16ba8b1cb41p-44},
    {0x1.12358e75d3033p+0, -0x1.1973bd1464000p-4, -0x1.560a154f930b3p-44},
    {0x1.1111111111111p+0, -0x1.08598b59e4000p-4, 0x1.7e9dd7009902cp-46},
    {0x1.0fef010fef011p+0, -0x1.eea31c0068000p-5, -0x1.c3de83606d891p-44},
    {0x1.0ecf56be69c90p+0, -0x1.ccb73cddd8000p-5, -0x1.967c36e09f5fep-44},
    {0x1.0db20a88f4696p+0, -0x1.aaef2d0fb0000p-5, -0x1.1085a353bb42ep-45},
    {0x1.0c9714fbcda3bp+0, -0x1.894aa149f8000p-5, -0x1.9a55a8be97661p-44},
    {0x1.0b7e6ec259dc8p+0, -0x1.67c94f2d48000p-5, -0x1.db2a0827cca0cp-44},
    {0x1.0a6810a6810a7p+0, -0x1.466aed42e0000p-5, 0x1.c073375bdfd28p-45},
    {0x1.0953f39010954p+0, -0x1.252f32f8d0000p-5, -0x1.8401ae021b67bp-45},
    {0x1.0842108421084p+0, -0x1.0415d89e78000p-5, 0x1.ddfc7f461c516p-44},
    {0x1.073260a47f7c6p+0, -0x1.c63d2ec150000p-6, 0x1.54a3ce030a687p-44},
    {0x1.0624dd2f1a9fcp+0, -0x1.8492528c90000p-6, 0x1.a9dba325a0c34p-45},
    {0x1.05197f7d73404p+0, -0x1.432a925980000p-6, -0x1.97739928637fep-47},
    {0x1.0410410410410p+0, -0x1.0205658930000p-6, -0x1.60dd27c8e8417p-44},
    {0x1.03091b51f5e1ap+0, -0x1.82448a3880000p-7, -0x1.4506412c584e0p-44},
    {0x1.0204081020408p+0, -0x1.0101575880000p-7, -0x1.bcd251998b506p-44},
    {0x1.0101010101010p+0, -0x1.0080559580000p-8, -0x1.164afcb31c67bp-45},
    {0x1.0000000000000p+0, 0x0.0p+0, 0x0.0p+0},
    {0x1.fc07f01fc07f0p-1, 0x1.fe02a6b100000p-8, 0x1.9e63f0dda40e4p-46},
    {0x1.f81f81f81f820p-1, 0x1.fc0a8b0fc0000p-7, 0x1.e1e7cf6d3a69cp-50},
    {0x1.f44659e4a4271p-1, 0x1.7b91b07d60000p-6, -0x1.3b685b602ace4p-44},
    {0x1.f07c1f07c1f08p-1, 0x1.f829b0e780000p-6, 0x1.97c267c7e09e4p-45},
    {0x1.ecc07b301ecc0p-1, 0x1.39e87b9fe8000p-5, 0x1.eb3d480ad9015p-44},
    {0x1.e9131abf0b767p-1, 0x1.77458f6330000p-5, -0x1.1807ce586af09p-44},
    {0x1.e573ac901e574p-1, 0x1.b42dd71198000p-5, -0x1.c8d7ae5d6704cp-46},
    {0x1.e1e1e1e1e1e1ep-1, 0x1.f0a30c0118000p-5, -0x1.d579e83368e91p-45},
    {0x1.de5d6e3f8868ap-1, 0x1.16536eea38000p-4, -0x1.472de768fa309p-46},
    {0x1.dae6076b981dbp-1, 0x1.341d7961bc000p-4, 0x1.1cfb299837610p-44},
    {0x1.d77b654b82c34p-1, 0x1.51b073f060000p-4, 0x1.83ba9278e686ap-44},
    {0x1.d41d41d41d41dp-1, 0x1.6f0d28ae58000p-4, -0x1.4b2241b664613p-44},
    {0x1.d0cb58f6ec074p-1, 0x1.8c345d6318000p-4, 0x1.b22b5acb42a66p-44},
    {0x1.cd85689039b0bp-1, 0x1.a926d3a4ac000p-4, 0x1.561c50bd22a9cp-44},
    {0x1.ca4b3055ee191p-1, 0x1.c5e548f5bc000p-4, 0x1.d0c97585fbe06p-46},
    {0x1.c71c71c71c71cp-1, 0x1.e27076e2b0000p-4, -0x1.a2c2c2af0003cp-45},
    {0x1.c3f8f01c3f8f0p-1, 0x1.fec9131dc0000p-4, -0x1.54455d1ae6607p-44},
    {0x1.c0e070381c0e0p-1, 0x1.0d77e7cd08000p-3, 0x1.cb6cd2ee2f482p-44},
    {0x1.bdd2b899406f7p-1, 0x1.1b72ad52f6000p-3, 0x1.e86041811a396p-45},
    {0x1.bacf914c1bad0p-1, 0x1.29552f8200000p-3, -0x1.5bd67f4471dfcp-44},
    {0x1.b7d6c3dda338bp-1, 0x1.371fc201e8000p-3, 0x1.eea079b2d8abcp-44},
    {0x1.b4e81b4e81b4fp-1, 0x1.44d2b6ccb8000p-3, -0x1.71f416135783cp-46},
    {0x1.b2036406c80d9p-1, 0x1.526e5e3a1c000p-3, -0x1.790aa37fc5238p-44},
    {0x1.af286bca1af28p-1, 0x1.5ff3070a7a000p-3, -0x1.8546f183bebf2p-44},
    {0x1.ac5701ac5701bp-1, 0x1.6d60fe719e000p-3, -0x1.bc91557134767p-44},
    {0x1.a98ef606a63bep-1, 0x1.7ab890210e000p-3, -0x1.be51072534a58p-45},
    {0x1.a6d01a6d01a6dp-1, 0x1.87fa06520c000p-3, 0x1.22130401202fcp-44},
    {0x1.a41a41a41a41ap-1, 0x1.9525a9cf46000p-3, -0x1.294937d9f158fp-44},
    {0x1.a16d3f97a4b02p-1, 0x1.a23bc1fe2c000p-3, -0x1.53d6d91dc9f0bp-44},
    {0x1.9ec8e951033d9p-1, 0x1.af3c94e80c000p-3, -0x1.92e633fcd9066p-52},
    {0x1.9c2d14ee4a102p-1, 0x1.bc286742d8000p-3, 0x1.9a873f39d121cp-44},
    {0x1.999999999999ap-1, 0x1.c8ff7c79aa000p-3, -0x1.7814f689f8434p-45},
    {0x1.970e4f80cb872p-1, 0x1.d5c216b4fc000p-3, -0x1.1b0d1bbca681bp-45},
    {0x1.948b0fcd6e9e0p-1, 0x1.e27076e2b0000p-3, -0x1.a302c2af0003cp-44},
    {0x1.920fb49d0e229p-1, 0x1.ef0adcbdc6000p-3, -0x1.b2a179c86af24p-45},
    {0x1.8f9c18f9c18fap-1, 0x1.fb9186d5e4000p-3, -0x1.d6b2aab993c87p-47},
    {0x1.8d3018d3018d3p-1, 0x1.0402594b4d000p-2, 0x1.037b89ef42d7fp-48},
    {0x1.8acb90f6bf3aap-1, 0x1.0a324e2739000p-2, 0x1.c4dee7ef4030ep-47},
    {0x1.886e5f0abb04ap-1, 0x1.1058bf9ae5000p-2, -0x1.4affd817d52cdp-44},
    {0x1.8618618618618p-1, 0x1.1675cababa000p-2, 0x1.83c0e731f55c4p-44},
    {0x1.83c977ab2beddp-1, 0x1.1c898c169a000p-2, -0x1.81260e5c62affp-44},
    {0x1.8181818181818p-1, 0x1.22941fbcf8000p-2, -0x1.a6876f5eb0963p-44},
    {0x1.7f405fd017f40p-1, 0x1.2895a13de8000p-2, 0x1.a917ad24c13f0p-44},
    {0x1.7d05f417d05f4p-1, 0x1.2e8e2bae12000p-2, -0x1.6791e99b72bd8p-45},
    {0x1.7ad2208e0ecc3p-1, 0x1.347dd9a988000p-2, -0x1.5522dd4c58092p-45},
    {0x1.78a4c8178a4c8p-1, 0x1.3a64c55694000p-2, 0x1.7a81cbcd735d0p-44},
    {0x1.767dce434a9b1p-1, 0x1.404308686a000p-2, 0x1.f8f043049f7d3p-44},
    {0x1.745d1745d1746p-1, 0x1.4618bc21c6000p-2, -0x1.3e02f484c84ccp-46},
};
__device__ __forceinline__ double gvi_log(double x) {
  const long long ix = __double_as_longlong(x);
  // subnormal, zero, negative, inf, NaN: (unsigned)(ix - 2^52) >= 0x7ff0... - 2^52
  if ((unsigned long long)(ix - 0x0010000000000000ll) >= 0x7fe0000000000000ull) return log(x);
  const long long tmp = ix - 0x3fe6100000000000ll;  // OFF = 0x3fe6000000000000 + 2^44
  const int i = (int)(tmp >> 45) & 127;
  const int k = (int)(tmp >> 52);
  const double z = __longlong_as_double(ix - (tmp & (long long)0xfff0000000000000ull));
  const double invc = GVI_LOG_TABLE[i][0], logc_hi = GVI_LOG_TABLE[i][1], logc_lo = GVI_LOG_TABLE[i][2];
  const double r = fma(z, invc, -1.0);
  const double kd = (double)k;
  const double hi = fma(kd, 0x1.62e42fefa3800p-1, logc_hi);   // exact
  const double lo = fma(kd, 0x1.ef35793c76730p-45, logc_lo);
  const double r2 = r * r;
  double p = 1.0 / 7.0;
  p = fma(p, r, -1.0 / 6.0);
  p = fma(p, r, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  return (hi + r) + fma(r2, p, lo);
}

// pointwise loss terms, shared by risk.cu and predict.cu ---------------------------------------------------------
// Same quantities as the reference's formulas, arranged for the FP64 pipe (a sqrt is ~10, a divide ~9, a log 14
// instructions): log sqrt(v) = log(v) / 2, log(sqrt(a) / sqrt(b)) = log(a / b) / 2, sqrt(a) sqrt(b) = sqrt(a b).
__device__ __forceinline__ double gvi_nll_term(double y, double m, double v) {
  // -norm.logpdf(y, loc=m, scale=sqrt(v))  (negative_log_likelihood.py:44-48)
  const double half_log_2pi = 0.91893853320467274178;
  const double r = y - m;
  return 0.5 * (r * r / v + gvi_log(v)) + half_log_2pi;
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
    case GVI_REG_PROJ_BHATTACHARYYA: {
      // dm^2 / (8 s) + log((s / 2) / sqrt(cp cq)) / 2 = dm^2 / (8 s) + log(s^2 / (4 cp cq)) / 4
      const double s = cp + cq;
      return 0.125 * dm * dm / s + 0.25 * gvi_log(0.25 * s * s / (cp * cq));
    }
    case GVI_REG_PROJ_GAUSSIAN_WASSERSTEIN:
      return dm * dm + cp + cq - 2.0 * sqrt(cp * cq);
    case GVI_REG_PROJ_HELLINGER: {
      const double s = cp + cq;
      return 1.0 - sqrt(2.0 * sqrt(cp * cq) / s) * exp(-0.25 * dm * dm / s);
    }
    case GVI_REG_PROJ_KL: {
      const double icq = 1.0 / cq;
      return 0.5 * ((cp + dm * dm) * icq - gvi_log(cp * icq) - 1.0);
    }
    case GVI_REG_PROJ_RENYI: {
      const double mix = alpha * cp + (1.0 - alpha) * cq;
      const double imix = 1.0 / mix;
      const double ia = 1.0 / (alpha - 1.0);  // loop invariant where the call is inlined
      return 0.5 * (gvi_log(cp / cq) + gvi_log(cp * imix) * ia + alpha * dm * dm * imix);
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
