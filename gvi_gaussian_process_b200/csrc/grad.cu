// grad.cu - the M x M part of the backward pass of the fused step, and the gradient entry point.
//
// With g_i = dR/dv_q,i (from the tile kernel) and a_i = A^-1 k_i:
//   dR/dlog_ls[d]  = sum_i g_i sum_j a_ij k_ij (x~_id - z~_jd)^2                       (tile kernel, T2_d)
//                  + < dK_ZZ/dlog_ls[d], S >,   S = sum_i g_i a_i a_i^T
//   dR/dlog_scaling = sum_i g_i (2 v_i - 2 jitter_abs |a_i|^2)                          (tile kernel)
// S is a weighted Gram (SYRK) over ALL points.  The tile kernel leaves the a_i of a chunk of points in a ring, already in
// DMMA operand order, and syrk_kernel accumulates 128x128 tiles of S on the tensor pipe, chunk after chunk; the
// contraction with dK_ZZ/dlog_ls is a small reduction kernel.  (S is accumulated from the a_i, the way reverse-mode
// differentiation of cho_solve does it: its error is ~cond(A) eps.  The algebraically equal A^-1 (sum g k k^T) A^-1
// amplifies the rounding of the inner sum by cond(A)^2 and was dropped.)  The same SYRK kernel fed by kgen_kernel
// gives C = sum_i g_i k_i k_i^T, the vector-Jacobian product of the SVGP family (gvi_weighted_gram_f64).
// Reference: jax.grad of generalised_variational_inference.py:73-94 (trainer.py:83-89).
#include "dgemm.cuh"

#include "predict.cuh"

size_t gvi_gp_predict_grad_wide_workspace_bytes(int64_t N, int M, int D);
int gvi_gp_predict_grad_wide(const double* F, int M, int D, const double* X, int64_t N, const double* g, int frozen,
                             double* out, void* workspace, cudaStream_t st);

namespace {
constexpr int CB = SYRK_CB;  // C tile edge
constexpr int PC = SYRK_PC;  // points per operand block
constexpr int STHREADS = 256;

// operand block layout: predict.cuh (the tile kernel writes a_i blocks, kgen_kernel writes k_i blocks)
constexpr int KPSTRIDE = SYRK_KPSTRIDE;
constexpr int STAGE = SYRK_STAGE;
constexpr int BLK = SYRK_BLK;                 // doubles per (column block, point group)
constexpr uint32_t BLK_BYTES = BLK * 8;       // 33152, a multiple of 16 (cp.async.bulk granularity)
constexpr int NSTAGE = 3;
__device__ __forceinline__ int frag_index(int c, int i) { return syrk_frag_index(c, i); }

// ---- stage 1: K(x_chunk, Z) once per chunk into an L2-sized ring, already in fragment order ------------------------
// One CTA per (32-point group, 128-column block): lane = point, warp w takes 16 columns, four at a time (independent
// exp chains); 48 registers -> several CTAs per SM keep the FP64 pipe busy.
template <int DR>
__global__ void __launch_bounds__(STHREADS) kgen_kernel(const double* __restrict__ F, const double* __restrict__ X,
                                                        const double* __restrict__ g, int64_t N, int64_t row0,
                                                        int npg, int M, int Mp, int D, double* __restrict__ Kc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pg = blockIdx.x, cb = blockIdx.y;
  const FactorLayout f = factor_layout(M, D);
  const double amp2 = F[0];
  const double* __restrict__ sqrtw = F + f.sqrtw_off;
  const double* __restrict__ Zs = F + f.zs_off;
  const int64_t r = row0 + (int64_t)pg * PC + lane;
  double xr[DR];
#pragma unroll
  for (int d = 0; d < DR; d++) xr[d] = (d < D && r < N) ? X[r * D + d] * sqrtw[d] : 0.0;
  const double live = (r < N) ? amp2 : 0.0;
  double* blk = Kc + ((size_t)cb * npg + pg) * BLK;
  if (warp == 0) blk[STAGE + lane] = (r < N) ? g[r] : 0.0;
  for (int c0 = cb * CB + warp * 16; c0 < cb * CB + warp * 16 + 16; c0 += 4) {
    double s4[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = c0 + q;
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < DR; d++) {
        double z = (d < D && c < Mp) ? Zs[(int64_t)c * D + d] : 0.0;
        double diff = xr[d] - z;
        s = fma(diff, diff, s);
      }
      s4[q] = s;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = c0 + q;
      const double k = (c < M ? live : 0.0) * gvi_exp(-0.5 * s4[q]);
      blk[frag_index(c % CB, lane)] = k;
    }
  }
}

// ---- mbarrier / bulk-copy (TMA 1-D) helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

struct SyrkParams {
  const double* Kc;  // [ncb][npg][BLK]
  int npg;           // point groups in this chunk
  int Mp;
  int nsplit;
  int first;         // first chunk: accumulators start at zero, otherwise they are reloaded from Cpart
  double* Cpart;     // [nsplit][Mp*Mp]
};

// ---- stage 2: C tile += (G K_j)^T K_l over the chunk.  One CTA = one 128x128 lower-triangular tile of C x one slice of
// the chunk's point groups.  K blocks arrive by cp.async.bulk (TMA 1-D) into a 3-stage shared-memory ring, completion
// on mbarriers; every warp issues 256 DMMAs per 32-point group for its 32x64 sub-tile; accumulators stay in registers.
__global__ void __launch_bounds__(STHREADS, 1) syrk_kernel(const SyrkParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);                       // [NSTAGE][2][BLK]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)NSTAGE * 2 * BLK);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int bj = 0, rem = blockIdx.x;
  while (rem > bj) { rem -= bj + 1; bj++; }
  const int bl = rem;
  const bool diag = (bj == bl);
  const int pg_begin = (int)((int64_t)P.npg * blockIdx.y / P.nsplit);
  const int pg_end = (int)((int64_t)P.npg * (blockIdx.y + 1) / P.nsplit);
  const int niter = pg_end - pg_begin;
  const int wj = warp >> 1, wl = warp & 1;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int it) {  // thread 0 only
    const int s = it % NSTAGE;
    double* dstA = ring + (size_t)s * 2 * BLK;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&full[s], diag ? BLK_BYTES : 2 * BLK_BYTES);
    bulk_g2s(dstA, P.Kc + ((size_t)bj * P.npg + pg_begin + it) * BLK, BLK_BYTES, &full[s]);
    if (!diag) bulk_g2s(dstA + BLK, P.Kc + ((size_t)bl * P.npg + pg_begin + it) * BLK, BLK_BYTES, &full[s]);
  };
  if (tid == 0)
    for (int it = 0; it < NSTAGE && it < niter; it++) issue(it);

  double acc[4][8][2];
  double* Cout = P.Cpart + (size_t)blockIdx.y * P.Mp * P.Mp;
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int r = bj * CB + (4 * wj + mt) * 8 + (lane >> 2);
        int c = bl * CB + (8 * wl + nt) * 8 + (lane & 3) * 2 + e;
        acc[mt][nt][e] = (!P.first && r < P.Mp && c < P.Mp) ? Cout[(size_t)r * P.Mp + c] : 0.0;
      }

  for (int it = 0; it < niter; it++) {
    const int s = it % NSTAGE;
    mbar_wait(&full[s], (it / NSTAGE) & 1);
    const double* Ab = ring + (size_t)s * 2 * BLK;
    const double* Bb = diag ? Ab : Ab + BLK;
    const double2* A2 = reinterpret_cast<const double2*>(Ab);
    const double2* B2 = reinterpret_cast<const double2*>(Bb);
#pragma unroll
    for (int kp = 0; kp < PC / 8; kp++) {
      const double g0 = Ab[STAGE + kp * 8 + (lane & 3)], g1 = Ab[STAGE + kp * 8 + 4 + (lane & 3)];
      double2 a[4], b[8];
#pragma unroll
      for (int mt = 0; mt < 4; mt++) {
        a[mt] = A2[kp * (KPSTRIDE / 2) + (4 * wj + mt) * 32 + lane];
        a[mt].x *= g0;  // the weight g_i rides on the A operand: (G K_j)^T K_l
        a[mt].y *= g1;
      }
#pragma unroll
      for (int nt = 0; nt < 8; nt++) b[nt] = B2[kp * (KPSTRIDE / 2) + (8 * wl + nt) * 32 + lane];
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
    }
    __syncthreads();  // every warp is done with stage s before it is refilled
    if (tid == 0 && it + NSTAGE < niter) issue(it + NSTAGE);
  }
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int r = bj * CB + (4 * wj + mt) * 8 + (lane >> 2);
        int c = bl * CB + (8 * wl + nt) * 8 + (lane & 3) * 2 + e;
        if (r < P.Mp && c < P.Mp) Cout[(size_t)r * P.Mp + c] = acc[mt][nt][e];
      }
}

// C = sum over splits (fixed order), mirrored to the full symmetric matrix (tiles with bj >= bl were computed)
__global__ void syrk_reduce_kernel(const double* __restrict__ Cpart, int nsplit, int Mp, double* __restrict__ C) {
  const int64_t total = (int64_t)Mp * Mp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(idx / Mp), c = (int)(idx % Mp);
    int64_t src = (r / CB >= c / CB) ? idx : (int64_t)c * Mp + r;
    double s = 0.0;
    for (int k = 0; k < nsplit; k++) s += Cpart[(size_t)k * total + src];
    C[idx] = s;
  }
}

// term3_d = sum_{j,l} S_jl * K_jl * (-1/2) (z~_jd - z~_ld)^2 ; one CTA per row j, fixed-order reductions
__global__ void __launch_bounds__(256) dkzz_contract_kernel(const double* __restrict__ S, const double* __restrict__ F,
                                                            int M, int Mp, int D, double* __restrict__ rowpart) {
  __shared__ double red[8];
  __shared__ double zj[32];
  const FactorLayout f = factor_layout(M, D);
  const double* __restrict__ Zs = F + f.zs_off;
  const int j = blockIdx.x, tid = threadIdx.x;
  if (tid < D) zj[tid] = Zs[(int64_t)j * D + tid];
  __syncthreads();
  const double amp2 = F[0];
  double t[32];
#pragma unroll
  for (int d = 0; d < 32; d++) t[d] = 0.0;
  for (int l = tid; l < M; l += 256) {
    double s = 0.0, diff2[32];
#pragma unroll
    for (int d = 0; d < 32; d++) {
      if (d < D) {
        double diff = zj[d] - Zs[(int64_t)l * D + d];
        diff2[d] = diff * diff;
        s += diff2[d];
      } else {
        diff2[d] = 0.0;
      }
    }
    const double coef = -0.5 * S[(size_t)j * Mp + l] * amp2 * gvi_exp(-0.5 * s);
#pragma unroll
    for (int d = 0; d < 32; d++) t[d] = fma(coef, diff2[d], t[d]);
  }
  for (int d = 0; d < D; d++) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 32; q++) v = (q == d) ? t[q] : v;  // static indexing keeps t[] in registers
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double a = 0.0;
      for (int w = 0; w < 8; w++) a += red[w];
      rowpart[(size_t)j * 32 + d] = a;
    }
  }
}

// out[4+d] += sum_j rowpart[j][d]  (fixed order)
__global__ void grad_finish_kernel(const double* __restrict__ rowpart, int M, int D, double* __restrict__ out) {
  const int d = threadIdx.x;
  if (d >= D) return;
  double a = 0.0;
  for (int j = 0; j < M; j++) a += rowpart[(size_t)j * 32 + d];
  out[4 + d] += a;
}

template <int DR>
int launch_kgen(const double* F, const double* X, const double* g, int64_t N, int64_t row0, int npg, int M, int Mp,
                int D, double* Kc, cudaStream_t st) {
  kgen_kernel<DR><<<dim3(npg, (Mp + CB - 1) / CB), STHREADS, 0, st>>>(F, X, g, N, row0, npg, M, Mp, D, Kc);
  GVI_CHECK_LAUNCH("kgen_kernel");
  return GVI_OK;
}

int launch_syrk(const SyrkParams& P, int ntri, cudaStream_t st) {
  size_t smem = (size_t)NSTAGE * 2 * BLK * sizeof(double) + NSTAGE * sizeof(uint64_t) + 64;
  static std::atomic<unsigned long long> configured{0};
  if (gvi_first_use_on_this_device(configured))
    GVI_CUDA(cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  syrk_kernel<<<dim3(ntri, P.nsplit), STHREADS, smem, st>>>(P);
  GVI_CHECK_LAUNCH("syrk_kernel");
  return GVI_OK;
}

#ifdef GVI_TEST_HOOKS
int g_chunk_waves_cap = 0;  // libgvigp_test.so only: gvi_test_set_grad_chunk_waves (include/gvigp_test.h)
#else
constexpr int g_chunk_waves_cap = 0;
#endif

struct GradWs {
  double *partials, *tile_partials, *g, *bscratch, *Cpart, *C, *rowpart, *Kc;
  double *mp, *vp, *pscratch;  // regulariser predictive of one chunk (forward launch) and that launch's panel scratch
  int* tile_counter;
  int64_t ntiles;
  int nsplit, ntri, ncb, chunk_pg;  // chunk_pg = point groups (of 32 points) per K chunk
  int64_t achunk_pts;               // points per chunk of the a_i ring (a multiple of every tile height and of 32)
};
size_t align256(size_t x) { return (x + 255) / 256 * 256; }
GradWs carve_grad(void* ws, int64_t N, int M, int D, size_t* total) {
  const FactorLayout f = factor_layout(M, D);
  GradWs w;
  const int Gc = (f.Mp + CB - 1) / CB;
  w.ntri = Gc * (Gc + 1) / 2;
  w.nsplit = GVI_NUM_SMS / w.ntri > 0 ? GVI_NUM_SMS / w.ntri : 1;
  const int T = 24;  // upper bound of the tile height (the b scratch is sized for the largest variant)
  unsigned char* p = (unsigned char*)ws;
  size_t o = 0;
  auto take = [&](size_t bytes) { double* r = (double*)(p + o); o += align256(bytes); return r; };
  w.tile_counter = (int*)take(STEP_COUNTER_DOUBLES * 8);
  w.partials = take((size_t)(GVI_NUM_SMS > STEP_REDUCE_BLOCKS ? GVI_NUM_SMS : STEP_REDUCE_BLOCKS) * STEP_PARTIAL_STRIDE * 8);
  {  // per-tile records of the gradient launch (tile height as the launcher will choose it)
    int tp = gvi_step_tile_points(f.Mp, D, true);
    if (tp <= 0) tp = 8;
    w.ntiles = (N + tp - 1) / tp;
    w.tile_partials = take((size_t)w.ntiles * STEP_PARTIAL_STRIDE * 8);
  }
  w.g = take((size_t)N * 8);
  w.bscratch = take((size_t)4 * GVI_NUM_SMS * f.Mp * T * 8);  // per CTA (up to two per SM): the b tile and q's parked K tile
  w.Cpart = take((size_t)w.nsplit * f.Mp * f.Mp * 8);
  w.C = take((size_t)f.Mp * f.Mp * 8);
  w.rowpart = take((size_t)f.Mp * 32 * 8);
  // K ring: ~48 MB so that a chunk written by kgen_kernel is still L2-resident when syrk_kernel streams it
  w.ncb = Gc;
  int64_t pgs = (48LL << 20) / ((int64_t)Gc * BLK * 8);
  pgs = pgs / w.nsplit * w.nsplit;
  if (pgs < w.nsplit) pgs = w.nsplit;
  const int64_t need = (N + PC - 1) / PC;
  if (pgs > need) pgs = need;
  w.chunk_pg = (int)pgs;
  // the same ring takes the a_i chunks of the tile kernel (gradient step).  That accumulation is compute bound and
  // streaming its operands from HBM costs nothing, so the chunks are long (up to 1 GB) to amortise the launch tails;
  // a chunk is a whole number of 148 x 24-point waves (3552 = 32 x 111 points: a multiple of every tile height).
  const int64_t wave = (int64_t)GVI_NUM_SMS * 24;
  const int64_t wave_bytes = wave / PC * Gc * BLK * 8;
  int64_t nw = (1LL << 30) / wave_bytes;  // (2 and 4 GB rings measured: no difference, 141.4 - 141.8 ms per step)
  if (g_chunk_waves_cap > 0 && nw > g_chunk_waves_cap) nw = g_chunk_waves_cap;
  const int64_t need_w = (N + wave - 1) / wave;
  if (nw > need_w) nw = need_w;
  if (nw < 1) nw = 1;
  w.achunk_pts = nw * wave;
  // (wide inputs, D > 32, take the GEMM route of grad_wide.cu and never fill the a_i ring)
  const size_t ring_k = (size_t)Gc * pgs * BLK * 8, ring_a = D > 32 ? 0 : (size_t)nw * wave_bytes;
  w.Kc = take(ring_k > ring_a ? ring_k : ring_a);
  w.mp = take((size_t)w.achunk_pts * 8);
  w.vp = take((size_t)w.achunk_pts * 8);
  w.pscratch = take(gvi_step_scratch_doubles(f.Mp) * 8);
  if (total) *total = o;
  return w;
}
}  // namespace

#ifdef GVI_TEST_HOOKS  // libgvigp_test.so only (include/gvigp_test.h): never part of the product library
namespace {
__global__ void test_exp_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = gvi_exp(x[i]);
}
}  // namespace
extern "C" int gvi_test_set_grad_chunk_waves(int waves) {
  const int prev = g_chunk_waves_cap;
  g_chunk_waves_cap = waves > 0 ? waves : 0;
  return prev;
}
extern "C" int gvi_test_exp_f64(const double* x, int64_t n, double* out, void* stream) {
  GVI_CHECK_ARG(x && out && n >= 1, "arguments");
  test_exp_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  GVI_CHECK_LAUNCH("test_exp_kernel");
  return GVI_OK;
}
#endif  // GVI_TEST_HOOKS

// C = sum_i g_i k_i k_i^T (symmetric, full matrix in w.C), chunk by chunk through the L2-sized K ring
static int gvi_weighted_gram(const double* Fq, const double* X, const double* g, int64_t N, int M, int D,
                             const GradWs& w, cudaStream_t st) {
  const FactorLayout f = factor_layout(M, D);
  int rc = GVI_OK;
  // C = sum_i g_i k_i k_i^T, chunk by chunk through the L2-sized K ring
  const int64_t total_pg = (N + PC - 1) / PC;
  for (int64_t pg0 = 0; pg0 < total_pg; pg0 += w.chunk_pg) {
    const int npg = (int)(total_pg - pg0 < w.chunk_pg ? total_pg - pg0 : w.chunk_pg);
    const int64_t row0 = pg0 * PC;
    if (D <= 1) rc = launch_kgen<1>(Fq, X, g, N, row0, npg, M, f.Mp, D, w.Kc, st);
    else if (D <= 4) rc = launch_kgen<4>(Fq, X, g, N, row0, npg, M, f.Mp, D, w.Kc, st);
    else if (D <= 8) rc = launch_kgen<8>(Fq, X, g, N, row0, npg, M, f.Mp, D, w.Kc, st);
    else if (D <= 16) rc = launch_kgen<16>(Fq, X, g, N, row0, npg, M, f.Mp, D, w.Kc, st);
    else rc = launch_kgen<32>(Fq, X, g, N, row0, npg, M, f.Mp, D, w.Kc, st);
    if (rc) return rc;
    SyrkParams sp{w.Kc, npg, f.Mp, w.nsplit, pg0 == 0 ? 1 : 0, w.Cpart};
    if ((rc = launch_syrk(sp, w.ntri, st))) return rc;
  }
  syrk_reduce_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(w.Cpart, w.nsplit, f.Mp, w.C);
  GVI_CHECK_LAUNCH("syrk_reduce_kernel");
  return GVI_OK;
}

// Gradient-mode tile launches chunk by chunk.  Each chunk leaves a_i = A^-1 k_i (and g_i) of its points in the ring in
// operand-block order and syrk_kernel folds it into S = sum_i g_i a_i a_i^T on the tensor pipe (w.C on return).  The
// per-tile records land at their global tile index, so the fixed-order reductions that follow see one launch.
static int gvi_grad_launch_accumulate_S(const StepParams& P0, const GradWs& w, int* grid_out, cudaStream_t st) {
  const FactorLayout f = factor_layout(P0.M, P0.D);
  const int T = gvi_step_tile_points(f.Mp, P0.D, true);
  if (T <= 0) {
    gvi_set_error("the gradient step supports M <= 3456 in this build (got M=%d)", P0.M);
    return GVI_EUNSUPPORTED;
  }
  int rc;
  for (int64_t row0 = 0; row0 < P0.N; row0 += w.achunk_pts) {
    const int64_t n = P0.N - row0 < w.achunk_pts ? P0.N - row0 : w.achunk_pts;
    StepParams P = P0;
    auto at = [&](const double* p) { return p ? p + row0 : nullptr; };
    auto atw = [&](double* p) { return p ? p + row0 : nullptr; };
    P.X = P0.X + row0 * P0.D;
    P.N = n;
    P.y = at(P0.y); P.m_q = at(P0.m_q); P.m_p_in = at(P0.m_p_in); P.v_p_in = at(P0.v_p_in); P.g_in = at(P0.g_in);
    P.g_out = atw(P0.g_out); P.dm_out = atw(P0.dm_out);
    for (int ps = 0; ps < P0.npass; ps++) {
      P.pass[ps].prior_mean = at(P0.pass[ps].prior_mean);
      P.pass[ps].mean_out = atw(P0.pass[ps].mean_out);
      P.pass[ps].var_out = atw(P0.pass[ps].var_out);
    }
    P.tile_partials = P0.tile_partials + (size_t)(row0 / T) * STEP_PARTIAL_STRIDE;
    if (P0.npass == 2) {
      // the regulariser GP p is constant: its predictive needs no gradient machinery, so it runs through the FORWARD
      // launch shape (two CTAs per SM: its kernel evaluations hide under the neighbour's DMMAs) into chunk-sized arrays
      // and the gradient launch takes (m_p, v_p) as inputs, like a prior-mode regulariser
      StepParams Pp{};
      Pp.npass = 1;
      Pp.pass[0] = P.pass[1];
      if (!Pp.pass[0].mean_out) Pp.pass[0].mean_out = w.mp;
      if (!Pp.pass[0].var_out) Pp.pass[0].var_out = w.vp;
      Pp.M = P0.M; Pp.Mp = P0.Mp; Pp.D = P0.D; Pp.G = P0.G;
      Pp.X = P.X; Pp.N = n;
      Pp.tile_points = 24;  // every chunk in the same launch shape (a short last chunk would otherwise take 8-point tiles
                            // on 16 warps, whose mean partial sums are grouped differently): chunking stays bit-invisible
      Pp.tile_counter = P0.tile_counter;
      Pp.acc_scratch = w.pscratch;
      const int gp = gvi_launch_step(Pp, st);
      if (gp < 0) return gp;
      P.npass = 1;
      P.m_p_in = Pp.pass[0].mean_out;
      P.v_p_in = Pp.pass[0].var_out;
    }
    const int64_t ntiles = (n + T - 1) / T;
    const int npg = (int)((ntiles * T + PC - 1) / PC);
    P.a_ring = w.Kc;
    P.a_npg = npg;
    if ((ntiles * T) % PC != 0 || n % T != 0)  // the last group is only partly written by the tiles: zero it first
      GVI_CUDA(cudaMemset2DAsync(w.Kc + (size_t)(npg - 1) * BLK, (size_t)npg * BLK * 8, 0, (size_t)BLK * 8, w.ncb, st));
    const int grid = gvi_launch_step(P, st);
    if (grid < 0) return grid;
    if (grid_out) *grid_out = grid;
    SyrkParams sp{w.Kc, npg, f.Mp, w.nsplit, row0 == 0 ? 1 : 0, w.Cpart};
    if ((rc = launch_syrk(sp, w.ntri, st))) return rc;
  }
  syrk_reduce_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(w.Cpart, w.nsplit, f.Mp, w.C);
  GVI_CHECK_LAUNCH("syrk_reduce_kernel");
  return GVI_OK;
}

// out[4+d] += <dK_ZZ/dlog_ls[d], S> with S in w.C
static int gvi_contract_kzz_term(const double* Fq, const FactorLayout& f, int M, int D, const GradWs& w, double* out,
                                 cudaStream_t st) {
  dkzz_contract_kernel<<<M, 256, 0, st>>>(w.C, Fq, M, f.Mp, D, w.rowpart);
  GVI_CHECK_LAUNCH("dkzz_contract_kernel");
  grad_finish_kernel<<<1, 32, 0, st>>>(w.rowpart, M, D, out);
  GVI_CHECK_LAUNCH("grad_finish_kernel");
  return GVI_OK;
}

extern "C" size_t gvi_fused_step_grad_workspace_bytes(int64_t N, int M, int D) {
  if (N < 1 || M < 1 || D < 1) return 0;
  size_t total = 0;
  carve_grad(nullptr, N, M, D, &total);
  if (D > 32 || gvi_step_tile_points(factor_layout(M, D).Mp, D, true) == 0) {  // the GEMM route (grad_wide.cu)
    const size_t wide = gvi_gp_predict_grad_wide_workspace_bytes(N, M, D);
    if (wide > total) total = wide;
  }
  return total;
}

extern "C" int gvi_fused_step_grad_f64(const void* factor_q, const void* factor_p, int M, int D, const double* X,
                                       const double* y, const double* m_q, const double* prior_mean_p,
                                       const double* m_p_in, const double* v_p_in, int64_t N, int reg_kind,
                                       double renyi_alpha, int frozen_cholesky, double* out, double* dm_out,
                                       double* var_q_out, double* mean_p_out, double* var_p_out, void* workspace,
                                       void* stream) {
  GVI_CHECK_ARG(out && dm_out && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  StepParams P{};
  int rc = gvi_fill_step_params(P, factor_q, factor_p, M, D, X, y, m_q, prior_mean_p, m_p_in, v_p_in, N, reg_kind,
                                renyi_alpha, var_q_out, mean_p_out, var_p_out);
  if (rc) return rc;
  GradWs w = carve_grad(workspace, N, M, D, nullptr);
  const FactorLayout f = factor_layout(M, D);
  const double* Fq = (const double*)factor_q;
  P.partials = w.partials;
  P.tile_counter = w.tile_counter;
  P.tile_partials = w.tile_partials;
  P.grad = 1;
  P.g_out = w.g;
  P.dm_out = dm_out;
  P.bscratch = w.bscratch;
  P.acc_scratch = w.pscratch;  // panel scratch of the two-CTA gradient shape (the p launch of a chunk is done with it)
  int grid;
  if (frozen_cholesky) {  // A does not depend on the trained parameters: no <dK_ZZ, S> term, one launch
    if ((grid = gvi_launch_step(P, st)) < 0) return grid;
  } else if ((rc = gvi_grad_launch_accumulate_S(P, w, nullptr, st))) {
    return rc;
  }
  if ((grid = gvi_launch_tile_reduce(w.tile_partials, w.ntiles, 1, w.partials, st)) < 0) return grid;
  if ((rc = gvi_launch_finish(P.partials, grid, N, 1, D, Fq, out, st))) return rc;
  if (frozen_cholesky) return GVI_OK;
  return gvi_contract_kzz_term(Fq, f, M, D, w, out, st);
}


// ---- vector-Jacobian product of the predictive variance (gvi_gp_predict_f64) w.r.t. the kernel hyper-parameters ----
// Given g_i = dLoss/dvar_i for an arbitrary downstream loss (e.g. the classification epilogue of classify.cu, or any
// host-framework graph), returns out[3] = dLoss/dlog_scaling and out[4+d] = dLoss/dlog_lengthscales[d]; out[0] =
// out[1] = 0 and out[2] = N so the vector has the layout of gvi_fused_step_grad_f64.  Same kernels as the fused
// gradient step: the tile kernel in gradient mode (a_i = L^-T L^-1 k_i), the weighted Gram C = sum_i g_i k_i k_i^T,
// S = A^-1 C A^-1 and <dK_ZZ/dtheta, S>.
extern "C" int gvi_gp_predict_grad_f64(const void* factor, int M, int D, const double* X, int64_t N,
                                       const double* g_var, int frozen_cholesky, double* out, void* workspace,
                                       void* stream) {
  GVI_CHECK_ARG(M >= 1 && D >= 1 && N >= 1, "sizes");
  GVI_CHECK_ARG(factor && X && g_var && out && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const FactorLayout f = factor_layout(M, D);
  const double* Fq = (const double*)factor;
  // wide inputs (D > 32) and inducing sets whose b tile does not fit shared memory (M > 3456): the contraction with
  // dK/dlog_ls runs as GEMMs over all d at once (grad_wide.cu) - any D, any M
  if (D > 32 || gvi_step_tile_points(f.Mp, D, true) == 0)
    return gvi_gp_predict_grad_wide(Fq, M, D, X, N, g_var, frozen_cholesky, out, workspace, st);
  GradWs w = carve_grad(workspace, N, M, D, nullptr);
  StepParams P{};
  P.pass[0] = PassDesc{Fq, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  P.npass = 1;
  P.M = M; P.Mp = f.Mp; P.D = D; P.G = f.G;
  P.X = X; P.N = N;
  P.reg_kind = GVI_REG_NONE;
  P.partials = w.partials;
  P.tile_counter = w.tile_counter;
  P.tile_partials = w.tile_partials;
  P.grad = 1;
  P.g_in = g_var;
  P.bscratch = w.bscratch;
  P.acc_scratch = w.pscratch;  // panel scratch of the two-CTA gradient shape (the p launch of a chunk is done with it)
  int grid, rc;
  if (frozen_cholesky) {
    if ((grid = gvi_launch_step(P, st)) < 0) return grid;
  } else if ((rc = gvi_grad_launch_accumulate_S(P, w, nullptr, st))) {
    return rc;
  }
  if ((grid = gvi_launch_tile_reduce(w.tile_partials, w.ntiles, 1, w.partials, st)) < 0) return grid;
  if ((rc = gvi_launch_finish(P.partials, grid, N, 1, D, Fq, out, st))) return rc;
  if (frozen_cholesky) return GVI_OK;
  return gvi_contract_kzz_term(Fq, f, M, D, w, out, st);
}

// ---- weighted Gram C = sum_i g_i k_i k_i^T as an entry point ----------------------------------------------------------
// For the SVGP family var_i = k_ii - k_i^T A^-1 k_i + k_i^T Sigma k_i with frozen kernel hyper-parameters, so
// dLoss/dSigma = sum_i g_i k_i k_i^T: the weighted-Gram (SYRK) kernel on the tensor pipe IS the vector-Jacobian product
// (cholesky_svgp_kernel.py:133-150, diagonal_svgp_kernel.py:117-129, log_svgp_kernel.py:108-124); each
// parameterisation of Sigma is an M x M map the host framework differentiates.  C_out is [M,M] row-major.
namespace {
__global__ void copy_block_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int M) {
  const int64_t total = (int64_t)M * M;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    dst[idx] = src[(idx / M) * lds + idx % M];
}
}  // namespace
extern "C" int gvi_weighted_gram_f64(const void* factor, int M, int D, const double* X, int64_t N, const double* g,
                                     double* C_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(M >= 1 && D >= 1 && N >= 1, "sizes");
  GVI_CHECK_ARG(D <= 32, "the weighted Gram supports D <= 32 in this build");
  GVI_CHECK_ARG(factor && X && g && C_out && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const FactorLayout f = factor_layout(M, D);
  GradWs w = carve_grad(workspace, N, M, D, nullptr);
  int rc = gvi_weighted_gram((const double*)factor, X, g, N, M, D, w, st);
  if (rc) return rc;
  copy_block_kernel<<<GVI_NUM_SMS * 2, 256, 0, st>>>(w.C, f.Mp, C_out, M);
  GVI_CHECK_LAUNCH("copy_block_kernel");
  return GVI_OK;
}

// ---- SVGP family: value and gradient w.r.t. the extra matrix E (the kernel hyper-parameters are frozen) -------------
// v_i = k_ii - ||L^-1 k_i||^2 + ||E k_i||^2  =>  dR/dE = 2 E C (lower triangle), C = sum_i g_i k_i k_i^T.
// out[0..2] as out3; dm_out[N]; dE_out[M*M] row-major (only the lower triangle is meaningful).  E is the dense matrix
// that was attached with gvi_gp_factor_set_extra_f64.
namespace {
__global__ void tril_scale_kernel(double* __restrict__ A, int M, int64_t ld, double scale) {
  const int64_t total = (int64_t)M * M;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(idx / M), c = (int)(idx % M);
    A[(int64_t)r * ld + c] = (c <= r) ? scale * A[(int64_t)r * ld + c] : 0.0;
  }
}
}  // namespace

extern "C" int gvi_svgp_step_grad_f64(const void* factor_q, const void* factor_p, int M, int D, const double* X,
                                      const double* y, const double* m_q, const double* prior_mean_p,
                                      const double* m_p_in, const double* v_p_in, int64_t N, int reg_kind,
                                      double renyi_alpha, const double* E, int64_t lde, double* out3, double* dm_out,
                                      double* dE_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(out3 && dm_out && dE_out && E && workspace, "null pointer");
  GVI_CHECK_ARG(lde >= M, "lde < M");
  cudaStream_t st = (cudaStream_t)stream;
  StepParams P{};
  int rc = gvi_fill_step_params(P, factor_q, factor_p, M, D, X, y, m_q, prior_mean_p, m_p_in, v_p_in, N, reg_kind,
                                renyi_alpha, nullptr, nullptr, nullptr);
  if (rc) return rc;
  GradWs w = carve_grad(workspace, N, M, D, nullptr);
  const FactorLayout f = factor_layout(M, D);
  const double* Fq = (const double*)factor_q;
  P.partials = w.partials;
  P.tile_counter = w.tile_counter;
  P.tile_partials = w.tile_partials;
  P.grad = 1;
  P.grad_pointwise_only = 1;
  P.g_out = w.g;
  P.dm_out = dm_out;
  P.bscratch = w.bscratch;
  P.acc_scratch = w.pscratch;  // panel scratch of the two-CTA gradient shape (the p launch of a chunk is done with it)
  int grid = gvi_launch_step(P, st);
  if (grid < 0) return grid;
  if ((grid = gvi_launch_tile_reduce(w.tile_partials, w.ntiles, 1, w.partials, st)) < 0) return grid;
  if ((rc = gvi_launch_finish(P.partials, grid, N, 0, D, Fq, out3, st))) return rc;
  if ((rc = gvi_weighted_gram(Fq, X, w.g, N, M, D, w, st))) return rc;
  // dE = 2 tril(E C): the in-tree tensor-pipe GEMM (dgemm.cu) on the M x M blocks
  if ((rc = gvi_dgemm_rm(false, false, M, M, M, E, lde, w.C, f.Mp, 0.0, dE_out, M, GVI_GEMM_K_FULL, false, st))) return rc;
  tril_scale_kernel<<<GVI_NUM_SMS * 2, 256, 0, st>>>(dE_out, M, M, 2.0);
  GVI_CHECK_LAUNCH("tril_scale_kernel");
  return GVI_OK;
}
