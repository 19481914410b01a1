// dgemm.cu - the library's own FP64 GEMM on the tensor pipe (DMMA m8n8k4), for the M^3- and N M^2-shaped products of the
// gradient routes that are not the fused tile kernel: the exact-GP NLL step (exact.cu; reference:
// experiments/regression/trainers.py:52-80 through src/gps/base/base.py:516-525), the variance VJP for wide inputs
// (grad_wide.cu; src/gps/base/classification_base.py:76-153 at D = 784) and dLoss/dE of the SVGP family (grad.cu).
// Row-major C[m,n] = op(A) op(B) + beta C.  What a library GEMM cannot know and this one uses:
//   - the triangular operand: L^-1 is lower triangular, so K L^-T only needs k <= n and T L^-1 only k >= n - the k loop
//     of an output tile stops (starts) at the tile's column range: half the flops of both products;
//   - the symmetric result: S = Aa^T diag(g) Aa is symmetric - only tiles on or below the diagonal are computed, their
//     mirror images are written in the same epilogue.
// 128 x 128 output tile per CTA, 16 warps (4 x 4, warp tile 32 x 32 = 4 x 4 DMMA tiles), k in chunks of 16 through a
// double-buffered shared-memory stage (register-staged global loads of chunk c+1 under the DMMAs of chunk c).  Operands
// sit in shared memory in the orientation they have in global memory ([row][k] or [k][row], padded so that the 8-byte
// fragment loads of a half-warp touch 32 distinct banks); no transposition pass.
#include "common.cuh"
#include "dgemm.cuh"

namespace {
constexpr int BK = 16, GT = 512;
// WT = edge of a warp's output tile: 32 -> 128 x 128 per CTA (the throughput shape), 16 -> 64 x 64 per CTA, two CTAs per
// SM (small products: M^3-shaped work at M ~ 1024 has only 64 tiles of 128 x 128 - 43 % of the SMs, each bound by its own
// tensor pipe - against 256 tiles of 64 x 64)
template <int WT>
struct Tile {
  static constexpr int BM = 4 * WT, BN = 4 * WT, NF = WT / 8;  // NF x NF DMMA tiles per warp
  static constexpr int PER = BM * BK / GT;                     // staged elements per thread and operand
  static constexpr int LDK = BK + 4;                           // row stride of a [row][k] stage
  static constexpr int LDR = BM + 4;                           // row stride of a [k][row] stage (BM == BN)
  static constexpr int STAGE_A = BM * LDK > BK * LDR ? BM * LDK : BK * LDR;  // doubles per operand stage, either orientation
  static constexpr int STAGE = 2 * STAGE_A;                                   // A + B
};

struct GemmParams {
  int m, n, k;
  const double* A;
  int64_t lda;
  const double* B;
  int64_t ldb;
  double beta;
  double* C;
  int64_t ldc;
  int kmode;  // GVI_GEMM_K_*
  int symm;   // compute tiles with m0 + BM > n0 only and mirror them (m == n, result symmetric)
};

// element (r, kk) of op(X) for a tile starting at (r0, k0): TX = false -> X[r][k] (k contiguous), true -> X[k][r]
template <int WT, bool TX>
__device__ __forceinline__ void load_stage_regs(const double* __restrict__ X, int64_t ldx, int r0, int rmax, int k0,
                                                int kmax, double (&v)[Tile<WT>::PER]) {
  constexpr int BM = Tile<WT>::BM;
  const int t = threadIdx.x;
#pragma unroll
  for (int e = 0; e < Tile<WT>::PER; e++) {
    const int idx = t + GT * e;
    int r, kk;
    if (!TX) {
      r = idx / BK;
      kk = idx % BK;
    } else {
      kk = idx / BM;
      r = idx % BM;
    }
    const int gr = r0 + r, gk = k0 + kk;
    double val = 0.0;
    if (gr < rmax && gk < kmax) val = TX ? X[(int64_t)gk * ldx + gr] : X[(int64_t)gr * ldx + gk];
    v[e] = val;
  }
}
template <int WT, bool TX>
__device__ __forceinline__ void store_stage(double* __restrict__ S, const double (&v)[Tile<WT>::PER]) {
  constexpr int BM = Tile<WT>::BM, LDK = Tile<WT>::LDK, LDR = Tile<WT>::LDR;
  const int t = threadIdx.x;
#pragma unroll
  for (int e = 0; e < Tile<WT>::PER; e++) {
    const int idx = t + GT * e;
    if (!TX) S[(idx / BK) * LDK + idx % BK] = v[e];
    else S[(idx / BM) * LDR + idx % BM] = v[e];
  }
}
// fragment element op(X)[row][kk] from a stage
template <int WT, bool TX>
__device__ __forceinline__ double frag(const double* __restrict__ S, int row, int kk) {
  return TX ? S[kk * Tile<WT>::LDR + row] : S[row * Tile<WT>::LDK + kk];
}

// TA: op(A) = A^T (A stored [k][m]);  TB: op(B) = B^T (B stored [n][k])
template <int WT, bool TA, bool TB>
__global__ void __launch_bounds__(GT, WT == 32 ? 1 : 2) dgemm_kernel(const GemmParams p) {
  using TL = Tile<WT>;
  constexpr int BM = TL::BM, BN = TL::BN, NF = TL::NF, PER = TL::PER, STAGE = TL::STAGE, STAGE_A = TL::STAGE_A;
  extern __shared__ __align__(16) double gemm_smem[];
  // tiles with the longest k range first (K_UPTO_N: the rightmost column tiles), so the tail of the grid is cheap
  const int bx = p.kmode == GVI_GEMM_K_UPTO_N ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int m0 = blockIdx.y * BM, n0 = bx * BN;
  if (p.symm && m0 + BM <= n0) return;  // strictly above the diagonal: written as the mirror of its transpose
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3;  // 4 x 4 warps, warp tile WT x WT
  const int fr = lane >> 2, fk = lane & 3;
  // k range of this output tile
  int kbeg = 0, kend = p.k;
  if (p.kmode == GVI_GEMM_K_UPTO_N) kend = min(p.k, n0 + BN);          // op(B)[k][n] = 0 for k > n
  else if (p.kmode == GVI_GEMM_K_FROM_N) kbeg = (n0 / BK) * BK;        // op(B)[k][n] = 0 for k < n
  double acc[NF][NF][2];
#pragma unroll
  for (int mt = 0; mt < NF; mt++)
#pragma unroll
    for (int nt = 0; nt < NF; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  // op(A) stage: TA means stored [k][m] -> "TX" orientation; op(B): stored [n][k] when TB (the [row][k] orientation)
  constexpr bool AX = TA, BX = !TB;
  double ra[PER], rb[PER];
  const int nchunk = (kend - kbeg + BK - 1) / BK;
  if (nchunk > 0) {
    load_stage_regs<WT, AX>(p.A, p.lda, m0, p.m, kbeg, kend, ra);
    load_stage_regs<WT, BX>(p.B, p.ldb, n0, p.n, kbeg, kend, rb);
  }
  for (int c = 0; c < nchunk; c++) {
    double* Sa = gemm_smem + (c & 1) * STAGE;
    double* Sb = Sa + STAGE_A;
    store_stage<WT, AX>(Sa, ra);
    store_stage<WT, BX>(Sb, rb);
    __syncthreads();
    if (c + 1 < nchunk) {  // the next chunk travels while this one is multiplied
      load_stage_regs<WT, AX>(p.A, p.lda, m0, p.m, kbeg + (c + 1) * BK, kend, ra);
      load_stage_regs<WT, BX>(p.B, p.ldb, n0, p.n, kbeg + (c + 1) * BK, kend, rb);
    }
#pragma unroll
    for (int k0 = 0; k0 < BK; k0 += 4) {
      double a[NF], b[NF];
#pragma unroll
      for (int mt = 0; mt < NF; mt++) a[mt] = frag<WT, AX>(Sa, WT * wm + 8 * mt + fr, k0 + fk);
#pragma unroll
      for (int nt = 0; nt < NF; nt++) b[nt] = frag<WT, BX>(Sb, WT * wn + 8 * nt + fr, k0 + fk);
#pragma unroll
      for (int mt = 0; mt < NF; mt++)
#pragma unroll
        for (int nt = 0; nt < NF; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
    // the stage written two iterations from now is this one: one barrier per chunk suffices with two buffers, because
    // the next store_stage goes to the other buffer and the one after that is separated from these reads by that barrier
  }
  // epilogue
#pragma unroll
  for (int mt = 0; mt < NF; mt++)
#pragma unroll
    for (int nt = 0; nt < NF; nt++) {
      const int r = m0 + WT * wm + 8 * mt + fr, c = n0 + WT * wn + 8 * nt + 2 * fk;
      if (r >= p.m) continue;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        if (c + e >= p.n) continue;
        double v = acc[mt][nt][e];
        double* dst = p.C + (int64_t)r * p.ldc + c + e;
        if (p.beta != 0.0) v = fma(p.beta, *dst, v);
        *dst = v;
        if (p.symm && m0 != n0) {
          // mirror image (tiles below the diagonal; inside a diagonal tile both halves were computed directly)
          double* mir = p.C + (int64_t)(c + e) * p.ldc + r;
          double w = acc[mt][nt][e];
          if (p.beta != 0.0) w = fma(p.beta, *mir, w);
          *mir = w;
        }
      }
    }
}

template <int WT, bool TA, bool TB>
int launch_wt(const GemmParams& p, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};
  constexpr int BM = Tile<WT>::BM, BN = Tile<WT>::BN;
  const size_t smem = (size_t)2 * Tile<WT>::STAGE * sizeof(double);
  if (gvi_first_use_on_this_device(configured))
    GVI_CUDA(cudaFuncSetAttribute(dgemm_kernel<WT, TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((p.n + BN - 1) / BN), (unsigned)((p.m + BM - 1) / BM));
  dgemm_kernel<WT, TA, TB><<<grid, GT, smem, st>>>(p);
  GVI_CHECK_LAUNCH("dgemm_kernel");
  return GVI_OK;
}
template <bool TA, bool TB>
int launch(const GemmParams& p, cudaStream_t st) {
  // tiles of 128 x 128 that do work (a symmetric result computes the lower triangle only): below ~3/4 of a wave the
  // product is bound by the tensor pipes of the few SMs that hold a tile -> quarter-size tiles, two CTAs per SM
  const int64_t tn = (p.n + 127) / 128, tm = (p.m + 127) / 128;
  const int64_t tiles = p.symm ? tn * (tn + 1) / 2 : tn * tm;
  if (tiles * 4 < (int64_t)GVI_NUM_SMS * 3) return launch_wt<16, TA, TB>(p, st);
  return launch_wt<32, TA, TB>(p, st);
}
}  // namespace

int gvi_dgemm_rm(bool ta, bool tb, int m, int n, int k, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double beta, double* C, int64_t ldc, int kmode, bool symmetric, cudaStream_t st) {
  if (m <= 0 || n <= 0) return GVI_OK;
  GVI_CHECK_ARG(A && B && C && k >= 0, "gemm arguments");
  GVI_CHECK_ARG(!symmetric || m == n, "a symmetric result is square");
  GemmParams p{m, n, k, A, lda, B, ldb, beta, C, ldc, kmode, symmetric ? 1 : 0};
  if (!ta && !tb) return launch<false, false>(p, st);
  if (!ta && tb) return launch<false, true>(p, st);
  if (ta && !tb) return launch<true, false>(p, st);
  return launch<true, true>(p, st);
}

namespace {
// y[m] (+)= op(A) x for row-major A: TA = false: A is [m][k]; TA = true: A is [k][m] (y = A^T x).  One CTA per 32 outputs.
template <bool TA>
__global__ void __launch_bounds__(256) dgemv_kernel(int m, int k, const double* __restrict__ A, int64_t lda,
                                                    const double* __restrict__ x, double beta, double* __restrict__ y) {
  __shared__ double part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.x * 32;
  double s = 0.0;
  if (TA) {  // coalesced over the outputs (lanes), k split over the warps
    const int r = r0 + lane;
    if (r < m)
      for (int kk = warp; kk < k; kk += 8) s = fma(A[(int64_t)kk * lda + r], x[kk], s);
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && r < m) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < 8; w++) v += part[w][lane];
      y[r] = beta != 0.0 ? fma(beta, y[r], v) : v;
    }
  } else {  // one warp per output row (4 rows each), lanes over k
    for (int q = 0; q < 4; q++) {
      const int r = r0 + 4 * warp + q;
      if (r >= m) break;
      double v = 0.0;
      for (int kk = lane; kk < k; kk += 32) v = fma(A[(int64_t)r * lda + kk], x[kk], v);
      v = warp_sum(v);
      if (lane == 0) y[r] = beta != 0.0 ? fma(beta, y[r], v) : v;
    }
  }
}
}  // namespace

int gvi_dgemv_rm(bool ta, int m, int k, const double* A, int64_t lda, const double* x, double beta, double* y,
                 cudaStream_t st) {
  if (m <= 0) return GVI_OK;
  GVI_CHECK_ARG(A && x && y && k >= 0, "gemv arguments");
  if (ta) dgemv_kernel<true><<<(m + 31) / 32, 256, 0, st>>>(m, k, A, lda, x, beta, y);
  else dgemv_kernel<false><<<(m + 31) / 32, 256, 0, st>>>(m, k, A, lda, x, beta, y);
  GVI_CHECK_LAUNCH("dgemv_kernel");
  return GVI_OK;
}

#ifdef GVI_TEST_HOOKS  // libgvigp_test.so only (include/gvigp_test.h): the GEMM / GEMV on their own, for parity and timing
extern "C" int gvi_test_dgemm_f64(int ta, int tb, int m, int n, int k, const double* A, int64_t lda, const double* B,
                                  int64_t ldb, double beta, double* C, int64_t ldc, int kmode, int symmetric,
                                  void* stream) {
  return gvi_dgemm_rm(ta != 0, tb != 0, m, n, k, A, lda, B, ldb, beta, C, ldc, kmode, symmetric != 0, (cudaStream_t)stream);
}
extern "C" int gvi_test_dgemv_f64(int ta, int m, int k, const double* A, int64_t lda, const double* x, double beta,
                                  double* y, void* stream) {
  return gvi_dgemv_rm(ta != 0, m, k, A, lda, x, beta, y, (cudaStream_t)stream);
}
#endif
