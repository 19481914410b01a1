// chol_coop.cu - the M x M factorisation chain of one GP as ONE cooperative launch: A = L L^T (right-looking, 64 x 64
// blocks, the next diagonal block factorised one step ahead, behind the trailing update) and X = L^-1 (block
// sub-diagonals, split-K partial tiles summed in a fixed order).  Replaces the per-panel launch chain of chol.cu
// (33 dependent launches at M = 1024, every panel CTA re-reading its rows from L2 fragment by fragment) for the
// cho_factor call sites of the reference (sparse_posterior_kernel.py:68-74, src/gps/base/base.py:516-518).
//
// All block products (64 x 64 x 64) run on the FP64 tensor pipe (DMMA m8n8k4) from shared-memory tiles; a grid-wide
// barrier separates dependent phases (2 per block step of the factorisation, 2 per sub-diagonal of the inverse).  The
// within-block arithmetic of the diagonal factorisation (column-by-column right-looking elimination, sqrt then divide)
// and of the diagonal-block inverse is the one chol.cu has always used, so every result that depends on a single 32 x 32
// block only (the reference's 2-point goldens) is bit-identical to the previous build.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {
constexpr int CB = 64;   // block size
constexpr int LS = 68;   // shared-memory row stride (doubles): 8-byte fragment loads of a half-warp hit 32 distinct banks
constexpr int HB = 32;   // half block: the diagonal factorisation works on 2 x 2 sub-blocks of 32
typedef double Tile[CB][LS];

// ---- 64 x 64 tile movement (256 threads) -----------------------------------------------------------------------
__device__ __forceinline__ void load_tile(Tile& dst, const double* __restrict__ src, int64_t ld) {
  for (int idx = threadIdx.x; idx < CB * CB / 2; idx += 256) {
    const int r = idx >> 5, c = (idx & 31) * 2;
    *reinterpret_cast<double2*>(&dst[r][c]) = *reinterpret_cast<const double2*>(src + (int64_t)r * ld + c);
  }
}
__device__ __forceinline__ void store_tile(double* __restrict__ dst, int64_t ld, const Tile& src) {
  for (int idx = threadIdx.x; idx < CB * CB / 2; idx += 256) {
    const int r = idx >> 5, c = (idx & 31) * 2;
    *reinterpret_cast<double2*>(dst + (int64_t)r * ld + c) = *reinterpret_cast<const double2*>(&src[r][c]);
  }
}

// ---- 64 x 64 x 64 product on the tensor pipe: 8 warps, warp (wm, wn) owns rows [16 wm, +16) x columns [32 wn, +32) ---
// TB = true : C += As Bs^T (Bs[n][k]);  TB = false: C += As Bs (Bs[k][n]).   acc[mt][nt][e] = C[16wm + 8mt + lane/4][32wn + 8nt + 2(lane%4) + e]
template <bool TB>
__device__ __forceinline__ void tile_mma(const Tile& As, const Tile& Bs, double (&acc)[2][4][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int fr = lane >> 2, fk = lane & 3;
#pragma unroll 4
  for (int k0 = 0; k0 < CB; k0 += 4) {
    double a[2], b[4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++) a[mt] = As[16 * wm + 8 * mt + fr][k0 + fk];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) b[nt] = TB ? Bs[32 * wn + 8 * nt + fr][k0 + fk] : Bs[k0 + fk][32 * wn + 8 * nt + fr];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}
__device__ __forceinline__ void acc_zero(double (&acc)[2][4][2]) {
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
}
// fragment-wise visit of the warp's part of a 64 x 64 tile: f(row, col, value pair)
template <typename F>
__device__ __forceinline__ void acc_visit(double (&acc)[2][4][2], F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp & 3, wn = warp >> 2;
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) f(16 * wm + 8 * mt + (lane >> 2), 32 * wn + 8 * nt + 2 * (lane & 3), acc[mt][nt]);
}

// ---- diagonal block: L = chol(T) and its inverse --------------------------------------------------------------
// 32 x 32 Cholesky of the sub-block at (o, o) of T, one warp, lane r owns row r - the arithmetic of chol.cu's panel kernel
__device__ __forceinline__ void chol32(Tile& T, int o, int lane, int col0, int* __restrict__ info) {
  double row[HB];
#pragma unroll
  for (int c = 0; c < HB; c++) row[c] = T[o + lane][o + c];
#pragma unroll
  for (int j = 0; j < HB; j++) {
    const double djj = __shfl_sync(0xffffffffu, row[j], j);
    if (lane == j && !(djj > 0.0) && info) atomicMin(info, col0 + j + 1);
    const double sq = sqrt(djj);
    row[j] = (lane == j) ? sq : row[j] / sq;
    T[o + lane][o + j] = row[j];  // column j of L (valid for lanes >= j)
    __syncwarp();
#pragma unroll
    for (int k = j + 1; k < HB; k++) {
      const double lkj = T[o + k][o + j];  // broadcast
      if (lane >= k) row[k] = fma(-row[j], lkj, row[k]);
    }
  }
#pragma unroll
  for (int c = 0; c < HB; c++) T[o + lane][o + c] = (c <= lane) ? row[c] : 0.0;
  __syncwarp();
}
// inverse of the lower-triangular 32 x 32 sub-block at (o, o) of T into Dv (same position): lane j owns column j,
// x[i] = (delta_ij - sum_{k=j}^{i-1} L[i][k] x[k]) / L[i][i] - the arithmetic of chol.cu's diag_inverse_kernel
__device__ __forceinline__ void inv32(const Tile& T, Tile& Dv, int o, int lane) {
  double x[HB];
#pragma unroll
  for (int i = 0; i < HB; i++) {
    double v = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; k++) v = fma(-T[o + i][o + k], (k >= lane) ? x[k] : 0.0, v);
    x[i] = (i >= lane) ? v / T[o + i][o + i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < HB; i++) Dv[o + i][o + lane] = x[i];
}

// T (64 x 64, symmetric input, lower part used) -> L in T (upper part zeroed), L^-1 in Dv; W is scratch.  256 threads.
__device__ void diag_factor(Tile& T, Tile& Dv, Tile& W, int col0, int* __restrict__ info) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  if (warp == 0) chol32(T, 0, lane, col0, info);
  __syncthreads();
  // L21 = A21 L11^-T (warp 0, lane owns a row: forward substitution) while warp 1 inverts L11
  if (warp == 0) {
    double x[HB];
#pragma unroll
    for (int c = 0; c < HB; c++) x[c] = T[HB + lane][c];
#pragma unroll
    for (int c = 0; c < HB; c++) {
      double v = x[c];
#pragma unroll
      for (int k = 0; k < c; k++) v = fma(-x[k], T[c][k], v);
      x[c] = v / T[c][c];
    }
#pragma unroll
    for (int c = 0; c < HB; c++) T[HB + lane][c] = x[c];
  } else if (warp == 1) {
    inv32(T, Dv, 0, lane);
  }
  __syncthreads();
  // A22 -= L21 L21^T (lower part), 4 entries per thread, k ascending
  for (int idx = tid; idx < HB * HB; idx += 256) {
    const int r = idx >> 5, c = idx & 31;
    if (c <= r) {
      double v = T[HB + r][HB + c];
#pragma unroll 8
      for (int k = 0; k < HB; k++) v = fma(-T[HB + r][k], T[HB + c][k], v);
      T[HB + r][HB + c] = v;
    }
  }
  __syncthreads();
  if (warp == 0) chol32(T, HB, lane, col0 + HB, info);
  __syncthreads();
  if (warp == 0) inv32(T, Dv, HB, lane);
  // zero blocks: upper right of L and of L^-1
  for (int idx = tid; idx < HB * HB; idx += 256) {
    const int r = idx >> 5, c = idx & 31;
    T[r][HB + c] = 0.0;
    Dv[r][HB + c] = 0.0;
  }
  __syncthreads();
  // Dv21 = -D2 (L21 D1): W = L21 D1, then Dv21 = -D2 W
  for (int idx = tid; idx < HB * HB; idx += 256) {
    const int r = idx >> 5, c = idx & 31;
    double v = 0.0;
#pragma unroll 8
    for (int k = 0; k < HB; k++) v = fma(T[HB + r][k], Dv[k][c], v);
    W[r][c] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < HB * HB; idx += 256) {
    const int r = idx >> 5, c = idx & 31;
    double v = 0.0;
#pragma unroll 8
    for (int k = 0; k < HB; k++) v = fma(Dv[HB + r][HB + k], W[k][c], v);
    Dv[HB + r][c] = -v;
  }
  __syncthreads();
}

struct CoopArgs {
  double* A;    // [Mw, Mw] in: symmetric (lower part used), out: L (strict upper part of the diagonal blocks zeroed)
  double* X;    // [Mw, Mw] out: L^-1 (upper block triangle zeroed)
  double* P;    // split-K partial tiles of the inverse: (G/2)^2 x 64 x 64 doubles
  int Mw, G;
  int* info;
};

__global__ void __launch_bounds__(256, 1) chol_inverse_coop_kernel(const CoopArgs a) {
  extern __shared__ __align__(16) unsigned char coop_smem[];
  Tile& As = *reinterpret_cast<Tile*>(coop_smem);
  Tile& Bs = *reinterpret_cast<Tile*>(coop_smem + sizeof(Tile));
  Tile& Cs = *reinterpret_cast<Tile*>(coop_smem + 2 * sizeof(Tile));
  cg::grid_group grid = cg::this_grid();
  const int G = a.G, nb = gridDim.x, b = blockIdx.x;
  const int64_t ld = a.Mw;
  double* const A = a.A;
  double* const X = a.X;
  auto blk = [&](double* base, int i, int j) { return base + ((int64_t)i * CB) * ld + (int64_t)j * CB; };

  // upper block triangle of X := 0 (the diagonal and lower blocks are all written below)
  for (int t = b; t < G * G; t += nb) {
    const int i = t / G, j = t % G;
    if (j > i)
      for (int idx = threadIdx.x; idx < CB * CB; idx += 256) blk(X, i, j)[(int64_t)(idx >> 6) * ld + (idx & 63)] = 0.0;
  }
  if (b == 0) {  // step 0: the first diagonal block
    load_tile(Cs, blk(A, 0, 0), ld);
    diag_factor(Cs, As, Bs, 0, a.info);
    store_tile(blk(A, 0, 0), ld, Cs);
    store_tile(blk(X, 0, 0), ld, As);
  }
  grid.sync();
  for (int k = 0; k < G - 1; k++) {
    // ---- panel: L[i,k] = A[i,k] Dinv_k^T, i > k --------------------------------------------------------------
    bool have_d = false;
    for (int i = k + 1 + b; i < G; i += nb) {
      __syncthreads();
      if (!have_d) {
        load_tile(Bs, blk(X, k, k), ld);
        have_d = true;
      }
      load_tile(As, blk(A, i, k), ld);
      __syncthreads();
      double acc[2][4][2];
      acc_zero(acc);
      tile_mma<true>(As, Bs, acc);
      double* dst = blk(A, i, k);
      acc_visit(acc, [&](int r, int c, double (&v)[2]) {
        *reinterpret_cast<double2*>(dst + (int64_t)r * ld + c) = make_double2(v[0], v[1]);
      });
    }
    grid.sync();
    // ---- trailing update A[i,j] -= L[i,k] L[j,k]^T for k < j <= i; task 0 = (k+1, k+1) goes on to factorise it ----
    const int n = G - k - 1;             // blocks per side of the trailing matrix
    const int ntask = n * (n + 1) / 2;
    // CTA 0 takes task 0 only (and then the diagonal factorisation); the other tasks go round the remaining CTAs
    for (int t = (nb == 1 ? 0 : (b == 0 ? 0 : b)); t < ntask; t += (nb == 1 ? 1 : (b == 0 ? ntask : nb - 1))) {
      // t -> (i, j), row-major over the lower triangle: t = i'(i'+1)/2 + j'
      int ip = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while ((ip + 1) * (ip + 2) / 2 <= t) ip++;
      while (ip * (ip + 1) / 2 > t) ip--;
      const int jp = t - ip * (ip + 1) / 2;
      const int i = k + 1 + ip, j = k + 1 + jp;
      __syncthreads();
      load_tile(As, blk(A, i, k), ld);
      load_tile(Bs, blk(A, j, k), ld);
      __syncthreads();
      double acc[2][4][2];
      acc_zero(acc);
      tile_mma<true>(As, Bs, acc);
      double* dst = blk(A, i, j);
      const bool diag_next = (t == 0);
      acc_visit(acc, [&](int r, int c, double (&v)[2]) {
        double2 old = *reinterpret_cast<const double2*>(dst + (int64_t)r * ld + c);
        old.x -= v[0];
        old.y -= v[1];
        if (diag_next) *reinterpret_cast<double2*>(&Cs[r][c]) = old;
        else *reinterpret_cast<double2*>(dst + (int64_t)r * ld + c) = old;
      });
      if (diag_next) {  // look-ahead: this CTA factorises block (k+1, k+1) while the others finish the update
        diag_factor(Cs, As, Bs, (k + 1) * CB, a.info);
        store_tile(blk(A, k + 1, k + 1), ld, Cs);
        store_tile(blk(X, k + 1, k + 1), ld, As);
      }
    }
    grid.sync();
  }
  // ---- X = L^-1 by block sub-diagonals d = i - j: X[i,j] = -Dinv_i sum_{k=j}^{i-1} L[i,k] X[k,j] --------------------
  for (int d = 1; d < G; d++) {
    const int nblk = G - d;        // blocks on this sub-diagonal
    const int ntask = nblk * d;    // one product per task: (j, kk), k = j + kk
    for (int t = b; t < ntask; t += nb) {
      const int j = t / d, kk = t % d;
      const int i = j + d, kq = j + kk;
      __syncthreads();
      load_tile(As, blk(A, i, kq), ld);
      load_tile(Bs, blk(X, kq, j), ld);
      __syncthreads();
      double acc[2][4][2];
      acc_zero(acc);
      tile_mma<false>(As, Bs, acc);
      double* dst = a.P + (int64_t)t * CB * CB;
      acc_visit(acc, [&](int r, int c, double (&v)[2]) {
        *reinterpret_cast<double2*>(dst + r * CB + c) = make_double2(v[0], v[1]);
      });
    }
    grid.sync();
    for (int j = b; j < nblk; j += nb) {
      const int i = j + d;
      __syncthreads();
      // Bs = sum of the d partial tiles in task order (fixed: the result does not depend on which CTA produced which)
      for (int idx = threadIdx.x; idx < CB * CB / 2; idx += 256) {
        const int r = idx >> 5, c = (idx & 31) * 2;
        const double* p = a.P + (int64_t)(j * d) * CB * CB + r * CB + c;
        double2 s = *reinterpret_cast<const double2*>(p);
        for (int kk = 1; kk < d; kk++) {
          const double2 q = *reinterpret_cast<const double2*>(p + (int64_t)kk * CB * CB);
          s.x += q.x;
          s.y += q.y;
        }
        *reinterpret_cast<double2*>(&Bs[r][c]) = s;
      }
      load_tile(As, blk(X, i, i), ld);  // Dinv_i
      __syncthreads();
      double acc[2][4][2];
      acc_zero(acc);
      tile_mma<false>(As, Bs, acc);
      double* dst = blk(X, i, j);
      acc_visit(acc, [&](int r, int c, double (&v)[2]) {
        *reinterpret_cast<double2*>(dst + (int64_t)r * ld + c) = make_double2(-v[0], -v[1]);
      });
    }
    grid.sync();
  }
}
}  // namespace

size_t gvi_chol_coop_scratch_doubles(int Mw) {
  const int64_t G = Mw / CB;
  const int64_t tasks = ((G + 1) / 2) * (G / 2) + 1;  // max over d of (G - d) d
  return (size_t)tasks * CB * CB;
}

// A, X: [Mw, Mw] row-major, Mw a multiple of 64 (identity-padded by the caller); *info must hold INT_MAX on entry
// (atomicMin of 1 + the first non-positive pivot column).  One cooperative launch on `st`.
int gvi_launch_chol_inverse_coop(double* A, double* X, double* P, int Mw, int* info, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};
  const size_t smem = 3 * sizeof(Tile);
  if (gvi_first_use_on_this_device(configured))
    GVI_CUDA(cudaFuncSetAttribute(chol_inverse_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CoopArgs args{A, X, P, Mw, Mw / CB, info};
  const int G = Mw / CB;
  int want = G * (G - 1) / 2;  // widest phase: the first trailing update
  const int quarter = ((G + 1) / 2) * (G / 2);
  if (quarter > want) want = quarter;
  int grid = want < 1 ? 1 : (want < GVI_NUM_SMS ? want : GVI_NUM_SMS);
  int dev = 0, sms = 0, coop = 0;
  GVI_CUDA(cudaGetDevice(&dev));
  GVI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GVI_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) {
    gvi_set_error("the device does not support cooperative launches");
    return GVI_EUNSUPPORTED;
  }
  if (grid > sms) grid = sms;
  void* params[] = {(void*)&args};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)chol_inverse_coop_kernel, dim3(grid), dim3(256), params, smem, st);
  gvi_count_launch();
  if (e != cudaSuccess) {
    gvi_set_error("CUDA error launching chol_inverse_coop_kernel: %s", cudaGetErrorString(e));
    return GVI_ECUDA;
  }
  return GVI_OK;
}
