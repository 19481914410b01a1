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

#ifdef GVI_COOP_STAMPS  // lab builds (tools/lab/chol_lab.cu): CTA 0 records %globaltimer at phase boundaries
__device__ long long* g_lab_stamps = nullptr;
#define STAMP(id)                                                                          \
  do {                                                                                     \
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_lab_stamps) {                             \
      long long t__;                                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                              \
      g_lab_stamps[2 * g_lab_stamps[0] + 2] = (id);                                        \
      g_lab_stamps[2 * g_lab_stamps[0] + 3] = t__;                                         \
      g_lab_stamps[0]++;                                                                   \
    }                                                                                      \
  } while (0)
#else
#define STAMP(id)
#endif

// ---- diagonal block: L = chol(T) and its inverse --------------------------------------------------------------
// 32 x 32 Cholesky of the sub-block at (o, o) of T, one warp, lane r owns row r (column-by-column right-looking).
// FAST = false: pivot sq = sqrt(d), column = row / sq - IEEE square root and division, the arithmetic chol.cu's panel kernel
//   has always used (bit-identical results for matrices that fit one 32 x 32 block, e.g. the reference's 2-point goldens).
// FAST = true : r = rsqrt(d), pivot = d r, column = row r.  A factorisation is a serial chain of M pivots; an IEEE
//   sqrt followed by a divide is ~60 dependent FP64 instructions (~800 cycles for a lone warp), i.e. 0.4 ms at M = 1024
//   before any other work - the reciprocal square root (MUFU seed + Newton steps) and a multiply are ~10.
template <bool FAST>
__device__ __forceinline__ void chol32(Tile& T, int o, Tile& Wscratch, int col0, int* __restrict__ info) {
  // Works on a [32][33] copy (stride 33: a warp's accesses to its own rows are bank-conflict free) in panels of 8 columns:
  // warp 0 eliminates a panel in registers (lane = row, pivots and multipliers travel by shuffle), then all 256 threads
  // apply the panel to the trailing sub-block.  Every element still receives its fma updates in ascending column order
  // followed by the pivot scaling, i.e. the arithmetic of the plain column-by-column elimination; but the serial chain is
  // one shuffle + pivot + fma per column instead of a warp-wide sweep over shared memory (a fully unrolled 32-column
  // version was instruction-fetch bound, a compact shared-memory loop latency bound: 17-18 us per call either way).
  double(*D)[HB + 1] = reinterpret_cast<double(*)[HB + 1]>(&Wscratch[0][0]);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < HB * HB; idx += 256) D[idx >> 5][idx & 31] = T[o + (idx >> 5)][o + (idx & 31)];
  __syncthreads();
#pragma unroll 1
  for (int jb = 0; jb < HB; jb += 8) {
    if (warp == 0) {
      double p[8];
#pragma unroll
      for (int jj = 0; jj < 8; jj++) p[jj] = D[lane][jb + jj];
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        const int j = jb + jj;
        const double djj = __shfl_sync(0xffffffffu, p[jj], j);
        if (lane == j && !(djj > 0.0) && info) atomicMin(info, col0 + j + 1);
        if (FAST) {
          const double r = rsqrt(djj);
          p[jj] = (lane == j) ? djj * r : p[jj] * r;
        } else {
          const double sq = sqrt(djj);
          p[jj] = (lane == j) ? sq : p[jj] / sq;
        }
#pragma unroll
        for (int kk = jj + 1; kk < 8; kk++) {
          const double lkj = __shfl_sync(0xffffffffu, p[jj], jb + kk);  // L[jb + kk][j]
          if (lane >= jb + kk) p[kk] = fma(-p[jj], lkj, p[kk]);
        }
      }
#pragma unroll
      for (int jj = 0; jj < 8; jj++)
        if (lane >= jb + jj) D[lane][jb + jj] = p[jj];
    }
    __syncthreads();
    // trailing sub-block: D[r][c] -= sum_{j in panel} L[r][j] L[c][j], r >= c >= jb + 8, j ascending
    const int n = HB - jb - 8;
    for (int idx = tid; idx < n * n; idx += 256) {
      const int r = jb + 8 + idx / n, c = jb + 8 + idx % n;
      if (c <= r) {
        double v = D[r][c];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) v = fma(-D[r][jb + jj], D[c][jb + jj], v);
        D[r][c] = v;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < HB * HB; idx += 256) {
    const int r = idx >> 5, c = idx & 31;
    T[o + r][o + c] = (c <= r) ? D[r][c] : 0.0;
  }
  __syncthreads();
}
// Inverse of the lower-triangular 32 x 32 sub-block at (o, o) of T into Dv (same position), all 256 threads, W scratch.
// Leaves: the four 8 x 8 diagonal blocks by the column recurrence x[i] = (delta_ij - sum_{k=j}^{i-1} L[i][k] x[k]) / L[i][i]
// (lane = column; the arithmetic chol.cu's diag_inverse_kernel used for whole 32 x 32 blocks, so results that depend
// on the first 8 columns only - the reference's 2-point goldens - keep their bits).  Then two doubling levels:
// inv([[a, 0], [b, c]]) = [[a^-1, 0], [-c^-1 b a^-1, c^-1]] with the two small products spread over the CTA.  The serial
// chain is 8 divisions long instead of 32 and the dependent fma chains are 8 / 8 / 16 long instead of 496 in total.
__device__ __forceinline__ void inv32_blocked(const Tile& T, Tile& Dv, Tile& W, int o) {
  const int tid = threadIdx.x;
  if (tid < HB) {
    const int lb = tid >> 3, j = tid & 7, oo = o + 8 * lb;  // leaf lb, column j
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double v = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; k++) v = fma(-T[oo + i][oo + k], (k >= j) ? x[k] : 0.0, v);
      x[i] = (i >= j) ? v / T[oo + i][oo + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) Dv[oo + i][oo + j] = x[i];
  }
  __syncthreads();
#pragma unroll
  for (int h = 8; h < HB; h *= 2) {  // pairs of h x h inverses -> 2h x 2h inverses
    const int npair = HB / (2 * h), outs = npair * h * h;
    // W = b a^-1   (b = T[lower-left h x h of the pair], a^-1 = Dv[upper-left])
    for (int idx = tid; idx < outs; idx += 256) {
      const int pr = idx / (h * h), r = (idx / h) % h, c = idx % h, oo = o + pr * 2 * h;
      double v = 0.0;
      for (int k = c; k < h; k++) v = fma(T[oo + h + r][oo + k], Dv[oo + k][oo + c], v);
      W[oo + h + r][oo + c] = v;
    }
    __syncthreads();
    // Dv[lower-left] = -c^-1 W, upper-right = 0
    for (int idx = tid; idx < outs; idx += 256) {
      const int pr = idx / (h * h), r = (idx / h) % h, c = idx % h, oo = o + pr * 2 * h;
      double v = 0.0;
      for (int k = 0; k <= r; k++) v = fma(Dv[oo + h + r][oo + h + k], W[oo + h + k][oo + c], v);
      Dv[oo + h + r][oo + c] = -v;
      Dv[oo + r][oo + h + c] = 0.0;
    }
    __syncthreads();
  }
}

// 32 x 32 x 32 product of sub-blocks on the tensor pipe: 8 warps, warp w owns the 8 x 8 output tiles 2w and 2w + 1 of the
// 4 x 4 tile grid.  A block at (ar, ac) of tile A; TB: B^T with B block at (br, bc) read as B[n][k], else B[k][n].
// acc[t][e] = C[8 (tile / 4) + lane / 4][8 (tile % 4) + 2 (lane % 4) + e], tile = 2 warp + t.
template <bool TB>
__device__ __forceinline__ void mma32(const Tile& A, int ar, int ac, const Tile& B, int br, int bc, double (&acc)[2][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < HB; k0 += 4) {
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const int tile = 2 * warp + t, mt = tile >> 2, nt = tile & 3;
      const double av = A[ar + 8 * mt + fr][ac + k0 + fk];
      const double bv = TB ? B[br + 8 * nt + fr][bc + k0 + fk] : B[br + k0 + fk][bc + 8 * nt + fr];
      dmma884(acc[t][0], acc[t][1], av, bv);
    }
  }
}
template <typename F>
__device__ __forceinline__ void mma32_visit(double (&acc)[2][2], F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 2; t++) {
    const int tile = 2 * warp + t;
    f(8 * (tile >> 2) + (lane >> 2), 8 * (tile & 3) + 2 * (lane & 3), acc[t]);
  }
}

// T (64 x 64, symmetric input, lower part used) -> L in T (upper part zeroed), L^-1 in Dv; W is scratch.  256 threads.
template <bool FAST>
__device__ void diag_factor(Tile& T, Tile& Dv, Tile& W, int col0, int* __restrict__ info) {
  const int tid = threadIdx.x;
  double acc[2][2];
  __syncthreads();
  STAMP(20);
  chol32<FAST>(T, 0, W, col0, info);
  STAMP(21);
  inv32_blocked(T, Dv, W, 0);  // D1 = L11^-1
  STAMP(22);
  // L21 = A21 D1^T (the triangular solve as a product)
  mma32<true>(T, HB, 0, Dv, 0, 0, acc);
  __syncthreads();
  mma32_visit(acc, [&](int r, int c, double (&v)[2]) {
    T[HB + r][c] = v[0];
    T[HB + r][c + 1] = v[1];
  });
  __syncthreads();
  // A22 -= L21 L21^T
  mma32<true>(T, HB, 0, T, HB, 0, acc);
  mma32_visit(acc, [&](int r, int c, double (&v)[2]) {
    T[HB + r][HB + c] -= v[0];
    T[HB + r][HB + c + 1] -= v[1];
  });
  // upper right blocks of L and of L^-1 are zero
  for (int idx = tid; idx < HB * HB; idx += 256) {
    const int r = idx >> 5, c = idx & 31;
    T[r][HB + c] = 0.0;
    Dv[r][HB + c] = 0.0;
  }
  __syncthreads();
  STAMP(23);
  chol32<FAST>(T, HB, W, col0 + HB, info);
  STAMP(24);
  inv32_blocked(T, Dv, W, HB);  // D2 = L22^-1
  STAMP(25);
  // Dv21 = -D2 (L21 D1)
  mma32<false>(T, HB, 0, Dv, 0, 0, acc);
  mma32_visit(acc, [&](int r, int c, double (&v)[2]) {
    W[r][c] = v[0];
    W[r][c + 1] = v[1];
  });
  __syncthreads();
  mma32<false>(Dv, HB, HB, W, 0, 0, acc);
  mma32_visit(acc, [&](int r, int c, double (&v)[2]) {
    Dv[HB + r][c] = -v[0];
    Dv[HB + r][c + 1] = -v[1];
  });
  __syncthreads();
  STAMP(26);
}

struct CoopArgs {
  double* A;    // [Mw, Mw] in: symmetric (lower part used), out: L (strict upper part of the diagonal blocks zeroed)
  double* X;    // [Mw, Mw] out: L^-1 (upper block triangle zeroed)
  double* P;    // split-K partial tiles of the inverse: (G/2)^2 x 64 x 64 doubles
  int Mw, G;
  int* info;
  int fast_pivots;    // see chol32
  long long* stamps;  // lab builds only
};

// chunk length of the split-K partial products of the inverse (products accumulated in registers per task)
__host__ __device__ inline int inv_chunk(int G) { return G <= 16 ? 1 : (G <= 32 ? 2 : 4); }

__global__ void __launch_bounds__(256, 1) chol_inverse_coop_kernel(const CoopArgs a) {
  extern __shared__ __align__(16) unsigned char coop_smem[];
  Tile& As = *reinterpret_cast<Tile*>(coop_smem);
  Tile& Bs = *reinterpret_cast<Tile*>(coop_smem + sizeof(Tile));
  Tile& Cs = *reinterpret_cast<Tile*>(coop_smem + 2 * sizeof(Tile));
  cg::grid_group grid = cg::this_grid();
  const int G = a.G, nb = gridDim.x, b = blockIdx.x;
  const int KC = inv_chunk(G);
  const int64_t ld = a.Mw;
  double* const A = a.A;
  double* const X = a.X;
  auto blk = [&](double* base, int i, int j) { return base + ((int64_t)i * CB) * ld + (int64_t)j * CB; };

  // X[i,j] = -Dinv_i sum_{q=j}^{i-1} L[i,q] X[q,j] (block row i of L^-1) rides along with the factorisation: row i needs
  // L[i, :i] (final after the panel of step i-1), the rows of X above it and Dinv_i (the look-ahead diagonal of step i-1),
  // so its split-K partial products are extra tasks of the PANEL phase of step i and their fixed-order sums times -Dinv_i
  // extra tasks of the UPDATE phase: L^-1 costs no barrier of its own, and the CTAs the shrinking trailing update leaves
  // idle do the work.
  // partial products of row i: task (j, ch) = sum over q in chunk ch of [j, i) of L[i,q] X[q,j]  ->  P[prefix(j) + ch]
  auto x_partials = [&](int i, int first_task, int stride_tasks, int my_first) {
    int tix = first_task;
    for (int j = 0; j < i; j++) {
      const int nch = (i - j + KC - 1) / KC;
      for (int ch = 0; ch < nch; ch++, tix++) {
        if (tix % stride_tasks != my_first) continue;
        double acc[2][4][2];
        acc_zero(acc);
        const int q0 = j + ch * KC, q1 = min(i, q0 + KC);
        for (int q = q0; q < q1; q++) {
          __syncthreads();
          load_tile(As, blk(A, i, q), ld);
          load_tile(Bs, blk(X, q, j), ld);
          __syncthreads();
          tile_mma<false>(As, Bs, acc);
        }
        double* dst = a.P + (int64_t)(tix - first_task) * CB * CB;
        acc_visit(acc, [&](int r, int c, double (&v)[2]) {
          *reinterpret_cast<double2*>(dst + r * CB + c) = make_double2(v[0], v[1]);
        });
      }
    }
    return tix;
  };
  // X[i,j] = -Dinv_i (sum of the partials of (i, j), in chunk order)
  auto x_reduce = [&](int i, int first_task, int stride_tasks, int my_first) {
    int tix = first_task, pfx = 0;
    for (int j = 0; j < i; j++, tix++) {
      const int nch = (i - j + KC - 1) / KC;
      if (tix % stride_tasks == my_first) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < CB * CB / 2; idx += 256) {
          const int r = idx >> 5, c = (idx & 31) * 2;
          const double* p = a.P + (int64_t)pfx * CB * CB + r * CB + c;
          double2 s = *reinterpret_cast<const double2*>(p);
          for (int ch = 1; ch < nch; ch++) {
            const double2 q = *reinterpret_cast<const double2*>(p + (int64_t)ch * CB * CB);
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
      pfx += nch;
    }
    return tix;
  };

  // upper block triangle of X := 0 (the diagonal and lower blocks are all written below)
  for (int t = b; t < G * G; t += nb) {
    const int i = t / G, j = t % G;
    if (j > i)
      for (int idx = threadIdx.x; idx < CB * CB; idx += 256) blk(X, i, j)[(int64_t)(idx >> 6) * ld + (idx & 63)] = 0.0;
  }
  STAMP(0);
  if (b == 0) {  // step 0: the first diagonal block
    load_tile(Cs, blk(A, 0, 0), ld);
    if (a.fast_pivots) diag_factor<true>(Cs, As, Bs, 0, a.info);
    else diag_factor<false>(Cs, As, Bs, 0, a.info);
    store_tile(blk(A, 0, 0), ld, Cs);
    store_tile(blk(X, 0, 0), ld, As);
  }
  STAMP(1);
  grid.sync();
  STAMP(2);
  // CTA 0 carries the critical chain (update of the next diagonal block, then its factorisation); the other CTAs
  // share everything else.  With a single CTA it does it all.
  const int nw = nb > 1 ? nb - 1 : 1;        // CTAs sharing the non-critical tasks
  const int me = nb > 1 ? b - 1 : 0;         // index among them (-1: CTA 0 of a multi-CTA grid takes none)
  for (int k = 0; k < G - 1; k++) {
    // ---- panel: L[i,k] = A[i,k] Dinv_k^T, i > k; plus the partial products of row k of L^-1 -------------------------
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
    // the panel took CTAs 0 .. G-k-2 (round robin): start the X tasks behind them
    x_partials(k, (G - k - 1) % nb, nb, b);
    STAMP(3);
    grid.sync();
    STAMP(4);
    // ---- trailing update A[i,j] -= L[i,k] L[j,k]^T for k < j <= i; task 0 = (k+1, k+1) goes on to factorise it;
    //      plus the finished blocks of row k of L^-1 -----------------------------------------------------------------
    const int n = G - k - 1;             // blocks per side of the trailing matrix
    const int ntask = n * (n + 1) / 2;
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
      // the block being updated travels while the product runs
      double* dst = blk(A, i, j);
      double2 old[2][4];
      {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++)
            old[mt][nt] = *reinterpret_cast<const double2*>(dst + (int64_t)(16 * (warp & 3) + 8 * mt + (lane >> 2)) * ld +
                                                            32 * (warp >> 2) + 8 * nt + 2 * (lane & 3));
      }
      __syncthreads();
      double acc[2][4][2];
      acc_zero(acc);
      tile_mma<true>(As, Bs, acc);
      const bool diag_next = (t == 0);
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          old[mt][nt].x -= acc[mt][nt][0];
          old[mt][nt].y -= acc[mt][nt][1];
        }
      {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const int r = 16 * (warp & 3) + 8 * mt + (lane >> 2), c = 32 * (warp >> 2) + 8 * nt + 2 * (lane & 3);
            if (diag_next) *reinterpret_cast<double2*>(&Cs[r][c]) = old[mt][nt];
            else *reinterpret_cast<double2*>(dst + (int64_t)r * ld + c) = old[mt][nt];
          }
      }
      if (diag_next) {  // look-ahead: this CTA factorises block (k+1, k+1) while the others finish the update
        STAMP(5);
        diag_factor<true>(Cs, As, Bs, (k + 1) * CB, a.info);  // G > 1 implies M > 32
        store_tile(blk(A, k + 1, k + 1), ld, Cs);
        store_tile(blk(X, k + 1, k + 1), ld, As);
      }
    }
    if (me >= 0) x_reduce(k, (ntask - 1) % nw, nw, me);
    STAMP(6);
    grid.sync();
    STAMP(7);
  }
  // ---- the last block row of L^-1 ----------------------------------------------------------------------------------
  if (G > 1) {
    x_partials(G - 1, 0, nb, b);
    grid.sync();
    x_reduce(G - 1, 0, nb, b);
  }
}
}  // namespace

size_t gvi_chol_coop_scratch_doubles(int Mw) {
  const int G = Mw / CB, KC = inv_chunk(G);
  int64_t tasks = 1;  // partial tiles of the last (longest) block row of the inverse
  for (int j = 0; j < G - 1; j++) tasks += (G - 1 - j + KC - 1) / KC;
  return (size_t)tasks * CB * CB;
}

// A, X: [Mw, Mw] row-major, Mw a multiple of 64 (identity-padded by the caller beyond the true size M); *info must hold INT_MAX on entry
// (atomicMin of 1 + the first non-positive pivot column).  One cooperative launch on `st`.
int gvi_launch_chol_inverse_coop(double* A, double* X, double* P, int Mw, int M, int* info, cudaStream_t st,
                                 long long* stamps = nullptr) {
  static std::atomic<unsigned long long> configured{0};
  const size_t smem = 3 * sizeof(Tile);
#ifdef GVI_COOP_STAMPS
  cudaMemcpyToSymbolAsync(g_lab_stamps, &stamps, sizeof(stamps), 0, cudaMemcpyHostToDevice, st);
#endif
  if (gvi_first_use_on_this_device(configured))
    GVI_CUDA(cudaFuncSetAttribute(chol_inverse_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CoopArgs args{A, X, P, Mw, Mw / CB, info, M > HB ? 1 : 0, stamps};
  const int G = Mw / CB;
  int want = G * (G - 1) / 2 + 1;  // widest phase: the first trailing update / the last row of the inverse
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
