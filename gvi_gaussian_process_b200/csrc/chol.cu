// chol.cu - per-step M x M work for one GP: A = K_ZZ + jitter I (+ noise I), Cholesky and L^-1 (one cooperative launch,
// chol_coop.cu), fragment-order packing of L^-1 for the DMMA predict kernel, alpha = A^-1 y.  The per-panel launch chain
// below (chol_panel_kernel) now only serves the public in-place gvi_chol_factor_f64.
// Reference: src/utils/matrix_operations.py:7-28 (a4); cho_factor call sites sparse_posterior_kernel.py:68-74 and
// src/gps/base/base.py:516-521 (a5/a7).  M <= a few thousand: O(M^3) here is <2% of the O(N M^2) predict pass.
#include "common.cuh"

int gvi_launch_ard_gram(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                        const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                        cudaStream_t st);
size_t gvi_gram_dmma_workspace_doubles(int64_t m1, int64_t m2, int D);
int gvi_launch_ard_gram_dmma(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                             const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                             double* workspace, cudaStream_t st);

namespace {
constexpr int NB = 32;

// ---- a4: trace-relative (or absolute) diagonal regulariser, single CTA, fixed-order trace --------------------
__global__ void __launch_bounds__(1024) add_diag_kernel(double* __restrict__ A, int M, int64_t ld, double reg,
                                                        int is_abs, const double* __restrict__ log_noise,
                                                        double* __restrict__ jitter_out,
                                                        int* __restrict__ info_init = nullptr) {
  __shared__ double part[32];
  __shared__ double add_s;
  const int tid = threadIdx.x;
  double s = 0.0;
  if (!is_abs)
    for (int i = tid; i < M; i += 1024) s += A[(int64_t)i * ld + i];
  s = warp_sum(s);
  if ((tid & 31) == 0) part[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double tr = 0.0;
    for (int w = 0; w < 32; w++) tr += part[w];
    double jit = is_abs ? reg : reg * tr / (double)M;
    if (jitter_out) *jitter_out = jit;
    if (info_init) *info_init = 0x7fffffff;  // the factorisation's atomicMin word
    add_s = jit + (log_noise ? exp(log_noise[0]) : 0.0);
  }
  __syncthreads();
  const double add = add_s;
  for (int i = tid; i < M; i += 1024) A[(int64_t)i * ld + i] += add;
}

// identity padding of rows/cols [M, Mp) so that the padded factor is block-diagonal with I
__global__ void pad_identity_kernel(double* __restrict__ A, int M, int Mp) {
  int64_t total = (int64_t)Mp * Mp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(idx / Mp), c = (int)(idx % Mp);
    if (r >= M || c >= M) A[idx] = (r == c) ? 1.0 : 0.0;
  }
}

__device__ __forceinline__ double ld_guard(const double* A, int64_t ld, int M, int r, int c) {
  return (r < M && c < M) ? A[(int64_t)r * ld + c] : (r == c ? 1.0 : 0.0);
}

// 32x32 tile product on the FP64 tensor pipe with the contraction split over warps: this warp takes the k-steps
// (4 columns each) ks = slice, slice + nslice, ... < nks.  A rows come from Ar[r][k] (k contiguous, row stride lda).
// BT = true : B given as rows Br[c][k] (k contiguous)  -> acc[r][c] += sum_k Ar[r][k] Br[c][k]   (L_i L_j^T)
// BT = false: B given as Br[k][c] (c contiguous)       -> acc[r][c] += sum_k Ar[r][k] Br[k][c]   (L X)
// acc[mt][nt][e] = C[8mt + lane/4][8nt + 2(lane%4) + e].  Loads go straight from global/L2 into fragments.
template <bool BT>
__device__ __forceinline__ void tile32_mma(const double* __restrict__ Ar, int64_t lda, const double* __restrict__ Br,
                                           int64_t ldb, int nks, int slice, int nslice, int lane,
                                           double (&acc)[4][4][2]) {
  const int fr = lane >> 2, fk = lane & 3;
#pragma unroll 2
  for (int ks = slice; ks < nks; ks += nslice) {
    const int k0 = ks * 4;
    double a[4], b[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      a[t] = Ar[(int64_t)(8 * t + fr) * lda + k0 + fk];
      b[t] = BT ? Br[(int64_t)(8 * t + fr) * ldb + k0 + fk] : Br[(int64_t)(k0 + fk) * ldb + 8 * t + fr];
    }
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}
__device__ __forceinline__ void tile32_store(double (*dst)[NB + 1], int lane, const double (&acc)[4][4][2]) {
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) dst[8 * mt + (lane >> 2)][8 * nt + (lane & 3) * 2 + e] = acc[mt][nt][e];
}

// ---- blocked left-looking Cholesky: launch t handles the 32-wide block column jb = t ---------------------------
// Panel CTA (ib > jb): S_d = A[jb,jb] - L[jb,:] L[jb,:]^T (recomputed by every CTA, cheap),
// S_r = A[ib,jb] - L[ib,:] L[jb,:]^T, L_d = chol(S_d), L[ib,jb] = S_r L_d^-T.  The two 32 x 32 x (32 jb) products run
// on the tensor pipe (DMMA), 4 warps each, contraction split 4 ways, fragments loaded straight from L2.
// The diagonal block L_d itself is written one launch LATER (CTA 0 of launch t+1 recomputes and stores block t),
// because panel CTAs of launch t still read the original A[jb,jb] while they run.
// Requires M % 32 == 0 (gvi_gp_factor_f64 pads; the public gvi_chol_factor_f64 uses chol_panel_guard_kernel).
__global__ void __launch_bounds__(256) chol_panel_kernel(double* __restrict__ A, int M, int64_t ld, int t,
                                                         int has_diag, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char chol_smem[];
  typedef double Tile[NB][NB + 1];
  Tile* part = reinterpret_cast<Tile*>(chol_smem);  // [4]
  Tile& Sd = part[4];
  Tile& Sr = part[5];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool diag_cta = has_diag && blockIdx.x == 0;
  const int jb = diag_cta ? t - 1 : t;
  const int ib = diag_cta ? jb : t + 1 + (int)blockIdx.x - has_diag;
  const double* Lj = A + (int64_t)jb * NB * ld;
  const double* Li = A + (int64_t)ib * NB * ld;
  const int nks = jb * NB / 4;
  // warps 0-3: S_d partials, warps 4-7: S_r partials
  for (int which = 0; which < 2; which++) {
    if (which == 1 && diag_cta) break;
    if ((warp >> 2) == which) {
      double acc[4][4][2];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
      tile32_mma<true>(which == 0 ? Lj : Li, ld, Lj, ld, nks, warp & 3, 4, lane, acc);
      tile32_store(part[warp & 3], lane, acc);
    }
    __syncthreads();
    if (which == 0 || !diag_cta) {
      const double* Ablk = (which == 0 ? Lj : Li) + jb * NB;
      Tile& dst = which == 0 ? Sd : Sr;
      for (int idx = tid; idx < NB * NB; idx += 256) {
        int r = idx >> 5, c = idx & 31;
        dst[r][c] = Ablk[(int64_t)r * ld + c] - (part[0][r][c] + part[1][r][c] + part[2][r][c] + part[3][r][c]);
      }
    }
    __syncthreads();
  }
  if (warp == 0) {
    // right-looking 32x32 Cholesky in shared memory: lane r owns row r (registers), column j is broadcast from smem
    double row[NB];
#pragma unroll
    for (int c = 0; c < NB; c++) row[c] = Sd[lane][c];
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const double djj = __shfl_sync(0xffffffffu, row[j], j);
      if (lane == j && !(djj > 0.0) && diag_cta && info) atomicMin(info, jb * NB + j + 1);
      const double sq = sqrt(djj);
      row[j] = (lane == j) ? sq : row[j] / sq;
      Sd[lane][j] = row[j];  // column j of L (valid for lanes >= j)
      __syncwarp();
#pragma unroll
      for (int k = j + 1; k < NB; k++) {
        const double lkj = Sd[k][j];  // broadcast
        if (lane >= k) row[k] = fma(-row[j], lkj, row[k]);
      }
    }
#pragma unroll
    for (int c = 0; c < NB; c++) Sd[lane][c] = (c <= lane) ? row[c] : 0.0;
    __syncwarp();
    if (diag_cta) {
      double* out = A + (int64_t)(jb * NB + lane) * ld + jb * NB;
      for (int c = 0; c <= lane; c++) out[c] = Sd[lane][c];
    } else {
      // X L_d^T = S_r, lane owns one row of X
      double x[NB];
#pragma unroll
      for (int c = 0; c < NB; c++) x[c] = Sr[lane][c];
#pragma unroll
      for (int c = 0; c < NB; c++) {
        double v = x[c];
#pragma unroll
        for (int k = 0; k < c; k++) v = fma(-x[k], Sd[c][k], v);
        x[c] = v / Sd[c][c];
      }
      double* out = A + (int64_t)(ib * NB + lane) * ld + jb * NB;
#pragma unroll
      for (int c = 0; c < NB; c++) out[c] = x[c];
    }
  }
}

// generic-size variant with bounds guards (public gvi_chol_factor_f64 on matrices whose size is not a multiple of 32)
__global__ void __launch_bounds__(256) chol_panel_guard_kernel(double* __restrict__ A, int M, int64_t ld, int t,
                                                               int has_diag, int* __restrict__ info) {
  __shared__ double Li[NB][NB + 1];
  __shared__ double Lj[NB][NB + 1];
  __shared__ double Sd[NB][NB + 1];
  __shared__ double Sr[NB][NB + 1];
  const int tid = threadIdx.x;
  const bool diag_cta = has_diag && blockIdx.x == 0;
  const int jb = diag_cta ? t - 1 : t;
  const int ib = diag_cta ? jb : t + 1 + (int)blockIdx.x - has_diag;
  const int tx = tid & 15, ty = tid >> 4;
  double accd[2][2] = {{0, 0}, {0, 0}}, accr[2][2] = {{0, 0}, {0, 0}};
  for (int k0 = 0; k0 < jb * NB; k0 += NB) {
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
      int r = idx >> 5, c = idx & 31;
      Lj[r][c] = ld_guard(A, ld, M, jb * NB + r, k0 + c);
      if (!diag_cta) Li[r][c] = ld_guard(A, ld, M, ib * NB + r, k0 + c);
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < NB; k++) {
      double j0 = Lj[2 * tx][k], j1 = Lj[2 * tx + 1][k];
      double d0 = Lj[2 * ty][k], d1 = Lj[2 * ty + 1][k];
      accd[0][0] = fma(d0, j0, accd[0][0]);
      accd[0][1] = fma(d0, j1, accd[0][1]);
      accd[1][0] = fma(d1, j0, accd[1][0]);
      accd[1][1] = fma(d1, j1, accd[1][1]);
      if (!diag_cta) {
        double i0 = Li[2 * ty][k], i1 = Li[2 * ty + 1][k];
        accr[0][0] = fma(i0, j0, accr[0][0]);
        accr[0][1] = fma(i0, j1, accr[0][1]);
        accr[1][0] = fma(i1, j0, accr[1][0]);
        accr[1][1] = fma(i1, j1, accr[1][1]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      int r = 2 * ty + a, c = 2 * tx + b;
      Sd[r][c] = ld_guard(A, ld, M, jb * NB + r, jb * NB + c) - accd[a][b];
      if (!diag_cta) Sr[r][c] = ld_guard(A, ld, M, ib * NB + r, jb * NB + c) - accr[a][b];
    }
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    double row[NB];
#pragma unroll
    for (int c = 0; c < NB; c++) row[c] = Sd[lane][c];
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double djj = __shfl_sync(0xffffffffu, row[j], j);
      if (lane == j && !(djj > 0.0) && diag_cta && info) atomicMin(info, jb * NB + j + 1);
      double s = sqrt(djj);
      row[j] = (lane == j) ? s : row[j] / s;
#pragma unroll
      for (int k = j + 1; k < NB; k++) {
        double lkj = __shfl_sync(0xffffffffu, row[j], k);
        if (lane >= k) row[k] = fma(-row[j], lkj, row[k]);
      }
    }
#pragma unroll
    for (int c = 0; c < NB; c++) Sd[lane][c] = (c <= lane) ? row[c] : 0.0;
    __syncwarp();
    if (diag_cta) {
      int r = jb * NB + lane;
      if (r < M)
        for (int c = 0; c <= lane; c++) A[(int64_t)r * ld + jb * NB + c] = Sd[lane][c];
    } else {
      double x[NB];
#pragma unroll
      for (int c = 0; c < NB; c++) x[c] = Sr[lane][c];
#pragma unroll
      for (int c = 0; c < NB; c++) {
        double v = x[c];
#pragma unroll
        for (int k = 0; k < c; k++) v = fma(-x[k], Sd[c][k], v);
        x[c] = v / Sd[c][c];
      }
      int r = ib * NB + lane;
      if (r < M) {
#pragma unroll
        for (int c = 0; c < NB; c++)
          if (jb * NB + c < M) A[(int64_t)r * ld + jb * NB + c] = x[c];
      }
    }
  }
}

// ---- header, scaled inducing points -----------------------------------------------------------------------------
__global__ void factor_header_kernel(double* __restrict__ F, FactorLayout f, const double* __restrict__ Z,
                                     const double* __restrict__ log_scaling, const double* __restrict__ log_ls,
                                     int ls_len, double abs_jitter, int frozen) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0) {
    double amp = exp(log_scaling[0]);
    F[0] = amp * amp;
    F[1] = amp * amp;
    F[4] = abs_jitter;
    F[5] = frozen ? 1.0 : 0.0;
    F[6] = F[7] = 0.0;
  }
  if (tid < f.D) F[f.sqrtw_off + tid] = sqrt(exp(log_ls[ls_len == 1 ? 0 : tid]));
  for (int64_t idx = tid; idx < (int64_t)f.Mp * f.D; idx += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(idx / f.D), d = (int)(idx % f.D);
    F[f.zs_off + idx] = (r < f.M) ? Z[idx] * sqrt(exp(log_ls[ls_len == 1 ? 0 : d])) : 0.0;
  }
}

// F[7] = max_j |z~_j - z~_0|^2: the inducing points' half of the guard of the tile kernel's tensor-pipe phase 1
// (predict.cu).  One CTA, after the header kernel.
__global__ void __launch_bounds__(256) zs_radius_kernel(double* __restrict__ F, FactorLayout f) {
  __shared__ double red[256];
  const double* __restrict__ Zs = F + f.zs_off;
  double best = 0.0;
  for (int j = threadIdx.x; j < f.M; j += 256) {
    double s = 0.0;
    for (int d = 0; d < f.D; d++) {
      const double v = Zs[(int64_t)j * f.D + d] - Zs[d];
      s = fma(v, v, s);
    }
    best = (s > best || s != s) ? s : best;  // a NaN radius fails the guard
  }
  red[threadIdx.x] = best;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      const double o = red[threadIdx.x + st];
      if (o > red[threadIdx.x] || o != o) red[threadIdx.x] = o;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) F[7] = red[0];
}

// header of a state built from a caller-supplied Gram: no kernel parameters (k(x,x) arrives per point)
__global__ void factor_header_gram_kernel(double* __restrict__ F, FactorLayout f, double abs_jitter) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0) {
    F[0] = 1.0;
    F[1] = 0.0;
    F[4] = abs_jitter;
    F[5] = 0.0;
    F[6] = F[7] = 0.0;
  }
  if (tid < f.D) F[f.sqrtw_off + tid] = 0.0;
  for (int64_t idx = tid; idx < (int64_t)f.Mp * f.D; idx += (int64_t)gridDim.x * blockDim.x) F[f.zs_off + idx] = 0.0;
}
__global__ void copy_gram_kernel(const double* __restrict__ K, int64_t ldk, int M, double* __restrict__ A, int Mp) {
  const int64_t total = (int64_t)M * M;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    A[(idx / M) * Mp + idx % M] = K[(idx / M) * ldk + idx % M];
}

// ---- pack L^-1 in fragment order: group g (rows R g .. R g + R - 1, R = GVI_GROUP_ROWS), kp (cols 8kp..8kp+7),
// mt (8-row tile), lane, e:  element = Linv[R g + 8 mt + lane/4][8 kp + 4 e + lane%4]; zero above the diagonal and in
// the padding.  One CTA per group.
__global__ void __launch_bounds__(256) pack_linv_kernel(const double* __restrict__ X, int64_t Mp, int M, int G,
                                                        double* __restrict__ W) {
  const int g = blockIdx.x;
  double2* out = reinterpret_cast<double2*>(W + wpack_group_off(g));
  const int items = wpack_nkp(g) * GVI_MT * 32;
  for (int it = threadIdx.x; it < items; it += 256) {
    int lane = it & 31, mt = (it >> 5) % GVI_MT, kp = it / (32 * GVI_MT);
    int c = GVI_GROUP_ROWS * g + 8 * mt + (lane >> 2);
    double2 v;
    int j0 = 8 * kp + (lane & 3), j1 = j0 + 4;
    v.x = (c < M && j0 <= c) ? X[(int64_t)c * Mp + j0] : 0.0;
    v.y = (c < M && j1 <= c) ? X[(int64_t)c * Mp + j1] : 0.0;
    out[it] = v;
  }
}

// L^-T in the same fragment order: group g holds rows j in [R g, R g + R) and the k-range c in [R g, Mp):
// element = Linv[c = 8kp + 4e + lane%4][j = R g + 8mt + lane/4] for kp >= R g / 8.  One CTA per group.
__global__ void __launch_bounds__(256) pack_linvT_kernel(const double* __restrict__ X, int64_t ldx, int Mp, int M,
                                                         int G, double* __restrict__ W) {
  const int g = blockIdx.x;
  double2* out = reinterpret_cast<double2*>(W + wpackT_group_off(g, G));
  const int kp0 = GVI_GROUP_ROWS * g / 8;
  const int items = (Mp / 8 - kp0) * GVI_MT * 32;
  for (int it = threadIdx.x; it < items; it += 256) {
    int lane = it & 31, mt = (it >> 5) % GVI_MT, kpl = it / (32 * GVI_MT);
    int j = GVI_GROUP_ROWS * g + 8 * mt + (lane >> 2);
    int c0 = 8 * (kp0 + kpl) + (lane & 3), c1 = c0 + 4;
    double2 v;
    v.x = (j < M && c0 < M && c0 >= j) ? X[(int64_t)c0 * ldx + j] : 0.0;
    v.y = (j < M && c1 < M && c1 >= j) ? X[(int64_t)c1 * ldx + j] : 0.0;
    out[it] = v;
  }
}
__global__ void copy_linv_kernel(const double* __restrict__ X, int64_t ldx, int Mp, int M, double* __restrict__ out) {
  const int64_t total = (int64_t)Mp * Mp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(idx / Mp), c = (int)(idx % Mp);
    out[idx] = (r < M && c <= r) ? X[(int64_t)r * ldx + c] : 0.0;
  }
}

// ---- alpha = L^-T (L^-1 y): two triangular matvecs with the explicit inverse, fixed summation order ------------
__global__ void __launch_bounds__(256) linv_matvec_kernel(const double* __restrict__ X, int64_t ldx, int Mp, int M,
                                                          const double* __restrict__ y, double* __restrict__ t) {
  // warp per row c: t[c] = sum_{j<=c} X[c][j] y[j]
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Mp) return;
  double s = 0.0;
  if (warp < M)
    for (int j = lane; j <= warp; j += 32) s = fma(X[(int64_t)warp * ldx + j], y[j], s);
  s = warp_sum(s);
  if (lane == 0) t[warp] = s;
}
constexpr int MV_SEG = 8;  // row segments of the transposed matvec (partials combined in segment order)
__global__ void __launch_bounds__(256) linvT_matvec_kernel(const double* __restrict__ X, int64_t ldx, int Mp, int M,
                                                           const double* __restrict__ t, double* __restrict__ part_out) {
  // part_out[seg][j] = sum over the rows c of segment seg, c >= j, of X[c][j] t[c]: 32 columns per CTA (coalesced across
  // lanes), the segment's rows split over the 8 warps with 4 independent accumulators each, fixed-order combination
  __shared__ double part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int seg = blockIdx.y;
  const int rows_per_seg = (M + MV_SEG - 1) / MV_SEG;
  const int c_lo = seg * rows_per_seg, c_hi = min(M, c_lo + rows_per_seg);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (j < M) {
    int c = max(c_lo, blockIdx.x * 32) + warp;
    for (; c + 24 < c_hi; c += 32) {
      const double x0 = X[(int64_t)c * ldx + j], x1 = X[(int64_t)(c + 8) * ldx + j];
      const double x2 = X[(int64_t)(c + 16) * ldx + j], x3 = X[(int64_t)(c + 24) * ldx + j];
      if (c >= j) s0 = fma(x0, t[c], s0);
      if (c + 8 >= j) s1 = fma(x1, t[c + 8], s1);
      if (c + 16 >= j) s2 = fma(x2, t[c + 16], s2);
      if (c + 24 >= j) s3 = fma(x3, t[c + 24], s3);
    }
    for (; c < c_hi; c += 8)
      if (c >= j) s0 = fma(X[(int64_t)c * ldx + j], t[c], s0);
  }
  part[warp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (warp == 0 && j < Mp) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) v += part[w][lane];
    part_out[(int64_t)seg * Mp + j] = v;
  }
}
__global__ void __launch_bounds__(256) matvec_combine_kernel(const double* __restrict__ part, int Mp, double* __restrict__ a) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= Mp) return;
  double v = 0.0;
#pragma unroll
  for (int s = 0; s < MV_SEG; s++) v += part[(int64_t)s * Mp + j];
  a[j] = v;
}
__global__ void zero_kernel(double* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = 0.0;
}
// after the factorisation: info word -> 0 / 1 + first non-PD column, into the factor header and the caller's info
__global__ void info_finish_kernel(double* F, int* info, int Mp, int* caller_info) {
  int v = *info;
  if (v == 0x7fffffff) v = 0;
  *info = v;
  F[3] = (v > Mp) ? 0.0 : (double)v;
  if (caller_info) *caller_info = v;
}
__global__ void info_init_kernel(int* info) { *info = 0x7fffffff; }
__global__ void info_final_kernel(int* info) {
  if (*info == 0x7fffffff) *info = 0;
}
}  // namespace

static int configure_factor_kernels() {
  static std::atomic<unsigned long long> configured{0};
  if (gvi_first_use_on_this_device(configured))
    GVI_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(6 * NB * (NB + 1) * sizeof(double))));
  return GVI_OK;
}

static int launch_chol(double* A, int M, int64_t ld, int* info, cudaStream_t st) {
  if (int rc = configure_factor_kernels()) return rc;
  const int G = (M + NB - 1) / NB;
  for (int t = 0; t <= G; t++) {
    const int has_diag = t > 0 ? 1 : 0;
    const int ctas = has_diag + (t < G ? G - t - 1 : 0);
    if (ctas == 0) continue;
    if (M % NB == 0)
      chol_panel_kernel<<<ctas, 256, 6 * NB * (NB + 1) * sizeof(double), st>>>(A, M, ld, t, has_diag, info);
    else
      chol_panel_guard_kernel<<<ctas, 256, 0, st>>>(A, M, ld, t, has_diag, info);
    GVI_CHECK_LAUNCH("chol_panel_kernel");
  }
  return GVI_OK;
}

extern "C" int gvi_add_diagonal_regulariser_f64(double* A, int M, int64_t ld, double reg, int is_abs, void* stream) {
  GVI_CHECK_ARG(A && M >= 1 && ld >= M, "matrix");
  add_diag_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(A, M, ld, reg, is_abs, nullptr, nullptr);
  GVI_CHECK_LAUNCH("add_diag_kernel");
  return GVI_OK;
}

extern "C" int gvi_chol_factor_f64(double* A, int M, int64_t ld, int* info, void* stream) {
  GVI_CHECK_ARG(A && M >= 1 && ld >= M, "matrix");
  cudaStream_t st = (cudaStream_t)stream;
  if (info) {
    info_init_kernel<<<1, 1, 0, st>>>(info);
    GVI_CHECK_LAUNCH("info_init_kernel");
  }
  int rc = launch_chol(A, M, ld, info, st);
  if (rc) return rc;
  if (info) {
    info_final_kernel<<<1, 1, 0, st>>>(info);
    GVI_CHECK_LAUNCH("info_final_kernel");
  }
  return GVI_OK;
}

extern "C" size_t gvi_gp_factor_bytes(int M, int D) {
  if (M < 1 || D < 1) return 0;
  return (size_t)factor_layout(M, D).total * sizeof(double);
}
size_t gvi_chol_coop_scratch_doubles(int Mw);
int gvi_launch_chol_inverse_coop(double* A, double* X, double* P, int Mw, int M, int* info, cudaStream_t st,
                                 long long* stamps = nullptr);
static inline int factor_ws_dim(int M) { return (M + 63) / 64 * 64; }  // the factorisation works on 64 x 64 blocks

extern "C" size_t gvi_gp_factor_workspace_bytes(int M) {
  if (M < 1) return 0;
  const size_t Mw = (size_t)factor_ws_dim(M);
  return (2 * Mw * Mw + 2 * Mw + 8 + gvi_chol_coop_scratch_doubles((int)Mw)) * sizeof(double);
}

static int factor_impl(const double* Z, int M, int D, const double* log_scaling, const double* log_ls, int ls_len,
                       const double* chol_log_scaling, const double* chol_log_ls, int chol_ls_len, double reg,
                       int is_abs, const double* log_noise, const double* y_z, void* factor_out, void* workspace,
                       int* info, void* stream, const double* Kzz = nullptr, int64_t ldk = 0);

// K_ZZ supplied by the caller (any base kernel): no kernel parameters, no scaled inducing points in the state
extern "C" int gvi_gp_factor_gram_f64(const double* Kzz, int64_t ld, int M, double reg, int is_abs,
                                      const double* log_noise, const double* y_z, void* factor_out, void* workspace,
                                      int* info, void* stream) {
  GVI_CHECK_ARG(Kzz && ld >= M, "Kzz");
  return factor_impl(nullptr, M, 1, nullptr, nullptr, 1, nullptr, nullptr, 0, reg, is_abs, log_noise, y_z, factor_out,
                     workspace, info, stream, Kzz, ld);
}

extern "C" int gvi_gp_factor_f64(const double* Z, int M, int D, const double* log_scaling, const double* log_ls,
                                 int ls_len, double reg, int is_abs, const double* log_noise, const double* y_z,
                                 void* factor_out, void* workspace, int* info, void* stream) {
  return factor_impl(Z, M, D, log_scaling, log_ls, ls_len, nullptr, nullptr, 0, reg, is_abs, log_noise, y_z, factor_out,
                     workspace, info, stream);
}

extern "C" int gvi_gp_factor_frozen_f64(const double* Z, int M, int D, const double* log_scaling,
                                        const double* log_ls, int ls_len, const double* chol_log_scaling,
                                        const double* chol_log_ls, int chol_ls_len, double reg, int is_abs,
                                        void* factor_out, void* workspace, int* info, void* stream) {
  GVI_CHECK_ARG(chol_log_scaling && chol_log_ls, "null pointer");
  GVI_CHECK_ARG(chol_ls_len == 1 || chol_ls_len == D, "chol log_lengthscales must have 1 or D entries");
  return factor_impl(Z, M, D, log_scaling, log_ls, ls_len, chol_log_scaling, chol_log_ls, chol_ls_len, reg, is_abs,
                     nullptr, nullptr, factor_out, workspace, info, stream);
}

__global__ void set_flag_kernel(double* F, int idx, double v) { F[idx] = v; }

// host-side record of which factor buffers carry an extra matrix (the launcher must keep their K tile in one panel)
#include <mutex>
#include <unordered_set>
static std::mutex g_extra_mutex;
static std::unordered_set<const void*> g_extra_factors;
static void note_extra(const void* factor, bool on) {
  std::lock_guard<std::mutex> lock(g_extra_mutex);
  if (on) g_extra_factors.insert(factor); else g_extra_factors.erase(factor);
}
bool gvi_factor_has_extra(const void* factor) {
  std::lock_guard<std::mutex> lock(g_extra_mutex);
  return g_extra_factors.count(factor) != 0;
}

extern "C" int gvi_gp_factor_set_extra_f64(void* factor, int M, int D, const double* E, int64_t ld, void* stream) {
  GVI_CHECK_ARG(factor && M >= 1 && D >= 1, "factor");
  cudaStream_t st = (cudaStream_t)stream;
  const FactorLayout f = factor_layout(M, D);
  double* F = (double*)factor;
  note_extra(factor, E != nullptr);
  if (!E) {
    set_flag_kernel<<<1, 1, 0, st>>>(F, 6, 0.0);
    GVI_CHECK_LAUNCH("set_flag_kernel");
    return GVI_OK;
  }
  GVI_CHECK_ARG(ld >= M, "ld < M");
  pack_linv_kernel<<<f.G, 256, 0, st>>>(E, ld, M, f.G, F + f.wextra_off);
  GVI_CHECK_LAUNCH("pack_linv_kernel");
  set_flag_kernel<<<1, 1, 0, st>>>(F, 6, 1.0);
  GVI_CHECK_LAUNCH("set_flag_kernel");
  return GVI_OK;
}

static int factor_impl(const double* Z, int M, int D, const double* log_scaling, const double* log_ls, int ls_len,
                       const double* chol_log_scaling, const double* chol_log_ls, int chol_ls_len, double reg,
                       int is_abs, const double* log_noise, const double* y_z, void* factor_out, void* workspace,
                       int* info, void* stream, const double* Kzz, int64_t ldk) {
  GVI_CHECK_ARG((Kzz || (Z && log_scaling && log_ls)) && factor_out && workspace, "null pointer");
  GVI_CHECK_ARG(M >= 1 && D >= 1, "sizes");
  GVI_CHECK_ARG(ls_len == 1 || ls_len == D, "log_lengthscales must have 1 or D entries");
  cudaStream_t st = (cudaStream_t)stream;
  const FactorLayout f = factor_layout(M, D);
  double* F = (double*)factor_out;
  note_extra(factor_out, false);  // a fresh factorisation clears the extra-matrix flag (F[6] = 0 in the header kernel)
  // workspace: A and X are [Mw, Mw] with Mw = M rounded up to the 64-wide blocks of the cooperative factorisation
  const int Mw = factor_ws_dim(M);
  double* A = (double*)workspace;
  double* X = A + (size_t)Mw * Mw;
  double* tvec = X + (size_t)Mw * Mw;
  int* winfo = (int*)(tvec + 2 * Mw);
  double* P = tvec + 2 * Mw + 8;
  int rc;
  if (Kzz) {
    factor_header_gram_kernel<<<64, 256, 0, st>>>(F, f, is_abs ? reg : 0.0);
    GVI_CHECK_LAUNCH("factor_header_gram_kernel");
    copy_gram_kernel<<<148 * 4, 256, 0, st>>>(Kzz, ldk, M, A, Mw);
    GVI_CHECK_LAUNCH("copy_gram_kernel");
  } else {
    factor_header_kernel<<<64, 256, 0, st>>>(F, f, Z, log_scaling, log_ls, ls_len, is_abs ? reg : 0.0,
                                             chol_log_scaling != nullptr);
    GVI_CHECK_LAUNCH("factor_header_kernel");
    zs_radius_kernel<<<1, 256, 0, st>>>(F, f);
    GVI_CHECK_LAUNCH("zs_radius_kernel");
  }
  // K_ZZ of the Cholesky: the caller's Gram, the kernel's own parameters, or the frozen (regulariser) parameters
  // Wide inputs (D >= 24): the x.z^T contraction on the tensor pipe (gram_dmma.cu), its operand staging placed in the
  // X half of the workspace, which the factorisation only writes later (M = 2048, D = 784: 1.2 ms -> 0.2 ms per K_ZZ)
  const double* g_ls = chol_log_scaling ? chol_log_scaling : log_scaling;
  const double* g_ll = chol_log_scaling ? chol_log_ls : log_ls;
  const int g_len = chol_log_scaling ? chol_ls_len : ls_len;
  if (Kzz)
    rc = GVI_OK;
  else if (D >= 24 && gvi_gram_dmma_workspace_doubles(M, M, D) <= (size_t)Mw * Mw)
    rc = gvi_launch_ard_gram_dmma(Z, M, Z, M, D, g_ls, g_ll, g_len, A, Mw, X, st);
  else
    rc = gvi_launch_ard_gram(Z, M, Z, M, D, g_ls, g_ll, g_len, A, Mw, st);
  if (rc) return rc;
  add_diag_kernel<<<1, 1024, 0, st>>>(A, M, Mw, reg, is_abs, log_noise, F + 2, winfo);
  GVI_CHECK_LAUNCH("add_diag_kernel");
  if (Mw != M) {
    pad_identity_kernel<<<148, 256, 0, st>>>(A, M, Mw);
    GVI_CHECK_LAUNCH("pad_identity_kernel");
  }
  // Cholesky and L^-1 in ONE cooperative launch (chol_coop.cu)
  if ((rc = gvi_launch_chol_inverse_coop(A, X, P, Mw, M, winfo, st))) return rc;
  info_finish_kernel<<<1, 1, 0, st>>>(F, winfo, f.Mp, info);
  GVI_CHECK_LAUNCH("info_finish_kernel");
  {
    pack_linv_kernel<<<f.G, 256, 0, st>>>(X, Mw, M, f.G, F + f.wpack_off);
    GVI_CHECK_LAUNCH("pack_linv_kernel");
    pack_linvT_kernel<<<f.G, 256, 0, st>>>(X, Mw, f.Mp, M, f.G, F + f.wpackT_off);
    GVI_CHECK_LAUNCH("pack_linvT_kernel");
    copy_linv_kernel<<<148 * 4, 256, 0, st>>>(X, Mw, f.Mp, M, F + f.linv_off);
    GVI_CHECK_LAUNCH("copy_linv_kernel");
  }
  if (y_z) {
    linv_matvec_kernel<<<(f.Mp * 32 + 255) / 256, 256, 0, st>>>(X, Mw, f.Mp, M, y_z, tvec);
    GVI_CHECK_LAUNCH("linv_matvec_kernel");
    // the partials live in the (now free) split-K scratch of the cooperative kernel
    linvT_matvec_kernel<<<dim3(f.Mp / 32, MV_SEG), 256, 0, st>>>(X, Mw, f.Mp, M, tvec, P);
    GVI_CHECK_LAUNCH("linvT_matvec_kernel");
    matvec_combine_kernel<<<(f.Mp + 255) / 256, 256, 0, st>>>(P, f.Mp, F + f.alpha_off);
    GVI_CHECK_LAUNCH("matvec_combine_kernel");
  } else {
    zero_kernel<<<4, 256, 0, st>>>(F + f.alpha_off, f.Mp);
    GVI_CHECK_LAUNCH("zero_kernel");
  }
  return GVI_OK;
}
