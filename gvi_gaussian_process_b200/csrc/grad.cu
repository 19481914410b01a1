// grad.cu - the M x M part of the backward pass of the fused step, and the gradient entry point.
//
// With g_i = dR/dv_q,i (from the tile kernel) and a_i = A^-1 k_i:
//   dR/dlog_ls[d]  = sum_i g_i sum_j a_ij k_ij (x~_id - z~_jd)^2                       (tile kernel, T2_d)
//                  + < dK_ZZ/dlog_ls[d], S >,   S = sum_i g_i a_i a_i^T = A^-1 C A^-1,  C = sum_i g_i k_i k_i^T
//   dR/dlog_scaling = sum_i g_i (2 v_i - 2 jitter_abs |a_i|^2)                          (tile kernel)
// C is a weighted Gram (SYRK) over ALL points: syrk_kernel regenerates K(x,Z) column blocks on the FP64 pipe and
// accumulates 128x128 tiles of C on the tensor pipe (DMMA), interleaved in one instruction stream; S needs four
// M^3 products with the dense L^-1 (plain library GEMMs: cuBLAS); the contraction with dK_ZZ/dlog_ls is a small
// reduction kernel.  Reference: jax.grad of generalised_variational_inference.py:73-94 (trainer.py:83-89).
#include <cublas_v2.h>

#include "predict.cuh"

namespace {
constexpr int CB = 128;  // C tile edge
constexpr int PC = 32;   // points per chunk
constexpr int STHREADS = 256;

// Stage buffer: [kp (4)][tile (16)][lane (32)][2] doubles = the DMMA fragments of 128 columns x 32 points; each
// kp block is padded by 4 doubles so that the 32 lanes of a generating warp (lane = point) hit 16 distinct banks.
constexpr int KPSTRIDE = 16 * 32 * 2 + 4;
constexpr int STAGE = 4 * KPSTRIDE;
__device__ __forceinline__ int frag_index(int c, int i) {
  return (i >> 3) * KPSTRIDE + (((c >> 3) * 32) + ((c & 7) * 4 + (i & 3))) * 2 + ((i >> 2) & 1);
}

struct SyrkParams {
  const double* F;   // factor of q
  const double* X;
  const double* g;   // [N]
  int64_t N;
  int M, Mp, D;
  int nsplit;
  double* Cpart;     // [nsplit][Mp*Mp]
};

// One CTA = one 128x128 tile of C (lower triangle) x one slice of the points.  Per 32-point chunk every warp issues
// 256 DMMAs (its 32x64 sub-tile) and, in the same instruction stream, generates its share of the NEXT chunk's
// k(x_i, z_c) values on the FP64 pipe (lane = point, 32 columns per warp), so the two pipes overlap.
template <int DR>
__global__ void __launch_bounds__(STHREADS, 1) syrk_kernel(const SyrkParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* KA = reinterpret_cast<double*>(smem_raw);   // [2][STAGE] weighted (g_i k_ij), A-fragment order
  double* KB = KA + 2 * STAGE;                         // [2][STAGE] k_il, B-fragment order
  double* zsm = KB + 2 * STAGE;                        // [256][DR] scaled inducing points of the two column blocks
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // lower-triangular tile (bj >= bl) from the linear block index
  int bj = 0, rem = blockIdx.x;
  while (rem > bj) { rem -= bj + 1; bj++; }
  const int bl = rem;
  const FactorLayout f = factor_layout(P.M, P.D);
  const double amp2 = P.F[0];
  const double* __restrict__ sqrtw = P.F + f.sqrtw_off;
  const double* __restrict__ Zs = P.F + f.zs_off;
  for (int idx = tid; idx < 256 * DR; idx += STHREADS) {
    int c = idx / DR, d = idx % DR;
    int zc = (c < CB ? bj : bl) * CB + (c & (CB - 1));
    zsm[idx] = (d < P.D && zc < P.Mp) ? Zs[(int64_t)zc * P.D + d] : 0.0;
  }
  // generating role: warps 0-3 -> A block (weighted by g_i), warps 4-7 -> B block; 32 columns per warp
  const int which = warp >> 2;
  const int gcol0 = (warp & 3) * 32;  // first column (inside the 128-wide block) this warp generates
  double* mine = which == 0 ? KA : KB;

  const int64_t nchunks = (P.N + PC - 1) / PC;
  const int64_t c_begin = nchunks * blockIdx.y / P.nsplit, c_end = nchunks * (blockIdx.y + 1) / P.nsplit;

  double xr[DR], wr = 0.0;  // this lane's point of the chunk being generated (scaled) and its weight
  auto load_point = [&](int64_t chunk) {
    const int64_t r = chunk * PC + lane;
#pragma unroll
    for (int d = 0; d < DR; d++) xr[d] = (d < P.D && r < P.N) ? P.X[r * P.D + d] * sqrtw[d] : 0.0;
    wr = (r < P.N) ? (which == 0 ? P.g[r] : 1.0) : 0.0;
  };
  // four columns at once: all loads first, then four independent exp chains (ILP), then the stores
  auto gen4 = [&](int buf, int cc0) {
    double s4[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double* zc = zsm + (which * CB + gcol0 + cc0 + q) * DR;
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < DR; d++) {
        double diff = xr[d] - zc[d];
        s = fma(diff, diff, s);
      }
      s4[q] = s;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = gcol0 + cc0 + q;
      const int zglob = (which == 0 ? bj : bl) * CB + c;
      s4[q] = (zglob < P.M ? amp2 : 0.0) * wr * gvi_exp(-0.5 * s4[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) mine[buf * STAGE + frag_index(gcol0 + cc0 + q, lane)] = s4[q];
  };

  double acc[4][8][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int wj = warp >> 1, wl = warp & 1;

  __syncthreads();  // zsm ready
  if (c_begin < c_end) {
    load_point(c_begin);
#pragma unroll
    for (int cc = 0; cc < 32; cc += 4) gen4(0, cc);
  }
  for (int64_t chunk = c_begin; chunk < c_end; chunk++) {
    const int cur = (int)((chunk - c_begin) & 1), nxt = cur ^ 1;
    __syncthreads();  // stage `cur` complete; every warp has finished reading stage `nxt` (previous iteration)
    load_point(chunk + 1);  // past the slice end this yields weight 0 / harmless values in the unused stage
    const double2* A2 = reinterpret_cast<const double2*>(KA + cur * STAGE);
    const double2* B2 = reinterpret_cast<const double2*>(KB + cur * STAGE);
#pragma unroll
    for (int kp = 0; kp < PC / 8; kp++) {
      double2 a[4], b[8];
#pragma unroll
      for (int mt = 0; mt < 4; mt++) a[mt] = A2[kp * (KPSTRIDE / 2) + (4 * wj + mt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < 8; nt++) b[nt] = B2[kp * (KPSTRIDE / 2) + (8 * wl + nt) * 32 + lane];
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
      // generation of the next chunk rides between the DMMA blocks (FP64 pipe || tensor pipe)
      gen4(nxt, kp * 8);
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
      gen4(nxt, kp * 8 + 4);
    }
  }
  // write the partial tile: lane holds C[row = lane/4][col = (lane%4)*2 + e] of each 8x8 tile
  double* Cout = P.Cpart + (size_t)blockIdx.y * P.Mp * P.Mp;
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

thread_local cublasHandle_t g_cublas = nullptr;

// row-major C[m x n] = op(A) op(B) through column-major cuBLAS (swap operands)
int gemm_rm(cublasHandle_t h, bool ta, bool tb, int m, int n, int k, const double* A, int lda, const double* B,
            int ldb, double* C, int ldc) {
  const double one = 1.0, zero = 0.0;
  cublasStatus_t s = cublasDgemm(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, n, m, k, &one, B,
                                 ldb, A, lda, &zero, C, ldc);
  if (s != CUBLAS_STATUS_SUCCESS) {
    gvi_set_error("cublasDgemm failed with status %d", (int)s);
    return GVI_ECUDA;
  }
  return GVI_OK;
}

template <int DR>
int launch_syrk(const SyrkParams& P, int ntri, cudaStream_t st) {
  size_t smem = ((size_t)4 * STAGE + 256 * DR) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    GVI_CUDA(cudaFuncSetAttribute(syrk_kernel<DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  syrk_kernel<DR><<<dim3(ntri, P.nsplit), STHREADS, smem, st>>>(P);
  GVI_CHECK_LAUNCH("syrk_kernel");
  return GVI_OK;
}

struct GradWs {
  double *partials, *g, *bscratch, *Cpart, *C, *T1, *T2, *rowpart;
  int nsplit, ntri;
};
size_t align256(size_t x) { return (x + 255) / 256 * 256; }
GradWs carve_grad(void* ws, int64_t N, int M, int D, size_t* total) {
  const FactorLayout f = factor_layout(M, D);
  GradWs w;
  const int Gc = (f.Mp + CB - 1) / CB;
  w.ntri = Gc * (Gc + 1) / 2;
  w.nsplit = GVI_NUM_SMS / w.ntri > 0 ? GVI_NUM_SMS / w.ntri : 1;
  int T = gvi_step_tile_points(f.Mp);
  if (T == 0) T = 8;
  unsigned char* p = (unsigned char*)ws;
  size_t o = 0;
  auto take = [&](size_t bytes) { double* r = (double*)(p + o); o += align256(bytes); return r; };
  w.partials = take((size_t)GVI_NUM_SMS * STEP_PARTIAL_STRIDE * 8);
  w.g = take((size_t)N * 8);
  w.bscratch = take((size_t)GVI_NUM_SMS * f.Mp * T * 8);
  w.Cpart = take((size_t)w.nsplit * f.Mp * f.Mp * 8);
  w.C = take((size_t)f.Mp * f.Mp * 8);
  w.T1 = take((size_t)f.Mp * f.Mp * 8);
  w.T2 = take((size_t)f.Mp * f.Mp * 8);
  w.rowpart = take((size_t)f.Mp * 32 * 8);
  if (total) *total = o;
  return w;
}
}  // namespace

namespace {
__global__ void test_exp_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = gvi_exp(x[i]);
}
}  // namespace
extern "C" int gvi_test_exp_f64(const double* x, int64_t n, double* out, void* stream) {
  GVI_CHECK_ARG(x && out && n >= 1, "arguments");
  test_exp_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  GVI_CHECK_LAUNCH("test_exp_kernel");
  return GVI_OK;
}

extern "C" size_t gvi_fused_step_grad_workspace_bytes(int64_t N, int M, int D) {
  if (N < 1 || M < 1 || D < 1) return 0;
  size_t total = 0;
  carve_grad(nullptr, N, M, D, &total);
  return total;
}

extern "C" int gvi_fused_step_grad_f64(const void* factor_q, const void* factor_p, int M, int D, const double* X,
                                       const double* y, const double* m_q, const double* prior_mean_p,
                                       const double* m_p_in, const double* v_p_in, int64_t N, int reg_kind,
                                       double renyi_alpha, double* out, double* dm_out, double* var_q_out,
                                       double* mean_p_out, double* var_p_out, void* workspace, void* stream) {
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
  P.grad = 1;
  P.g_out = w.g;
  P.dm_out = dm_out;
  P.bscratch = w.bscratch;
  int grid = gvi_launch_step(P, st);
  if (grid < 0) return grid;
  if ((rc = gvi_launch_finish(P.partials, grid, N, 1, D, Fq, out, st))) return rc;
  // C = sum_i g_i k_i k_i^T
  SyrkParams sp{Fq, X, w.g, N, M, f.Mp, D, w.nsplit, w.Cpart};
  if (D <= 1) rc = launch_syrk<1>(sp, w.ntri, st);
  else if (D <= 4) rc = launch_syrk<4>(sp, w.ntri, st);
  else if (D <= 8) rc = launch_syrk<8>(sp, w.ntri, st);
  else if (D <= 16) rc = launch_syrk<16>(sp, w.ntri, st);
  else rc = launch_syrk<32>(sp, w.ntri, st);
  if (rc) return rc;
  syrk_reduce_kernel<<<GVI_NUM_SMS * 4, 256, 0, st>>>(w.Cpart, w.nsplit, f.Mp, w.C);
  GVI_CHECK_LAUNCH("syrk_reduce_kernel");
  // S = L^-T (L^-1 C L^-T) L^-1  (plain library GEMMs)
  if (!g_cublas) {
    if (cublasCreate(&g_cublas) != CUBLAS_STATUS_SUCCESS) {
      gvi_set_error("cublasCreate failed");
      return GVI_ECUDA;
    }
  }
  cublasSetStream(g_cublas, st);
  cublasSetPointerMode(g_cublas, CUBLAS_POINTER_MODE_HOST);
  const double* Linv = Fq + f.linv_off;
  const int n = f.Mp;
  if ((rc = gemm_rm(g_cublas, false, false, n, n, n, Linv, n, w.C, n, w.T1, n))) return rc;   // T1 = Linv C
  if ((rc = gemm_rm(g_cublas, false, true, n, n, n, w.T1, n, Linv, n, w.T2, n))) return rc;   // T2 = T1 Linv^T
  if ((rc = gemm_rm(g_cublas, true, false, n, n, n, Linv, n, w.T2, n, w.T1, n))) return rc;   // T1 = Linv^T T2
  if ((rc = gemm_rm(g_cublas, false, false, n, n, n, w.T1, n, Linv, n, w.T2, n))) return rc;  // S = T1 Linv
  dkzz_contract_kernel<<<M, 256, 0, st>>>(w.T2, Fq, M, f.Mp, D, w.rowpart);
  GVI_CHECK_LAUNCH("dkzz_contract_kernel");
  grad_finish_kernel<<<1, 32, 0, st>>>(w.rowpart, M, D, out);
  GVI_CHECK_LAUNCH("grad_finish_kernel");
  return GVI_OK;
}
