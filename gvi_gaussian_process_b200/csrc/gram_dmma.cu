// gram_dmma.cu - ARD Gram for wide inputs (D >= 24, caller-supplied workspace; correct for any D): k(x,z) = amp^2 exp(-1/2 (|x~|^2 + |z~|^2 - 2 x~.z~)) with the
// x~.z~ contraction on the FP64 tensor pipe (DMMA m8n8k4) and the squared-distance / exp epilogue fused.
// x~ = (x - c) sqrt(w): scaling by the ARD weights and a common shift c (the first row of x2; distances are
// translation invariant, the shift keeps |x~|^2 of the order of the distances so the expansion does not cancel).
//   1. gram_prep_kernel: x~ blocks in DMMA fragment order ([128 rows][32 features] per block) + per-chunk row norms
//   2. gram_dmma_kernel: one CTA per 128x128 output tile; the feature chunks stream through a 3-stage shared-memory
//      ring by cp.async.bulk (TMA 1-D) + mbarrier; 256 DMMAs per warp per chunk; epilogue writes the row-major Gram.
// Reference: src/kernels/standard/ard_kernel.py:58-66 via the double vmap of src/kernels/standard/base.py:95-103.
#include "common.cuh"

namespace {
constexpr int CB = 128;   // rows per block (both operands)
constexpr int KC = 32;    // features per chunk
constexpr int GTHREADS = 256;
constexpr int KPSTRIDE = 16 * 32 * 2 + 4;
constexpr int BLKD = 4 * KPSTRIDE;            // doubles per (row block, feature chunk)
constexpr uint32_t BLKD_BYTES = BLKD * 8;     // 32896, multiple of 16
constexpr int NSTAGE = 3;

__device__ __forceinline__ int frag_index(int c, int i) {
  return (i >> 3) * KPSTRIDE + (((c >> 3) * 32) + ((c & 7) * 4 + (i & 3))) * 2 + ((i >> 2) & 1);
}
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

// sqrtw[d] = sqrt(exp(log_ls[d])), centre_scaled[d] = x2[0][d] * sqrtw[d]
__global__ void gram_scale_kernel(const double* __restrict__ x2, int D, const double* __restrict__ log_ls, int ls_len,
                                  double* __restrict__ sqrtw, double* __restrict__ centre_scaled) {
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
    const double s = sqrt(exp(log_ls[ls_len == 1 ? 0 : d]));
    sqrtw[d] = s;
    centre_scaled[d] = x2[d] * s;
  }
}

// x~ = x * sqrtw - centre_scaled (sqrtw == nullptr: x is already scaled).
// grid (row blocks, feature chunks): lane = feature inside the chunk, each warp takes 16 rows
__global__ void __launch_bounds__(GTHREADS) gram_prep_kernel(const double* __restrict__ x, int64_t m, int D,
                                                             const double* __restrict__ sqrtw,
                                                             const double* __restrict__ centre_scaled, int nchunk,
                                                             double* __restrict__ blocks, double* __restrict__ normpart) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rb = blockIdx.x, ch = blockIdx.y;
  const int d = ch * KC + lane;
  const double sw = (d < D) ? (sqrtw ? sqrtw[d] : 1.0) : 0.0;
  const double c = (d < D) ? centre_scaled[d] : 0.0;
  double* blk = blocks + ((size_t)rb * nchunk + ch) * BLKD;
  for (int rl = warp * 16; rl < warp * 16 + 16; rl++) {
    const int64_t r = (int64_t)rb * CB + rl;
    const double v = (r < m && d < D) ? fma(x[r * D + d], sw, -c) : 0.0;
    blk[frag_index(rl, lane)] = v;
    const double s = warp_sum(v * v);
    if (lane == 0) normpart[((size_t)ch * gridDim.x + rb) * CB + rl] = s;
  }
}

struct GramParams {
  const double* Ab;   // [nrb1][nchunk][BLKD]
  const double* Bb;   // [nrb2][nchunk][BLKD]
  const double* n1;   // [nchunk][nrb1*128]
  const double* n2;   // [nchunk][nrb2*128]
  int nrb1, nrb2, nchunk, D;
  int64_t m1, m2, ld;
  const double* log_scaling;  // amplitude = exp(log_scaling)^2 ...
  const double* amp2_ptr;     // ... or read directly (factor header) when non-null
  double* out;
};

__global__ void __launch_bounds__(GTHREADS, 1) gram_dmma_kernel(const GramParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);                       // [NSTAGE][2][BLKD]
  double* n1s = ring + (size_t)NSTAGE * 2 * BLKD;                           // [128]
  double* n2s = n1s + CB;                                                    // [128]
  uint64_t* full = reinterpret_cast<uint64_t*>(n2s + CB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bj = blockIdx.y, bl = blockIdx.x;
  const int wj = warp >> 1, wl = warp & 1;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // squared norms of the tile's rows / columns: fixed-order sum over the feature chunks
  if (tid < CB) {
    double a = 0.0;
    for (int ch = 0; ch < P.nchunk; ch++) a += P.n1[((size_t)ch * P.nrb1 + bj) * CB + tid];
    n1s[tid] = a;
  } else {
    double a = 0.0;
    for (int ch = 0; ch < P.nchunk; ch++) a += P.n2[((size_t)ch * P.nrb2 + bl) * CB + tid - CB];
    n2s[tid - CB] = a;
  }
  __syncthreads();
  auto issue = [&](int it) {  // thread 0 only
    const int s = it % NSTAGE;
    double* dstA = ring + (size_t)s * 2 * BLKD;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&full[s], 2 * BLKD_BYTES);
    bulk_g2s(dstA, P.Ab + ((size_t)bj * P.nchunk + it) * BLKD, BLKD_BYTES, &full[s]);
    bulk_g2s(dstA + BLKD, P.Bb + ((size_t)bl * P.nchunk + it) * BLKD, BLKD_BYTES, &full[s]);
  };
  if (tid == 0)
    for (int it = 0; it < NSTAGE && it < P.nchunk; it++) issue(it);
  double acc[4][8][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int it = 0; it < P.nchunk; it++) {
    const int s = it % NSTAGE;
    mbar_wait(&full[s], (it / NSTAGE) & 1);
    const double2* A2 = reinterpret_cast<const double2*>(ring + (size_t)s * 2 * BLKD);
    const double2* B2 = A2 + BLKD / 2;
    const int nkp = min(KC / 8, (P.D - it * KC + 7) / 8);  // k-pairs of this chunk that hold features
#pragma unroll 1
    for (int kp = 0; kp < nkp; kp++) {
      double2 a[4], b[8];
#pragma unroll
      for (int mt = 0; mt < 4; mt++) a[mt] = A2[kp * (KPSTRIDE / 2) + (4 * wj + mt) * 32 + lane];
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
    __syncthreads();
    if (tid == 0 && it + NSTAGE < P.nchunk) issue(it + NSTAGE);
  }
  // epilogue: squared distance + exp into a shared-memory tile (the ring is free now), then full-row coalesced stores
  double amp2;
  if (P.amp2_ptr) {
    amp2 = P.amp2_ptr[0];
  } else {
    const double amp = exp(P.log_scaling[0]);
    amp2 = amp * amp;
  }
  constexpr int OSTRIDE = CB + 2;
  double* otile = ring;  // [128][130] doubles = 133 KB
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const int rl = (4 * wj + mt) * 8 + (lane >> 2);
    const double nr = n1s[rl];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int cl = (8 * wl + nt) * 8 + (lane & 3) * 2;
      double2 v;
      // clamp rounding-negative squared distances at 0 with a compare (fmax would turn a NaN input into k = amp^2)
      const double s0 = nr + n2s[cl] - 2.0 * acc[mt][nt][0], s1 = nr + n2s[cl + 1] - 2.0 * acc[mt][nt][1];
      v.x = amp2 * gvi_exp(-0.5 * (s0 < 0.0 ? 0.0 : s0));
      v.y = amp2 * gvi_exp(-0.5 * (s1 < 0.0 ? 0.0 : s1));
      *reinterpret_cast<double2*>(otile + rl * OSTRIDE + cl) = v;
    }
  }
  __syncthreads();
  const int64_t c0 = (int64_t)bl * CB;
  const bool vec_ok = ((P.ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0) && (c0 + CB <= P.m2);
  for (int rl = warp; rl < CB; rl += GTHREADS / 32) {
    const int64_t r = (int64_t)bj * CB + rl;
    if (r >= P.m1) break;
    double* orow = P.out + r * P.ld + c0;
    if (vec_ok) {
#pragma unroll
      for (int h = 0; h < 2; h++)
        reinterpret_cast<double2*>(orow)[h * 32 + lane] =
            *reinterpret_cast<const double2*>(otile + rl * OSTRIDE + (h * 32 + lane) * 2);
    } else {
      for (int cl = lane; cl < CB && c0 + cl < P.m2; cl += 32) orow[cl] = otile[rl * OSTRIDE + cl];
    }
  }
}
}  // namespace

size_t gvi_gram_dmma_workspace_doubles(int64_t m1, int64_t m2, int D) {
  const int64_t nrb1 = (m1 + CB - 1) / CB, nrb2 = (m2 + CB - 1) / CB, nchunk = (D + KC - 1) / KC;
  return (size_t)((nrb1 + nrb2) * nchunk * (BLKD + CB)) + 2 * (size_t)((D + 1) / 2 * 2);
}

// scaled-coordinate core: x1 (scaled by sqrtw1 unless null) against x2 (scaled by sqrtw2 unless null), both shifted by
// centre_scaled; amplitude from log_scaling or amp2_ptr
static int launch_gram_core(const double* x1, int64_t m1, const double* sqrtw1, const double* x2, int64_t m2,
                            const double* sqrtw2, const double* centre_scaled, int D, const double* log_scaling,
                            const double* amp2_ptr, double* out, int64_t ld, double* blocks_ws, cudaStream_t st) {
  const int nrb1 = (int)((m1 + CB - 1) / CB), nrb2 = (int)((m2 + CB - 1) / CB), nchunk = (D + KC - 1) / KC;
  double* Ab = blocks_ws;
  double* Bb = Ab + (size_t)nrb1 * nchunk * BLKD;
  double* n1 = Bb + (size_t)nrb2 * nchunk * BLKD;
  double* n2 = n1 + (size_t)nchunk * nrb1 * CB;
  gram_prep_kernel<<<dim3(nrb1, nchunk), GTHREADS, 0, st>>>(x1, m1, D, sqrtw1, centre_scaled, nchunk, Ab, n1);
  GVI_CHECK_LAUNCH("gram_prep_kernel");
  gram_prep_kernel<<<dim3(nrb2, nchunk), GTHREADS, 0, st>>>(x2, m2, D, sqrtw2, centre_scaled, nchunk, Bb, n2);
  GVI_CHECK_LAUNCH("gram_prep_kernel");
  GramParams P{Ab, Bb, n1, n2, nrb1, nrb2, nchunk, D, m1, m2, ld, log_scaling, amp2_ptr, out};
  size_t smem = ((size_t)NSTAGE * 2 * BLKD + 2 * CB) * sizeof(double) + NSTAGE * sizeof(uint64_t) + 64;
  static std::atomic<unsigned long long> configured{0};
  if (gvi_first_use_on_this_device(configured))
    GVI_CUDA(cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  gram_dmma_kernel<<<dim3(nrb2, nrb1), GTHREADS, smem, st>>>(P);
  GVI_CHECK_LAUNCH("gram_dmma_kernel");
  return GVI_OK;
}

int gvi_launch_ard_gram_dmma(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                             const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                             double* workspace, cudaStream_t st) {
  const int Dp = (D + 1) / 2 * 2;
  double* sqrtw = workspace;
  double* centre = workspace + Dp;
  gram_scale_kernel<<<(D + 255) / 256, 256, 0, st>>>(x2, D, log_ls, ls_len, sqrtw, centre);
  GVI_CHECK_LAUNCH("gram_scale_kernel");
  return launch_gram_core(x1, m1, sqrtw, x2, m2, sqrtw, centre, D, log_scaling, nullptr, out, ld, workspace + 2 * Dp,
                          st);
}

// K(x_chunk, Z) for the wide-input predictive path: the factor buffer F holds sqrtw, the scaled inducing points Z~
// (row 0 of which is the common shift) and amp2.  Rows are written with leading dimension Mp.
int gvi_launch_wide_gram(const double* F, const FactorLayout& f, const double* x, int64_t nr, double* out,
                         double* workspace, cudaStream_t st) {
  const int Dp = (f.D + 1) / 2 * 2;
  return launch_gram_core(x, nr, F + f.sqrtw_off, F + f.zs_off, f.M, nullptr, F + f.zs_off, f.D, nullptr, F, out, f.Mp,
                          workspace + 2 * Dp, st);
}

// K(Z, Z) from the scaled inducing points of the factor state (rows and columns < M; leading dimension Mp) - the
// Hadamard partner of S in the wide-input gradient (grad_wide.cu)
int gvi_launch_wide_gram_zz(const double* F, const FactorLayout& f, double* out, double* workspace, cudaStream_t st) {
  const int Dp = (f.D + 1) / 2 * 2;
  return launch_gram_core(F + f.zs_off, f.M, nullptr, F + f.zs_off, f.M, nullptr, F + f.zs_off, f.D, nullptr, F, out,
                          f.Mp, workspace + 2 * Dp, st);
}
