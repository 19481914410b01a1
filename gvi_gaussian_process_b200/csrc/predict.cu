// predict.cu - the hot kernel: per tile of T points, build K(x,Z) in shared memory (never in HBM), multiply by the
// packed L^-1 on the FP64 tensor pipe (DMMA m8n8k4), and reduce the squared row norms -> predictive variance;
// the same K tile gives the mean (K alpha).  Up to two GPs (q = approximate GP, p = exact regulariser GP) are
// processed per tile and the Gaussian NLL + pGVI/GWI terms are reduced in the epilogue, so one launch is the whole
// "Gram + predict + risk" step and only O(N) bytes cross HBM.  In gradient mode the same launch also produces
// dR/dm_q[N], g_i = dR/dv_q,i [N] and the point-sized parts of dR/dlog_lengthscales (a_i = L^-T b_i by a second
// triangular product per tile); the M x M part is finished in grad.cu.
//
// Reference arithmetic: sparse_posterior_kernel.py:63-96 in diagonal mode (src/kernels/base.py:132-147),
// src/gps/base/base.py:495-526 (exact posterior), negative_log_likelihood.py:41-55, regularisations/projected/*;
// the gradient is jax.grad of generalised_variational_inference.py:73-94 (experiments/shared/trainer.py:83-89).
// Algebra: v = k(x,x) - k^T A^-1 k = k(x,x) - ||L^-1 k||^2 with A = L L^T;  b = L^-1 k is a dependency-free
// triangular GEMM against the explicit inverse (validated against cho_solve in tests; see DESIGN.md numerics).
#include <cstdlib>

#include "predict.cuh"
// the main-loop variants below are `if constexpr (...) { ...; return; }` chains: in an instantiation that takes an
// early branch the rest of the body is unreachable by construction (nvcc warning 128)
#pragma nv_diag_suppress 128

namespace {

// Two launch shapes:
//   forward, M <= 1024: 8 warps per CTA, TWO CTAs per SM, each sweeping its K tile in column panels of about half the
//             shared memory (barriers and tile epilogues of one CTA are covered by the other; +1 % at M=1024, +10 %
//             at M=512).  NB: DFMA and DMMA share one datapath on B200 (tools/fp64_mix.cu: a half/half mix runs at
//             the SUM of the two times), so phase 1 (FP64) can never hide under phase 2 (DMMA) - the kernel's floor
//             is DMMA time + phase-1 FP64 time.
//   forward, larger M and gradient: 16 warps, one CTA per SM (the gradient keeps the whole b tile in shared memory).
constexpr int MT = GVI_MT;    // 8-row m-tiles per group
#ifndef GVI_PB
#define GVI_PB 4
#endif
constexpr int PB = GVI_PB;    // kernel evaluations in flight per thread in phase 1 (independent exp chains)
// A/B switches of tools/ab_build.py (profiles/ncu_final_r02.md has the measurements); the defaults are the product:
#ifndef GVI_P1_MMA
#define GVI_P1_MMA 1        // phase 1 on the tensor pipe behind the per-point guard (0: direct differences only)
#endif
#ifndef GVI_P1_TABLE
#define GVI_P1_TABLE 1      // its exp: table-driven, 12 FP64 instructions (0: gvi_exp, 20)
#endif
#ifndef GVI_SKIP_ZERO
#define GVI_SKIP_ZERO 1     // skip the zero m-tile of a group's last k-pair in phase 2 (two-CTA shape)
#endif
#ifndef GVI_GRAD_TWO_CTA
#define GVI_GRAD_TWO_CTA 1  // gradient mode in the two-CTA panel shape for M <= 1024 (0: one 16-warp CTA per SM)
#endif
constexpr double GVI_P1_BOUND = 512.0;  // guard of the tensor-pipe phase 1: (|u|max + |v|max)^2 in scaled units

// One GVI_GROUP_ROWS-row group of a packed triangular matrix times the T-point tile held in shared memory (fragment order):
// acc[mt][nt] += sum_{kp in [kp0, kp0+nkp)} A(group rows 8mt.., cols 8kp..) * Ktile(cols 8kp.., points 8nt..).
// A fragments come from L2 as one LDG.128 per lane per (kp, mt) (two k-steps), prefetched one kp ahead.
// `Ag` points at the first k-pair to process, `kp0` is the index of that k-pair inside the shared-memory tile;
// the caller initialises acc (zero, or a parked partial sum in the multi-panel path).
// PREFETCH_B: also fetch the B fragments (shared memory) one k-pair ahead - pays in the 16-warp one-CTA shape
// (M=2048: 67.1 -> 65.5 ms, M=4096: 108 -> 101 ms), costs in the 8-warp two-CTA shape (71.6 -> 72.9 ms at M=1024).
#ifndef GVI_DEEP_A
#define GVI_DEEP_A 1
#endif
// `kskip`: index of a k-pair whose m-tile 0 is known to be zero and is not multiplied (-1: none).  The last k-pair of a
// group of a packed LOWER-triangular matrix covers the columns 16 g + 8 .. 16 g + 15, strictly above the rows of m-tile 0:
// 6 of the group's 24 (g + 1) DMMAs, 0.8 % of phase 2 at M = 1024 and 1.5 % at M = 512.
template <int NT, bool PREFETCH_B>
__device__ __forceinline__ void group_mma(const double2* __restrict__ Ag, int kp0, int nkp,
                                          const double2* __restrict__ Ks2, int lane, double (&acc)[MT][NT][2],
                                          int kskip = -1) {
  if constexpr (!PREFETCH_B && GVI_DEEP_A) {
    // A fragments fetched TWO k-pairs ahead through three rotating register buffers (+8 registers): with two CTAs per
    // SM a CTA is alone on the tensor pipe while its neighbour generates a K tile, i.e. two warps per sub-partition -
    // one k-pair of lead (12 DMMAs x 2 warps ~ 400 cycles) is shorter than the L2 latency under load, two are not.
    double2 a0[MT], a1[MT], a2[MT];
    auto lda = [&](double2 (&a)[MT], int k) {
      const int kc = k < nkp ? k : nkp - 1;
#pragma unroll
      for (int mt = 0; mt < MT; mt++) a[mt] = __ldcg(Ag + (kc * MT + mt) * 32);
    };
    auto step = [&](const double2 (&a)[MT], int k) {
      double2 b[NT];
#pragma unroll
      for (int nt = 0; nt < NT; nt++) b[nt] = Ks2[((kp0 + k) * NT + nt) * 32 + lane];
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
    };
    auto tail_step = [&](const double2 (&a)[MT], int k) {  // the group's last k-pair: m-tile 0 is zero there
      double2 b[NT];
#pragma unroll
      for (int nt = 0; nt < NT; nt++) b[nt] = Ks2[((kp0 + k) * NT + nt) * 32 + lane];
#pragma unroll
      for (int mt = 1; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 1; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
    };
    lda(a0, 0);
    lda(a1, 1);
    const int n1 = kskip >= 0 ? nkp - 1 : nkp;  // full steps (the skipped m-tile can only be in the last k-pair)
    int k = 0;
    for (; k + 3 <= n1; k += 3) {
      lda(a2, k + 2);
      step(a0, k);
      lda(a0, k + 3);
      step(a1, k + 1);
      lda(a1, k + 4);
      step(a2, k + 2);
    }
    if (kskip < 0) {
      if (k < nkp) step(a0, k);
      if (k + 1 < nkp) step(a1, k + 1);
    } else if (n1 - k == 0) {
      tail_step(a0, k);
    } else if (n1 - k == 1) {
      step(a0, k);
      tail_step(a1, k + 1);
    } else {
      lda(a2, k + 2);
      step(a0, k);
      step(a1, k + 1);
      tail_step(a2, k + 2);
    }
    return;
  }
  if constexpr (!PREFETCH_B) {
    double2 a_cur[MT], a_nxt[MT];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) a_cur[mt] = __ldcg(Ag + mt * 32);
#pragma unroll 2
    for (int k = 0; k < nkp; k++) {
      const int kn = (k + 1 < nkp) ? k + 1 : k;
#pragma unroll
      for (int mt = 0; mt < MT; mt++) a_nxt[mt] = __ldcg(Ag + (kn * MT + mt) * 32);
      double2 b[NT];
#pragma unroll
      for (int nt = 0; nt < NT; nt++) b[nt] = Ks2[((kp0 + k) * NT + nt) * 32 + lane];
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a_cur[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a_cur[mt].y, b[nt].y);
#pragma unroll
      for (int mt = 0; mt < MT; mt++) a_cur[mt] = a_nxt[mt];
    }
    return;
  }
  double2 a_cur[MT], a_nxt[MT], b_cur[NT], b_nxt[NT];
#pragma unroll
  for (int mt = 0; mt < MT; mt++) a_cur[mt] = __ldcg(Ag + mt * 32);
#pragma unroll
  for (int nt = 0; nt < NT; nt++) b_cur[nt] = Ks2[(kp0 * NT + nt) * 32 + lane];
#pragma unroll 2
  for (int k = 0; k < nkp; k++) {
    const int kn = (k + 1 < nkp) ? k + 1 : k;
#pragma unroll
    for (int mt = 0; mt < MT; mt++) a_nxt[mt] = __ldcg(Ag + (kn * MT + mt) * 32);
#pragma unroll
    for (int nt = 0; nt < NT; nt++) b_nxt[nt] = Ks2[((kp0 + kn) * NT + nt) * 32 + lane];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a_cur[mt].x, b_cur[nt].x);
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a_cur[mt].y, b_cur[nt].y);
#pragma unroll
    for (int mt = 0; mt < MT; mt++) a_cur[mt] = a_nxt[mt];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) b_cur[nt] = b_nxt[nt];
  }
}

template <int NT, int DR, bool GRAD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 8 ? 2 : 1) gp_step_kernel(const StepParams P) {
  constexpr int T = 8 * NT;
  constexpr int THREADS = WARPS * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ksm = reinterpret_cast<double*>(smem_raw);            // [Mp*T] fragment order
  double* xs = Ksm + (size_t)P.Pc * T;                            // [T][DR] scaled x tile
  double* redm = xs + T * DR;                                     // [WARPS][T]
  double* res = redm + WARPS * T;                                 // [2 GPs][T][2] (mean, var)
  double* gs = res + 2 * T * 2;                                   // [T] g_i = dR/dv_q,i of the tile
  const double2* etab = reinterpret_cast<const double2*>(gs + T);  // [32] 2^(j/32) hi / lo (tensor-pipe phase 1)
  double* tail = gs + T + (GVI_P1_TABLE ? 64 : 0);
  const bool multi = P.Pc < P.Mp;
  // [G][T] squared norms per row group: shared memory when it fits, L2 scratch otherwise
  double* reds = P.reds_in_smem ? tail : P.reds_scratch + (size_t)blockIdx.x * P.G * T;
  if (P.reds_in_smem) tail += (size_t)P.G * T;
  double* gpart = tail;                                           // [G][DR+1] per-group gradient partials (GRAD)
  int* counter = reinterpret_cast<int*>(tail + (GRAD ? (size_t)P.G * (DR + 1) : 0));
  double* ascr = multi ? P.acc_scratch + (size_t)blockIdx.x * P.Mp * T : nullptr;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ntiles = (P.N + T - 1) / T;
  const FactorLayout f = factor_layout(P.M, P.D);
  double* bscr = GRAD ? P.bscratch + (size_t)2 * blockIdx.x * P.Mp * T : nullptr;
  // q's K tile is parked next to it (L2 resident): the T2 epilogue of phase 3 reads k_ij back instead of re-evaluating
  // the kernel (37 of its 64 FP64 instructions per entry; the FP64 pipe shares its datapath with the DMMAs)
  double* kscr = GRAD ? P.bscratch + ((size_t)2 * blockIdx.x + 1) * P.Mp * T : nullptr;

  if (GVI_P1_TABLE && threadIdx.x < 32)
    const_cast<double2*>(etab)[threadIdx.x] = GVI_EXP2_TABLE[threadIdx.x];  // visible after the first barrier of the loop
  long long& s_tile = *reinterpret_cast<long long*>(counter + 2);  // inside the 16 spare bytes after the group counter
  constexpr int TILE_STRIDE = GRAD ? STEP_PARTIAL_STRIDE : 2;
  for (int64_t tile = blockIdx.x;; tile += gridDim.x) {
    if (P.tile_counter) {  // dynamic tile order
      __syncthreads();
      if (tid == 0) s_tile = atomicAdd(P.tile_counter, 1);
      __syncthreads();
      tile = s_tile;
    }
    if (tile >= ntiles) break;
    const int64_t row0 = tile * T;
    for (int pi = 0; pi < P.npass; pi++) {
      // gradient mode runs p first so that q's tile (and b = L^-1 k of q) is the one left in shared memory
      const int ps = GRAD ? P.npass - 1 - pi : pi;
      const PassDesc& pd = P.pass[ps];
      const double* __restrict__ F = pd.F;
      const double amp2 = F[0];
      const bool keep_b = GRAD && ps == 0 && !P.grad_pointwise_only;
      // ---- stage the scaled x tile -------------------------------------------------------------------
      for (int idx = tid; idx < T * DR; idx += THREADS) {
        int i = idx / DR, d = idx % DR;
        double v = 0.0;
        if (!pd.Kpre && d < P.D && row0 + i < P.N) v = P.X[(row0 + i) * P.D + d] * F[f.sqrtw_off + d];
        xs[idx] = v;
      }
      if (tid == 0) *counter = 0;
      __syncthreads();
      double qfirst = 0.0;      // ||L^-1 k||^2 of threads < T when a second matrix follows
      bool have_extra = false;
      for (int c_lo = 0; c_lo < P.Mp; c_lo += P.Pc) {  // column panels (one iteration unless M is large)
      const int c_hi = min(c_lo + P.Pc, P.Mp);
      if (c_lo > 0) {
        __syncthreads();  // everyone is done with the previous panel's K tile
        if (tid == 0) *counter = 0;
      }
      // ---- phase 1: K panel (T x Pc) into shared memory in B-fragment order; mean partials ------------------
      bool p1_done = false;
      if constexpr (GVI_P1_MMA && DR >= 4) {  // DR = 32: the 24 A-fragment values spill at the 128-register cap
        // Tensor-pipe form of the exponent (5 <= D): s_ij = |u_i|^2/2 + |v_j|^2/2 - u_i.v_j with u = x~ - c~, v = z~ - c~
        // (c~ = z~_0), the dot product as DMMA m8n8k4 (A = -u for the tile's T points, B = v for 8 columns), the
        // accumulator started at |u_i|^2/2 + |v_j|^2/2 - log amp^2: D + 1 datapath slots per kernel evaluation instead of
        // 2 D + 4, in warp tiles of T x 8 with 2 NT independent exp chains per lane.  The expansion cancels, so a point
        // takes it only while (|u_i| + |v|max)^2 <= GVI_P1_BOUND (error of k <= (D + 3) eps bound / 2 ~ 3e-13 relative at
        // D = 8; |v|max^2 = F[7] from the factorisation); every warp evaluates the same guards from the same values.
        // Points that fail (outliers) are re-evaluated by direct differences; a tile without a passing point (short
        // lengthscales) takes the direct-difference code below as a whole.
        constexpr int KS = DR / 4;
        constexpr bool A_REGS = KS <= 4;  // D > 16: the 3 x 8 A values per lane would spill - re-read from the x tile
        const int q = lane & 3, r = lane >> 2;
        const double* __restrict__ Zs = F + f.zs_off;
        const double* __restrict__ alpha = F + f.alpha_off;
        double cz[KS], a[NT][A_REGS ? KS : 1], hnu[NT];
        bool live_mt[NT];
#pragma unroll
        for (int ks = 0; ks < KS; ks++) cz[ks] = (q + 4 * ks < P.D) ? Zs[q + 4 * ks] : 0.0;
        const double r2 = F[7];
        bool flag[NT];  // point 8 mt + r fails its guard: evaluated by direct differences (fix-up below)
        unsigned fast_live = 0u, slow_live = 0u;
#pragma unroll
        for (int mt = 0; mt < NT; mt++) {
          const bool live = row0 + 8 * mt + r < P.N;
          live_mt[mt] = live;
          double part = 0.0;
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            const double u = live ? xs[(8 * mt + r) * DR + q + 4 * ks] - cz[ks] : 0.0;
            if (A_REGS) a[mt][ks] = -u;
            part = fma(u, u, part);
          }
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          hnu[mt] = 0.5 * part;
          // PER POINT (not per tile): which arithmetic a point gets must not depend on the tile it falls into, so that
          // chunked / sharded / differently tiled launches stay bit identical
          flag[mt] = !(part + r2 + 2.0 * sqrt(part * r2) <= GVI_P1_BOUND);
          fast_live |= __ballot_sync(0xffffffffu, live && !flag[mt]);
          slow_live |= __ballot_sync(0xffffffffu, live && flag[mt]);
        }
        // no point of the tile passes (short lengthscales): the whole tile takes the direct-difference code below
        if (!pd.Kpre && !(P.debug_skip & 1) && fast_live != 0u) {
          p1_done = true;
          const double lamp = log(amp2);
          double mf[NT];
#pragma unroll
          for (int mt = 0; mt < NT; mt++) mf[mt] = 0.0;
          // the next column tile's inducing-point coordinates and alpha travel while this one is evaluated
          const int nct = (c_hi - c_lo) / 8;
          double zr[KS], ar[2];
          auto fetch = [&](int ct) {
            const int j0 = c_lo + 8 * ct;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
              zr[ks] = (q + 4 * ks < P.D) ? Zs[(int64_t)(j0 + r) * P.D + q + 4 * ks] : 0.0;  // cz is 0 there as well
#pragma unroll
            for (int e = 0; e < 2; e++) ar[e] = alpha[j0 + 2 * q + e];
          };
          if (warp < nct) fetch(warp);
          for (int ct = warp; ct < nct; ct += WARPS) {
            const int j0 = c_lo + 8 * ct;
            double b[KS], part = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
              const double v = zr[ks] - cz[ks];
              b[ks] = v;
              part = fma(v, v, part);
            }
            const double aj[2] = {ar[0], ar[1]};
            if (ct + WARPS < nct) fetch(ct + WARPS);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);  // |v|^2 of column j0 + r in the four lanes of row r
            double init[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const double hv = 0.5 * __shfl_sync(0xffffffffu, part, (2 * q + e) * 4);
              const int j = j0 + 2 * q + e;
              init[e] = (j < P.M) ? hv - lamp : 1e300;  // padding columns: exp(-1e300) = 0
            }
            double acc[NT][2];
#pragma unroll
            for (int mt = 0; mt < NT; mt++)
#pragma unroll
              for (int e = 0; e < 2; e++) acc[mt][e] = hnu[mt] + init[e];
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
              for (int mt = 0; mt < NT; mt++) {
                double am;
                if constexpr (A_REGS) am = a[mt][ks];
                else am = live_mt[mt] ? cz[ks] - xs[(8 * mt + r) * DR + q + 4 * ks] : 0.0;
                dmma884(acc[mt][0], acc[mt][1], am, b[ks]);
              }
#pragma unroll
            for (int mt = 0; mt < NT; mt++)
#pragma unroll
              for (int e = 0; e < 2; e++)
                acc[mt][e] = GVI_P1_TABLE ? gvi_exp_neg_t32<true>(acc[mt][e], etab, 1) : gvi_exp(-acc[mt][e]);
            // K tile entry (point i = 8 mt + r, column jl = 8 ct + 2 q + e) in the B-fragment order of phase 2
            double* __restrict__ kbase = Ksm + (size_t)ct * (NT * 64) + r * 8;
#pragma unroll
            for (int mt = 0; mt < NT; mt++)
#pragma unroll
              for (int e = 0; e < 2; e++) {
                const int c = 2 * q + e;
                kbase[mt * 64 + (c >> 2) + 2 * (c & 3)] = acc[mt][e];
                mf[mt] = fma(acc[mt][e], aj[e], mf[mt]);
              }
          }
#pragma unroll
          for (int mt = 0; mt < NT; mt++) {
            double v = mf[mt];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (q == 0 && !(flag[mt] && live_mt[mt]))
              redm[warp * T + 8 * mt + r] = (c_lo == 0) ? v : redm[warp * T + 8 * mt + r] + v;
          }
          if (slow_live != 0u) {
            // fix-up of the points that failed their guard (outliers; rare): the direct-difference arithmetic of the code
            // below, bit for bit, over the column tiles THIS warp wrote above (same-warp ordering of the K-tile stores)
            __syncwarp();
#pragma unroll 1
            for (int mt = 0; mt < NT; mt++) {
              const unsigned fm = __ballot_sync(0xffffffffu, flag[mt] && live_mt[mt]);
#pragma unroll 1
              for (int rr = 0; rr < 8; rr++) {
                if (!((fm >> (4 * rr)) & 1u)) continue;
                const int i = 8 * mt + rr;
                double msum = 0.0;
                for (int ct = warp + WARPS * (lane >> 3); ct < nct; ct += 4 * WARPS) {
                  const int jl = 8 * ct + (lane & 7), j = c_lo + jl;
                  double sa = 0.0, sb = 0.0;
#pragma unroll
                  for (int d = 0; d < DR; d += 2) {
                    const double z0 = (d < P.D) ? Zs[(int64_t)j * P.D + d] : 0.0;
                    const double z1 = (d + 1 < P.D) ? Zs[(int64_t)j * P.D + d + 1] : 0.0;
                    const double d0 = xs[i * DR + d] - z0, d1 = xs[i * DR + d + 1] - z1;
                    sa = fma(d0, d0, sa);
                    sb = fma(d1, d1, sb);
                  }
                  const double kk = ((j < P.M) ? amp2 : 0.0) * gvi_exp(-0.5 * (sa + sb));
                  Ksm[(size_t)(jl >> 3) * (NT * 64) + ((jl >> 2) & 1) + 2 * (jl & 3) + (i >> 3) * 64 + (i & 7) * 8] = kk;
                  msum = fma(kk, alpha[j], msum);
                }
                msum = warp_sum(msum);
                if (lane == 0) redm[warp * T + i] = (c_lo == 0) ? msum : redm[warp * T + i] + msum;
              }
            }
          }
        }
      }
      if (!p1_done && !(P.debug_skip & 1)) {
        double macc[T];
#pragma unroll
        for (int i = 0; i < T; i++) macc[i] = 0.0;
        const double* __restrict__ Zs = F + f.zs_off;
        const double* __restrict__ alpha = F + f.alpha_off;
        if (pd.Kpre) {
          // wide inputs: the K rows of this tile were built by the tensor-pipe Gram kernel; transpose them into the
          // fragment order (coalesced reads along j)
          for (int j = c_lo + tid; j < c_hi; j += THREADS) {
            const double aj = alpha[j];
            const int jl = j - c_lo;
            double* __restrict__ kcol = Ksm + (size_t)(jl >> 3) * (NT * 64) + ((jl >> 2) & 1) + 2 * (jl & 3);
#pragma unroll
            for (int i = 0; i < T; i++) {
              const double k = (j < P.M && row0 + i < P.N) ? __ldcg(pd.Kpre + (row0 + i) * pd.ldk + j) : 0.0;
              kcol[(i >> 3) * 64 + (i & 7) * 8] = k;
              macc[i] = fma(k, aj, macc[i]);
            }
          }
        } else
        for (int j = c_lo + tid; j < c_hi; j += THREADS) {
          double z[DR];
#pragma unroll
          for (int d = 0; d < DR; d++) z[d] = (d < P.D) ? Zs[(int64_t)j * P.D + d] : 0.0;
          const double aj = alpha[j];
          const double scale = (j < P.M) ? amp2 : 0.0;
          const int jl = j - c_lo;
          double* __restrict__ kcol = Ksm + (size_t)(jl >> 3) * (NT * 64) + ((jl >> 2) & 1) + 2 * (jl & 3);
          const double* __restrict__ xr = xs;
          // four points at a time: all loads first, four independent exp chains, then the stores (the compiler
          // cannot prove that the K-tile stores do not alias the x tile, so the batching is explicit)
#pragma unroll
          for (int i0 = 0; i0 < T; i0 += PB) {
            double s4[PB];
#pragma unroll
            for (int q = 0; q < PB; q++) {
              double sa = 0.0, sb = 0.0;
              if constexpr (DR % 2 == 0) {
                const double2* __restrict__ x2 = reinterpret_cast<const double2*>(xr + (i0 + q) * DR);
#pragma unroll
                for (int d = 0; d < DR; d += 2) {
                  const double2 xv = x2[d / 2];  // LDS.128 broadcast
                  const double d0 = xv.x - z[d], d1 = xv.y - z[d + 1];
                  sa = fma(d0, d0, sa);
                  sb = fma(d1, d1, sb);
                }
              } else {
#pragma unroll
                for (int d = 0; d < DR; d++) {
                  const double d0 = xr[(i0 + q) * DR + d] - z[d];
                  sa = fma(d0, d0, sa);
                }
              }
              s4[q] = sa + sb;
            }
#pragma unroll
            for (int q = 0; q < PB; q++) s4[q] = scale * gvi_exp(-0.5 * s4[q]);
#pragma unroll
            for (int q = 0; q < PB; q++) {
              const int i = i0 + q;
              kcol[(i >> 3) * 64 + (i & 7) * 8] = s4[q];
              macc[i] = fma(s4[q], aj, macc[i]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < T; i++) {
          double v = warp_sum(macc[i]);
          if (lane == 0) redm[warp * T + i] = (c_lo == 0) ? v : redm[warp * T + i] + v;
        }
      }
      __syncthreads();
      if (keep_b) {  // park q's K tile for phase 3 (the warps go on to phase 2 as soon as their share is on its way)
        const double2* Ks2r = reinterpret_cast<const double2*>(Ksm);
        double2* kscr2 = reinterpret_cast<double2*>(kscr) + (size_t)c_lo * T / 2;  // this panel's columns
        for (int idx = tid; idx < (c_hi - c_lo) * T / 2; idx += THREADS) __stcg(kscr2 + idx, Ks2r[idx]);
      }
      // ---- phase 2: b = L^-1 k on the tensor pipe, accumulate ||b||^2 ----------------------------------
      auto run_groups = [&](const double* __restrict__ W, const bool store_b) {
        const double2* __restrict__ Ks2 = reinterpret_cast<const double2*>(Ksm);
        const int kp_lo = c_lo / 8, kp_hi = c_hi / 8;
        const int g_lo = c_lo / GVI_GROUP_ROWS;  // groups below g_lo were completed by earlier panels
        while (true) {
          int gi = 0;
          if (lane == 0) gi = atomicAdd(counter, 1);
          gi = __shfl_sync(0xffffffffu, gi, 0);
          if (gi >= P.G - g_lo) break;
          const int g = P.G - 1 - gi;  // largest groups first
          const int nkp_g = wpack_nkp(g);
          const int kp_end = min(kp_hi, nkp_g);
          double acc[MT][NT][2];
          if (c_lo == 0) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
              for (int nt = 0; nt < NT; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
          } else {  // partial sums parked by the previous panel
            const double* src = ascr + (size_t)g * (GVI_GROUP_ROWS * T) + lane;
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
              for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) acc[mt][nt][e] = __ldcg(src + ((mt * NT + nt) * 2 + e) * 32);
          }
          group_mma<NT, WARPS == 16>(reinterpret_cast<const double2*>(W + wpack_group_off(g)) + (size_t)kp_lo * MT * 32 + lane, 0,
                        kp_end - kp_lo, Ks2, lane, acc, GVI_SKIP_ZERO && kp_end == nkp_g ? kp_end - kp_lo - 1 : -1);
          if (kp_end < nkp_g) {  // the group continues in the next panel
            double* dst = ascr + (size_t)g * (GVI_GROUP_ROWS * T) + lane;
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
              for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) __stcg(dst + ((mt * NT + nt) * 2 + e) * 32, acc[mt][nt][e]);
            continue;
          }
          // squared norms of this group's rows: lanes with equal lane%4 hold the same two points of every
          // n-tile.  Stored per GROUP (not per warp) so the final sum has a fixed order whatever warp ran the group.
#pragma unroll
          for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              double v = 0.0;
#pragma unroll
              for (int mt = 0; mt < MT; mt++) v = fma(acc[mt][nt][e], acc[mt][nt][e], v);
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              if (lane < 4) reds[g * T + nt * 8 + lane * 2 + e] = v;
            }
          if (store_b) {
            // b[c][i] goes to the per-CTA scratch (L2 resident) in the K-tile fragment order at "column" c
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
              const int c = GVI_GROUP_ROWS * g + 8 * mt + (lane >> 2);
              double* bcol = bscr + (size_t)(c >> 3) * (NT * 64) + ((c >> 2) & 1) + 2 * (c & 3);
#pragma unroll
              for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                  const int il = (lane & 3) * 2 + e;  // point inside the n-tile
                  __stcg(bcol + nt * 64 + il * 8, acc[mt][nt][e]);
                }
            }
          }
        }
      };
      if (!(P.debug_skip & 2)) run_groups(F + f.wpack_off, keep_b);
      // SVGP family: a second lower-triangular matrix E adds ||E k||^2 to the variance (same K tile, same loop)
      if (F[6] != 0.0 && !multi) {  // (launches with an extra matrix are forced to a single panel by the host)
        __syncthreads();
        if (tid < T) {
          double q0 = 0.0;
          for (int g = 0; g < P.G; g++) q0 += reds[g * T + tid];
          qfirst = q0;
        }
        if (tid == 0) *counter = 0;
        have_extra = true;
        __syncthreads();
        run_groups(F + f.wextra_off, false);
      }
      }  // column panels
      __syncthreads();
      // ---- tile epilogue: variance and mean of this pass --------------------------------------------
      if (tid < T) {
        double q = 0.0, mu = 0.0;
        if (!P.reds_in_smem) {
          for (int g = 0; g < P.G; g++) q += __ldcg(reds + g * T + tid);
        } else {
          for (int g = 0; g < P.G; g++) q += reds[g * T + tid];
        }
#pragma unroll
        for (int w = 0; w < WARPS; w++) mu += redm[w * T + tid];
        const int64_t r = row0 + tid;
        const double kxx = pd.kdiag ? (r < P.N ? pd.kdiag[r] : 0.0) : F[1];
        double var = have_extra ? kxx - qfirst + q : kxx - q;
        if (pd.noise_add) var += exp(pd.noise_add[0]);
        if (r < P.N) {
          if (pd.prior_mean) mu += pd.prior_mean[r];
          if (pd.mean_out) pd.mean_out[r] = mu;
          if (pd.var_out) pd.var_out[r] = var;
        }
        res[(ps * T + tid) * 2 + 0] = mu;
        res[(ps * T + tid) * 2 + 1] = var;
      }
      __syncthreads();
    }  // passes
    // ---- fused risk epilogue (+ pointwise derivatives in gradient mode) --------------------------------
    // Warp 0 owns the tile's T <= 24 points; its sums go to the tile's own record (reduced in a fixed order afterwards, so
    // the result does not depend on which CTA ran the tile) - nothing is carried across tiles in registers.
    if (P.partials && warp == 0) {
      const int64_t r = row0 + tid;
      double g_i = 0.0, part_nll = 0.0, part_reg = 0.0, part_gv = 0.0, part_g = 0.0;
      if (tid >= T) {
        // lanes beyond the tile contribute zeros to the shuffles below
      } else if (GRAD && P.g_in) {
        // vector-Jacobian product of the predictive variance: the caller supplies g_i = dLoss/dvar_i
        if (r < P.N) {
          g_i = P.g_in[r];
          part_gv = fma(g_i, res[(0 * T + tid) * 2 + 1], part_gv);
          part_g += g_i;
        }
      } else if (r < P.N) {
        const double vq = res[(0 * T + tid) * 2 + 1];
        const double mq = P.m_q ? P.m_q[r] : res[(0 * T + tid) * 2 + 0];
        double h_i = 0.0;
        if (P.y) {
          part_nll += gvi_nll_term(P.y[r], mq, vq);
          if (GRAD) {
            double dm, dv;
            gvi_nll_grad(P.y[r], mq, vq, dm, dv);
            h_i += dm;
            g_i += dv;
          }
        }
        if (P.reg_kind != GVI_REG_NONE) {
          double mp, vp;
          if (P.npass > 1) {
            mp = res[(1 * T + tid) * 2 + 0];
            vp = res[(1 * T + tid) * 2 + 1];
          } else {
            mp = P.m_p_in[r];
            vp = P.v_p_in[r];
          }
          part_reg += gvi_reg_term(P.reg_kind, P.renyi_alpha, mp, vp, mq, vq);
          if (GRAD) {
            double dm, dv;
            gvi_reg_grad(P.reg_kind, P.renyi_alpha, mp, vp, mq, vq, dm, dv);
            h_i += dm;
            g_i += dv;
          }
        }
        if (GRAD) {
          part_gv = fma(g_i, vq, part_gv);
          part_g += g_i;
          P.dm_out[r] = h_i;
          P.g_out[r] = g_i;
        }
      }
      if (GRAD && tid < T) gs[tid] = g_i;
      double* rec = P.tile_partials + (size_t)tile * TILE_STRIDE;
      const double a = warp_sum(part_nll), b = warp_sum(part_reg);
      if (lane == 0) {
        rec[0] = a;
        rec[1] = b;
      }
      if (GRAD) {
        const double c = warp_sum(part_gv), d = warp_sum(part_g);
        if (lane == 0) {
          rec[2] = c;
          rec[36] = d;
          rec[37] = 0.0;
        }
        if (P.grad_pointwise_only) rec[3 + lane] = 0.0, rec[4 + lane] = 0.0;  // no phase 3: T2_d and sum g |a|^2 are 0
      }
    }
    // ---- gradient phase 3: a = L^-T b for q, then sum_i g_i a_ij k_ij (x~_id - z~_jd)^2 and sum_i g_i |a_i|^2 ----
    if (GRAD && !P.grad_pointwise_only) {
      const double* __restrict__ F = P.pass[0].F;
      if (P.a_ring && tid < T) {  // the weights g_i ride at the end of every column block of the point's group
        const int64_t p = row0 + tid;
        for (int cb = 0; cb * SYRK_CB < P.Mp; cb++)
          __stcg(P.a_ring + ((size_t)cb * P.a_npg + (p >> 5)) * SYRK_BLK + SYRK_STAGE + (p & 31), gs[tid]);
      }
      const double2* __restrict__ Ks2 = reinterpret_cast<const double2*>(Ksm);
      const double* __restrict__ WT = F + f.wpackT_off;
      const double* __restrict__ Zs = F + f.zs_off;
      // a_j = sum_{c >= j} L^-1[c][j] b_c: group g needs the b rows from 16 g up.  The b tile comes back from the scratch
      // in panels of Pc rows (one panel when the whole tile fits: the 16-warp shape), the TOP panel first - every group
      // starts there - and groups that reach below a panel park their accumulators like phase 2 does.
      for (int c_hi = P.Mp; c_hi > 0;) {
      const int c_lo = (c_hi - 1) / P.Pc * P.Pc;
      __syncthreads();  // the previous panel's (or the passes') readers are done with the tile
      {
        double2* Ks2w = reinterpret_cast<double2*>(Ksm);
        const double2* scr2 = reinterpret_cast<const double2*>(bscr) + (size_t)c_lo * T / 2;
        for (int idx = tid; idx < (c_hi - c_lo) * T / 2; idx += THREADS) Ks2w[idx] = __ldcg(scr2 + idx);  // b panel replaces K
      }
      if (tid == 0) *counter = 0;
      __syncthreads();
      const int ng = c_hi / GVI_GROUP_ROWS;  // groups 0 .. ng - 1 reach into this panel
      while (true) {
        int g = 0;
        if (lane == 0) g = atomicAdd(counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);  // ascending g = largest k-range first
        if (g >= ng) break;
        double acc[MT][NT][2];
        if (c_hi == P.Mp) {
#pragma unroll
          for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        } else {  // partial sums parked by the panel above
          const double* src = ascr + (size_t)g * (GVI_GROUP_ROWS * T) + lane;
#pragma unroll
          for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
              for (int e = 0; e < 2; e++) acc[mt][nt][e] = __ldcg(src + ((mt * NT + nt) * 2 + e) * 32);
        }
        const int kp0 = GVI_GROUP_ROWS * g / 8;                  // first k-pair of the group (absolute)
        const int kp_beg = max(kp0, c_lo / 8), kp_end = c_hi / 8;  // its k-pairs inside this panel
        group_mma<NT, WARPS == 16>(reinterpret_cast<const double2*>(WT + wpackT_group_off(g, P.G)) +
                                       (size_t)(kp_beg - kp0) * MT * 32 + lane,
                                   kp_beg - c_lo / 8, kp_end - kp_beg, Ks2, lane, acc);
        if (GVI_GROUP_ROWS * g < c_lo) {  // the group continues in the panel below
          double* dst = ascr + (size_t)g * (GVI_GROUP_ROWS * T) + lane;
#pragma unroll
          for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
              for (int e = 0; e < 2; e++) __stcg(dst + ((mt * NT + nt) * 2 + e) * 32, acc[mt][nt][e]);
          continue;
        }
        if (P.a_ring) {
          // a_ij of this group -> operand blocks of the S = sum_i g_i a_i a_i^T accumulation (grad.cu): per (mt, nt)
          // the warp covers 8 columns x 8 points = 512 contiguous bytes of the block
#pragma unroll
          for (int mt = 0; mt < MT; mt++) {
            const int j = GVI_GROUP_ROWS * g + 8 * mt + (lane >> 2);
            const int cb = j / SYRK_CB, c = j % SYRK_CB;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
              const int64_t p0 = row0 + nt * 8;  // a multiple of 8: the n-tile stays inside one k-pair of one group
              double* blk = P.a_ring + ((size_t)cb * P.a_npg + (p0 >> 5)) * SYRK_BLK;
              // the operand order pairs point i with point i + 4: lanes q and q ^ 2 (q = lane % 4) swap one value so
              // that every lane owns two ADJACENT doubles -> one 16-byte store per lane, full 32-byte sectors
              const int q = lane & 3;
              const double send = (q & 2) ? acc[mt][nt][0] : acc[mt][nt][1];
              const double recv = __shfl_xor_sync(0xffffffffu, send, 2);
              double2 v;
              v.x = (q & 2) ? recv : acc[mt][nt][0];
              v.y = (q & 2) ? acc[mt][nt][1] : recv;
              // q = 0: points (0, 4); q = 2: points (1, 5); q = 1: points (2, 6); q = 3: points (3, 7)
              const int i_lo = (q & 1) * 2 + (q >> 1);
              __stcg(reinterpret_cast<double2*>(blk + syrk_frag_index(c, (int)(p0 & 31) + i_lo)), v);
            }
          }
        }
        // T2_d = sum_ij g_i a_ij k_ij (x~_id - z~_jd)^2 in two passes, so that neither D running sums nor D squared
        // differences are live across the exp (at the 128-register cap that version spilled ~115 values per group):
        // pass 1 turns every a_ij into coef_ij = g_i a_ij k_ij in place, pass 2 takes one dimension at a time
        double na2 = 0.0;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
          const int j = GVI_GROUP_ROWS * g + 8 * mt + (lane >> 2);
          // k_ij of the parked K tile (B-fragment order of phase 1)
          const double* kcol = kscr + (size_t)(j >> 3) * (NT * 64) + ((j >> 2) & 1) + 2 * (j & 3);
          double kij[NT][2];
#pragma unroll
          for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) kij[nt][e] = __ldcg(kcol + nt * 64 + ((lane & 3) * 2 + e) * 8);
#pragma unroll
          for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const double ga = gs[nt * 8 + (lane & 3) * 2 + e] * acc[mt][nt][e];
              na2 = fma(ga, acc[mt][nt][e], na2);
              acc[mt][nt][e] = ga * kij[nt][e];
            }
        }
#pragma unroll
        for (int d = 0; d < DR; d++) {
          double v = 0.0;
#pragma unroll
          for (int mt = 0; mt < MT; mt++) {
            const int j = GVI_GROUP_ROWS * g + 8 * mt + (lane >> 2);
            const double zd = (d < P.D) ? Zs[(int64_t)j * P.D + d] : 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
              for (int e = 0; e < 2; e++) {
                const double diff = xs[(nt * 8 + (lane & 3) * 2 + e) * DR + d] - zd;
                v = fma(acc[mt][nt][e] * diff, diff, v);
              }
          }
          v = warp_sum(v);
          if (lane == 0) gpart[g * (DR + 1) + d] = v;
        }
        na2 = warp_sum(na2);
        if (lane == 0) gpart[g * (DR + 1) + DR] = na2;
      }
      c_hi = c_lo;
      }  // b panels
      __syncthreads();
      {
        double* rec = P.tile_partials + (size_t)tile * TILE_STRIDE;
        if (tid <= DR) {  // thread d < DR -> T2_d, d == DR -> sum g |a|^2
          double v = 0.0;
          for (int g = 0; g < P.G; g++) v += gpart[g * (DR + 1) + tid];
          rec[tid < DR ? 4 + tid : 3] = v;
        }
        for (int d = DR + tid; d < 32; d += THREADS) rec[4 + d] = 0.0;
      }
      __syncthreads();
    }
  }  // tiles
}

// fixed-order final reduction of per-CTA partial sums.
// out[0..2] = {sum nll, sum reg, N}; gradient mode additionally
// out[3] = sum_i g_i dv_i/dlog_scaling = 2 sum g v - 2 jitter_abs sum g |a|^2   (A built from the same parameters)
//                                     = 4 sum g v - 2 k(x,x) sum g                (frozen Cholesky, Fq[5] == 1)
// out[4+d] = T2_d (point part of d/dlog_ls[d])
__global__ void __launch_bounds__(64) finish_partials_kernel(const double* __restrict__ partials, int nblocks,
                                                             int64_t N, int grad, int D, const double* __restrict__ Fq,
                                                             double* __restrict__ out) {
  const int tid = threadIdx.x;
  if (tid >= STEP_PARTIAL_STRIDE) return;
  double a = 0.0;
  for (int i = 0; i < nblocks; i++) a += partials[(size_t)i * STEP_PARTIAL_STRIDE + tid];
  __shared__ double tot[STEP_PARTIAL_STRIDE];
  tot[tid] = a;
  __syncthreads();
  if (tid == 0) {
    out[0] = tot[0];
    out[1] = tot[1];
    out[2] = (double)N;
    if (grad)
      out[3] = (Fq[5] != 0.0) ? 4.0 * tot[2] - 2.0 * Fq[1] * tot[36] : 2.0 * tot[2] - 2.0 * Fq[4] * tot[3];
  }
  if (grad && tid < D) out[4 + tid] = tot[4 + tid];
}

// stage A of the fixed-order reduction over per-tile records: block b sums the tiles [b*chunk, (b+1)*chunk) column by
// column (thread t owns tiles t, t+256, ...; then a shared-memory tree) into one STEP_PARTIAL_STRIDE record.
__global__ void __launch_bounds__(256) tile_reduce_kernel(const double* __restrict__ tp, int64_t ntiles, int stride,
                                                          double* __restrict__ records) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  const int64_t chunk = (ntiles + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(lo + chunk, ntiles);
  for (int c = 0; c < STEP_PARTIAL_STRIDE; c++) {
    double a = 0.0;
    if (c < stride)
      for (int64_t t = lo + tid; t < hi; t += 256) a += tp[(size_t)t * stride + c];
    red[tid] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (tid < s) red[tid] += red[tid + s];
      __syncthreads();
    }
    if (tid == 0) records[(size_t)blockIdx.x * STEP_PARTIAL_STRIDE + c] = red[0];
    __syncthreads();
  }
}

// dynamic shared memory of one CTA: K panel (Pc columns) + x tile + reductions (+ per-group arrays when they fit)
template <int NT, int DR>
size_t step_smem_bytes(int Mp, int Pc, bool grad, int warps, bool reds_in_smem) {
  constexpr int T = 8 * NT;
  const size_t G = Mp / GVI_GROUP_ROWS;
  size_t doubles = (size_t)Pc * T + T * DR + (size_t)warps * T + 4 * T + T + (GVI_P1_TABLE ? 64 : 0);
  if (reds_in_smem) doubles += G * T;
  if (grad) doubles += G * (DR + 1);
  return doubles * sizeof(double) + 16;
}

constexpr size_t SMEM_LIMIT = 227 * 1024;          // one CTA per SM
constexpr size_t SMEM_LIMIT_2CTA = 113 * 1024;     // two CTAs per SM (228 KB - 2 x 1 KB reserved, halved)

template <int NT, int DR, bool GRAD, int WARPS>
int launch_step_inst2(const StepParams& P, int grid_cap, cudaStream_t st) {
  constexpr int T = 8 * NT;
  size_t smem = step_smem_bytes<NT, DR>(P.Mp, P.Pc, GRAD, WARPS, P.reds_in_smem != 0);
  static std::atomic<unsigned long long> configured{0};  // per instantiation, one bit per device
  if (gvi_first_use_on_this_device(configured)) {
    GVI_CUDA(cudaFuncSetAttribute(gp_step_kernel<NT, DR, GRAD, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(WARPS == 8 ? SMEM_LIMIT_2CTA : SMEM_LIMIT)));
    if (WARPS == 8)
      GVI_CUDA(cudaFuncSetAttribute(gp_step_kernel<NT, DR, GRAD, WARPS>,
                                    cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  int64_t ntiles = (P.N + T - 1) / T;
  int grid = (int)(ntiles < grid_cap ? ntiles : grid_cap);
  StepParams Q = P;
  static const int static_tiles = gvi_debug_env("GVI_STATIC_TILES");
  if (ntiles <= grid || static_tiles) Q.tile_counter = nullptr;  // one tile per CTA: nothing to balance
  if (Q.tile_counter) GVI_CUDA(cudaMemsetAsync(Q.tile_counter, 0, sizeof(int), st));
  gp_step_kernel<NT, DR, GRAD, WARPS><<<grid, WARPS * 32, smem, st>>>(Q);
  GVI_CHECK_LAUNCH("gp_step_kernel");
  return grid;
}
template <int NT, int DR>
int launch_step_inst(const StepParams& P, int grid_cap, cudaStream_t st) {
  if constexpr (NT == 3) {
    if (P.grad && P.two_cta) return launch_step_inst2<NT, DR, true, 8>(P, grid_cap, st);
  }
  if (P.grad) return launch_step_inst2<NT, DR, true, 16>(P, grid_cap, st);
  if constexpr (NT == 3) {
    if (P.two_cta) return launch_step_inst2<NT, DR, false, 8>(P, grid_cap, st);
  }
  return launch_step_inst2<NT, DR, false, 16>(P, grid_cap, st);
}

template <int NT>
int launch_step_nt(const StepParams& P, int grid_cap, cudaStream_t st) {
  if (P.D <= 1 || P.pass[0].Kpre) return launch_step_inst<NT, 1>(P, grid_cap, st);
  if (P.D <= 4) return launch_step_inst<NT, 4>(P, grid_cap, st);
  if (P.D <= 8) return launch_step_inst<NT, 8>(P, grid_cap, st);
  if (P.D <= 16) return launch_step_inst<NT, 16>(P, grid_cap, st);
  return launch_step_inst<NT, 32>(P, grid_cap, st);
}
}  // namespace

// whole-tile (single panel, one CTA per SM) variants: T = 24, 16 or 8 points per tile; 0 if not even T = 8 fits
template <int NT>
static size_t step_smem_bytes_d(int Mp, int Pc, int D, bool grad, int warps, bool reds_in_smem) {
  if (D <= 1) return step_smem_bytes<NT, 1>(Mp, Pc, grad, warps, reds_in_smem);
  if (D <= 4) return step_smem_bytes<NT, 4>(Mp, Pc, grad, warps, reds_in_smem);
  if (D <= 8) return step_smem_bytes<NT, 8>(Mp, Pc, grad, warps, reds_in_smem);
  if (D <= 16) return step_smem_bytes<NT, 16>(Mp, Pc, grad, warps, reds_in_smem);
  return step_smem_bytes<NT, 32>(Mp, Pc, grad, warps, reds_in_smem);
}
int gvi_step_tile_points(int Mp, int D, bool grad) {
  if (step_smem_bytes_d<3>(Mp, Mp, D, grad, 16, true) <= SMEM_LIMIT) return 24;
  if (step_smem_bytes_d<2>(Mp, Mp, D, grad, 16, true) <= SMEM_LIMIT) return 16;
  if (step_smem_bytes_d<1>(Mp, Mp, D, grad, 16, true) <= SMEM_LIMIT) return 8;
  return 0;
}
// points per tile of a forward launch over N points
// (small launches are bound by the tile-pass latency of the SMs that hold a tile: measured at M = 1024, D = 8 a pass over
// T points takes about 12 + 5.6 T us, and a launch ceil(tiles / 148) of them - tools/small_time.py)
int gvi_step_forward_tile_points(int64_t N) {
  if (N >= (int64_t)GVI_NUM_SMS * 24 * 4) return 24;
  int best = 24;
  double best_cost = 0.0;
  for (int T = 24; T >= 8; T -= 8) {
    const int64_t tiles = (N + T - 1) / T;
    const double cost = (double)((tiles + GVI_NUM_SMS - 1) / GVI_NUM_SMS) * (12.0 + 5.6 * T);
    if (T == 24 || cost < best_cost) best = T, best_cost = cost;
  }
  return best;
}
// scratch (doubles) for launches that sweep the K tile in panels: parked accumulators + per-group norms, per CTA
size_t gvi_step_scratch_doubles(int Mp) {
  return (size_t)2 * GVI_NUM_SMS * ((size_t)Mp * 24 + (size_t)(Mp / GVI_GROUP_ROWS) * 24);
}

// forward launches: pick the panel width and the launch shape.  Returns false if nothing fits.
static bool plan_forward(StepParams& P, bool allow_two_cta, bool grad = false) {
  const int D = P.pass[0].Kpre ? 1 : P.D;
  if (allow_two_cta) {
    // two CTAs per SM: the widest panel (multiple of 64 columns) that leaves room for two CTAs; the per-group norms
    // move to the L2 scratch when that saves a panel (gradient mode at M = 1024: 512 instead of 448 columns)
    int best_pc = 0, best_reds = 1;
    for (int reds_smem = 1; reds_smem >= 0; reds_smem--) {
      int pc = P.Mp;  // whole tile if it fits, else the widest multiple of 64 columns
      if (step_smem_bytes_d<3>(P.Mp, pc, D, grad, 8, reds_smem) > SMEM_LIMIT_2CTA) {
        pc = P.Mp / 64 * 64;
        while (pc > 64 && step_smem_bytes_d<3>(P.Mp, pc, D, grad, 8, reds_smem) > SMEM_LIMIT_2CTA) pc -= 64;
      }
      if (step_smem_bytes_d<3>(P.Mp, pc, D, grad, 8, reds_smem) <= SMEM_LIMIT_2CTA && (pc == P.Mp || pc >= 256)) {
        const int panels = (P.Mp + pc - 1) / pc;
        if (best_pc == 0 || panels < (P.Mp + best_pc - 1) / best_pc) best_pc = pc, best_reds = reds_smem;
      }
    }
    if (best_pc > 0) {
      P.two_cta = 1;
      P.Pc = best_pc;
      P.reds_in_smem = best_reds;
      return true;
    }
  }
  P.two_cta = 0;
  for (int reds_smem = 1; reds_smem >= 0; reds_smem--) {
    if (step_smem_bytes_d<3>(P.Mp, P.Mp, D, false, 16, reds_smem) <= SMEM_LIMIT) {
      P.Pc = P.Mp;
      P.reds_in_smem = reds_smem;
      return true;
    }
    if (P.Mp > 1024 && step_smem_bytes_d<3>(P.Mp, 1024, D, false, 16, reds_smem) <= SMEM_LIMIT) {
      P.Pc = 1024;
      P.reds_in_smem = reds_smem;
      return true;
    }
  }
  return false;
}

bool gvi_factor_has_extra(const void* factor);

// returns grid size (>0) or a negative error code
int gvi_launch_step(const StepParams& P_in, cudaStream_t st) {
  StepParams P = P_in;
  P.has_extra = gvi_factor_has_extra(P.pass[0].F) ? 1 : 0;
  static const int debug_skip = gvi_debug_env("GVI_DEBUG_SKIP");
  static const int one_cta = gvi_debug_env("GVI_ONE_CTA");
  P.debug_skip = debug_skip;
  const bool pre = P.pass[0].Kpre != nullptr;
  if (P.D > 32 && !pre) {
    gvi_set_error("the fused step supports D <= 32 in this build (got D=%d); gvi_gp_predict_f64 handles wider inputs",
                  P.D);
    return GVI_EUNSUPPORTED;
  }
  if (!P.grad) {
    // the SVGP second product needs the whole K tile (single panel); everything else may use two CTAs per SM
    const bool allow_two = !one_cta && !P.has_extra && P.acc_scratch != nullptr && P.Mp <= 1024;
    if (!plan_forward(P, allow_two) || (P.has_extra && P.Pc < P.Mp)) {
      gvi_set_error("no launch shape fits M=%d (extra matrix: %d)", P.M, P.has_extra);
      return GVI_EUNSUPPORTED;
    }
    if ((P.Pc < P.Mp || !P.reds_in_smem) && !P.acc_scratch) {
      gvi_set_error("M=%d needs the panel scratch (pass a workspace of gvi_gp_predict_workspace_bytes)", P.M);
      return GVI_EINVAL;
    }
    if (P.acc_scratch) P.reds_scratch = P.acc_scratch + (size_t)2 * GVI_NUM_SMS * P.Mp * 24;
    // profiling aid (env GVI_DEBUG_GRID): cap the grid, e.g. 148 in the two-CTA shape = one 8-warp CTA per SM
    static const int debug_grid = gvi_debug_env("GVI_DEBUG_GRID");
    if (debug_grid > 0) return launch_step_nt<3>(P, debug_grid, st);
    // small launches: 16- or 8-point tiles spread the points over more SMs (a tile-pass is bound by one SM's tensor
    // pipe; N = 1024: 43 tiles of 24 points against 128 tiles of 8, 182 -> 59 us)
    const int T = P.tile_points ? P.tile_points : gvi_step_forward_tile_points(P.N);
    if (T != 24) P.two_cta = 0;  // the two-CTA shape exists for 24-point tiles only; Pc / reds_in_smem stay valid (less smem)
    if (T == 8) return launch_step_nt<1>(P, GVI_NUM_SMS, st);
    if (T == 16) return launch_step_nt<2>(P, GVI_NUM_SMS, st);
    return launch_step_nt<3>(P, P.two_cta ? 2 * GVI_NUM_SMS : GVI_NUM_SMS, st);
  }
  // gradient mode, M <= 1024 with a panel scratch: the forward launch shape - two 8-warp CTAs per SM, K and b tiles in
  // panels (phase 1 and the T2 epilogue of one CTA hide under the other's DMMAs instead of being exposed)
  if (GVI_GRAD_TWO_CTA && !one_cta && !P.has_extra && P.acc_scratch && P.Mp <= 1024 && !P.grad_pointwise_only &&
      gvi_step_tile_points(P.Mp, P.D, true) == 24 && plan_forward(P, true, true) && P.two_cta) {
    P.reds_scratch = P.acc_scratch + (size_t)2 * GVI_NUM_SMS * P.Mp * 24;
    return launch_step_nt<3>(P, 2 * GVI_NUM_SMS, st);
  }
  // otherwise the whole b tile stays in shared memory (one 16-warp CTA per SM)
  P.Pc = P.Mp;
  P.two_cta = 0;
  P.reds_in_smem = 1;
  switch (gvi_step_tile_points(P.Mp, P.D, true)) {
    case 24: return launch_step_nt<3>(P, GVI_NUM_SMS, st);
    case 16: return launch_step_nt<2>(P, GVI_NUM_SMS, st);
    case 8: return launch_step_nt<1>(P, GVI_NUM_SMS, st);
    default:
      gvi_set_error("the gradient step supports M <= 3456 in this build (got M=%d)", P.M);
      return GVI_EUNSUPPORTED;
  }
}

int gvi_launch_tile_reduce(const double* tile_partials, int64_t ntiles, int grad, double* records, cudaStream_t st) {
  tile_reduce_kernel<<<STEP_REDUCE_BLOCKS, 256, 0, st>>>(tile_partials, ntiles, grad ? STEP_PARTIAL_STRIDE : 2, records);
  GVI_CHECK_LAUNCH("tile_reduce_kernel");
  return STEP_REDUCE_BLOCKS;
}

int gvi_launch_finish(const double* partials, int nblocks, int64_t N, int grad, int D, const double* Fq, double* out,
                      cudaStream_t st) {
  finish_partials_kernel<<<1, 64, 0, st>>>(partials, nblocks, N, grad, D, Fq, out);
  GVI_CHECK_LAUNCH("finish_partials_kernel");
  return GVI_OK;
}

size_t gvi_gram_dmma_workspace_doubles(int64_t m1, int64_t m2, int D);
int gvi_launch_ard_gram_dmma(const double* x1, int64_t m1, const double* x2, int64_t m2, int D,
                             const double* log_scaling, const double* log_ls, int ls_len, double* out, int64_t ld,
                             double* workspace, cudaStream_t st);

int gvi_launch_wide_gram(const double* F, const FactorLayout& f, const double* x, int64_t nr, double* out,
                         double* workspace, cudaStream_t st);

// wide inputs (D > 32): K(x_chunk, Z) is built CHUNK rows at a time by the tensor-pipe Gram kernel into a ring and the
// tile kernel consumes it.  The ring need not stay in L2: 16 B per K entry written + read against ~M flop of tensor
// work on it (M = 2048: 5 ns of HBM time per point against 250 ns of arithmetic).  What matters is that both kernels
// run whole waves: a chunk is k x (148 x 24) rows, k <= 8 tiles per CTA of the tile kernel (dynamic tile order), and
// for k = 8 or 4 (with Mp a multiple of 256) also a whole number of waves of the Gram kernel's 128 x 128 tiles
// (28 416 rows = 222 row tiles; r02 first version: 3552-row chunks = 448 Gram tiles = 3.03 waves, 25 % of the Gram
// kernel idle, and one exposed K-tile transposition per launch).  Ring <= 512 MB.
constexpr int64_t WIDE_CHUNK = (int64_t)GVI_NUM_SMS * 24;
static int64_t wide_chunk_rows(int Mp, int64_t N) {
  int64_t k = (512LL << 20) / (WIDE_CHUNK * (int64_t)Mp * 8);
  k = k < 1 ? 1 : (k >= 8 ? 8 : (k >= 4 ? 4 : k));
  int64_t rows = k * WIDE_CHUNK;
  const int64_t need = (N + 23) / 24 * 24;
  return need < rows ? (need < 24 ? 24 : need) : rows;
}
static size_t wide_ring_doubles(int M, int Mp, int D, int64_t N, int nrings) {
  const int64_t rows = wide_chunk_rows(Mp, N);
  return (size_t)nrings * rows * Mp + gvi_gram_dmma_workspace_doubles(rows, M, D);
}

extern "C" size_t gvi_gp_predict_workspace_bytes_n(int M, int D, int64_t N) {
  if (M < 1 || D < 1 || N < 0) return 0;
  const int Mp = (M + 31) / 32 * 32;
  size_t doubles = STEP_COUNTER_DOUBLES + gvi_step_scratch_doubles(Mp);
  if (D > 32) doubles += wide_ring_doubles(M, Mp, D, N, 1);
  return doubles * sizeof(double);
}
// enough for any N (D > 32: the full-sized ring)
extern "C" size_t gvi_gp_predict_workspace_bytes(int M, int D) {
  return gvi_gp_predict_workspace_bytes_n(M, D, INT64_MAX / 2);
}

extern "C" int gvi_gp_predict_f64(const void* factor, int M, int D, const double* X, int64_t N,
                                  const double* prior_mean, const double* log_noise_add, double* mean_out,
                                  double* var_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(M >= 1 && D >= 1 && N >= 0, "sizes");
  if (N == 0) return GVI_OK;
  GVI_CHECK_ARG(factor && X, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  StepParams P{};
  const FactorLayout f = factor_layout(M, D);
  P.npass = 1;
  P.M = M; P.Mp = f.Mp; P.D = D; P.G = f.G;
  P.partials = nullptr;
  // workspace = [tile counter][panel scratch][wide-input ring ...]
  P.tile_counter = workspace ? (int*)workspace : nullptr;
  P.acc_scratch = workspace ? (double*)workspace + STEP_COUNTER_DOUBLES : nullptr;
  if (D <= 32) {
    P.pass[0] = PassDesc{(const double*)factor, prior_mean, log_noise_add, mean_out, var_out, nullptr, 0};
    P.X = X; P.N = N;
    int rc = gvi_launch_step(P, st);
    return rc < 0 ? rc : GVI_OK;
  }
  GVI_CHECK_ARG(workspace, "D > 32 needs a workspace of gvi_gp_predict_workspace_bytes(M, D)");
  const double* F = (const double*)factor;
  const int64_t rows = wide_chunk_rows(f.Mp, N);
  double* ring = (double*)workspace + STEP_COUNTER_DOUBLES + gvi_step_scratch_doubles(f.Mp);
  double* gws = ring + (size_t)rows * f.Mp;
  // the factor buffer carries sqrt(w), the pre-scaled inducing points and amp^2: everything the Gram needs
  for (int64_t r0 = 0; r0 < N; r0 += rows) {
    const int64_t nr = N - r0 < rows ? N - r0 : rows;
    int rc = gvi_launch_wide_gram(F, f, X + r0 * D, nr, ring, gws, st);
    if (rc) return rc;
    P.pass[0] = PassDesc{F, prior_mean ? prior_mean + r0 : nullptr, log_noise_add, mean_out ? mean_out + r0 : nullptr,
                         var_out ? var_out + r0 : nullptr, ring, (int64_t)f.Mp};
    P.X = nullptr; P.N = nr;
    rc = gvi_launch_step(P, st);
    if (rc < 0) return rc;
  }
  return GVI_OK;
}

// any base kernel: the K rows come from the caller (materialised by that kernel's own Gram routine, chunk by chunk on
// the host side), k(x_i, x_i) per point; the triangular products, norms and the mean run in the same tile kernel
extern "C" int gvi_gp_predict_gram_f64(const void* factor, int M, const double* Kxz, int64_t ld, const double* kdiag,
                                       int64_t N, const double* prior_mean, const double* log_noise_add,
                                       double* mean_out, double* var_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(M >= 1 && N >= 0, "sizes");
  if (N == 0) return GVI_OK;
  GVI_CHECK_ARG(factor && Kxz && kdiag && ld >= M, "null pointer / ld < M");
  StepParams P{};
  const FactorLayout f = factor_layout(M, 1);
  P.npass = 1;
  P.M = M; P.Mp = f.Mp; P.D = 1; P.G = f.G;
  P.partials = nullptr;
  P.tile_counter = workspace ? (int*)workspace : nullptr;
  P.acc_scratch = workspace ? (double*)workspace + STEP_COUNTER_DOUBLES : nullptr;
  P.pass[0] = PassDesc{(const double*)factor, prior_mean, log_noise_add, mean_out, var_out, Kxz, ld, kdiag};
  P.X = nullptr; P.N = N;
  int rc = gvi_launch_step(P, (cudaStream_t)stream);
  return rc < 0 ? rc : GVI_OK;
}

// fused-step workspace = [tile counter][reduction records][per-tile (nll, reg)][panel scratch]
static size_t fused_records_doubles() {
  const size_t n = (size_t)(2 * GVI_NUM_SMS > STEP_REDUCE_BLOCKS ? 2 * GVI_NUM_SMS : STEP_REDUCE_BLOCKS);
  return n * STEP_PARTIAL_STRIDE;
}
static size_t fused_tile_doubles(int64_t N) {
  const int T = gvi_step_forward_tile_points(N);
  return (size_t)((N + T - 1) / T) * 2 + 2;
}
extern "C" size_t gvi_fused_step_workspace_bytes(int64_t N, int M) {
  if (N < 1 || M < 1) return 0;
  const int Mp = (M + 31) / 32 * 32;
  return (STEP_COUNTER_DOUBLES + fused_records_doubles() + fused_tile_doubles(N) + gvi_step_scratch_doubles(Mp)) *
         sizeof(double);
}

int gvi_fill_step_params(StepParams& P, const void* factor_q, const void* factor_p, int M, int D, const double* X,
                         const double* y, const double* m_q, const double* prior_mean_p, const double* m_p_in,
                         const double* v_p_in, int64_t N, int reg_kind, double renyi_alpha, double* var_q_out,
                         double* mean_p_out, double* var_p_out) {
  GVI_CHECK_ARG(M >= 1 && D >= 1 && N >= 1, "sizes");
  GVI_CHECK_ARG(factor_q && X && m_q, "null pointer");
  GVI_CHECK_ARG(reg_kind >= GVI_REG_NONE && reg_kind <= GVI_REG_SQUARED_DIFFERENCE, "reg_kind");
  GVI_CHECK_ARG(reg_kind == GVI_REG_NONE || factor_p || (m_p_in && v_p_in),
                "regulariser needs factor_p or (m_p_in, v_p_in)");
  const FactorLayout f = factor_layout(M, D);
  P.pass[0] = PassDesc{(const double*)factor_q, nullptr, nullptr, nullptr, var_q_out, nullptr, 0};
  P.npass = 1;
  if (factor_p) {
    P.pass[1] = PassDesc{(const double*)factor_p, prior_mean_p, nullptr, mean_p_out, var_p_out, nullptr, 0};
    P.npass = 2;
  }
  P.M = M; P.Mp = f.Mp; P.D = D; P.G = f.G;
  P.X = X; P.N = N;
  P.y = y; P.m_q = m_q; P.m_p_in = m_p_in; P.v_p_in = v_p_in;
  P.reg_kind = reg_kind; P.renyi_alpha = renyi_alpha;
  return GVI_OK;
}

// D > 32 adds the two K-row rings (q and p) and the Gram kernel's workspace; equal to the call above for D <= 32
extern "C" size_t gvi_fused_step_workspace_bytes_d(int64_t N, int M, int D) {
  size_t bytes = gvi_fused_step_workspace_bytes(N, M);
  if (bytes == 0 || D <= 32) return bytes;
  const int Mp = (M + 31) / 32 * 32;
  return bytes + wide_ring_doubles(M, Mp, D, N, 2) * sizeof(double);
}

extern "C" int gvi_fused_step_f64(const void* factor_q, const void* factor_p, int M, int D, const double* X,
                                  const double* y, const double* m_q, const double* prior_mean_p,
                                  const double* m_p_in, const double* v_p_in, int64_t N, int reg_kind,
                                  double renyi_alpha, double* out3, double* var_q_out, double* mean_p_out,
                                  double* var_p_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(out3 && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  StepParams P{};
  int rc = gvi_fill_step_params(P, factor_q, factor_p, M, D, X, y, m_q, prior_mean_p, m_p_in, v_p_in, N, reg_kind,
                                renyi_alpha, var_q_out, mean_p_out, var_p_out);
  if (rc) return rc;
  P.tile_counter = (int*)workspace;
  P.partials = (double*)workspace + STEP_COUNTER_DOUBLES;
  P.tile_partials = P.partials + fused_records_doubles();
  P.acc_scratch = P.tile_partials + fused_tile_doubles(N);
  if (D > 32) {
    // wide inputs: the K rows of q and p are built chunk by chunk by the tensor-pipe Gram kernel into two rings that stay
    // L2 resident; the tile kernel consumes both and runs the risk epilogue - still no N x M array in HBM and one ABI
    // call, with 2-3 launches per chunk of <= 28 416 points (workspace: gvi_fused_step_workspace_bytes_d)
    const FactorLayout f = factor_layout(M, D);
    const int64_t rows = wide_chunk_rows(f.Mp, N);  // a multiple of 24: the chunks' tile records are contiguous
    double* ring_q = P.acc_scratch + gvi_step_scratch_doubles(f.Mp);
    double* ring_p = ring_q + (size_t)rows * f.Mp;
    double* gws = ring_p + (size_t)rows * f.Mp;
    const StepParams P0 = P;
    for (int64_t r0 = 0; r0 < N; r0 += rows) {
      const int64_t nr = N - r0 < rows ? N - r0 : rows;
      StepParams C = P0;
      auto off = [r0](const double* p) { return p ? p + r0 : nullptr; };
      auto offw = [r0](double* p) { return p ? p + r0 : nullptr; };
      for (int ps = 0; ps < C.npass; ps++) {
        PassDesc& pd = C.pass[ps];
        double* ring = ps == 0 ? ring_q : ring_p;
        rc = gvi_launch_wide_gram(pd.F, f, X + r0 * D, nr, ring, gws, st);
        if (rc) return rc;
        pd.prior_mean = off(pd.prior_mean);
        pd.mean_out = offw(pd.mean_out);
        pd.var_out = offw(pd.var_out);
        pd.Kpre = ring;
        pd.ldk = f.Mp;
      }
      C.X = nullptr; C.N = nr;
      C.y = off(C.y); C.m_q = off(C.m_q); C.m_p_in = off(C.m_p_in); C.v_p_in = off(C.v_p_in);
      C.tile_partials = P0.tile_partials + (r0 / 24) * 2;
      C.tile_points = 24;  // also for a short last chunk: one record per 24 points throughout
      int g = gvi_launch_step(C, st);
      if (g < 0) return g;
    }
    int nrec_w = gvi_launch_tile_reduce(P.tile_partials, (N + 23) / 24, 0, P.partials, st);
    if (nrec_w < 0) return nrec_w;
    return gvi_launch_finish(P.partials, nrec_w, N, 0, D, (const double*)factor_q, out3, st);
  }
  int grid = gvi_launch_step(P, st);
  if (grid < 0) return grid;
  // forward launches use 24-point tiles; per-tile records -> fixed-order reduction -> out3
  const int T = gvi_step_forward_tile_points(N);
  int nrec = gvi_launch_tile_reduce(P.tile_partials, (N + T - 1) / T, 0, P.partials, st);
  if (nrec < 0) return nrec;
  return gvi_launch_finish(P.partials, nrec, N, 0, D, (const double*)factor_q, out3, st);
}
