// predict.cuh - parameter block of the fused tile kernel, shared between predict.cu and grad.cu
#pragma once
#include "common.cuh"

struct PassDesc {
  const double* F;          // factor buffer
  const double* prior_mean; // [N] or null, added to K alpha
  const double* noise_add;  // device scalar log-noise or null: var += exp(*noise_add)
  double* mean_out;         // [N] or null
  double* var_out;          // [N] or null
  const double* Kpre;       // optional precomputed K(x, Z) rows [N][ldk] (wide inputs, D > 32: Gram built by the
  int64_t ldk;              //   tensor-pipe Gram kernel chunk by chunk in an L2-sized ring); null = compute in phase 1
  const double* kdiag;      // optional k(x_i, x_i) per point [N] (non-stationary kernels, gvi_gp_predict_gram_f64);
                            //   null = the constant F[1] of a stationary kernel
};

// per-CTA partial record: [0] sum nll, [1] sum reg, [2] sum g v, [3] sum g |a|^2, [4..36) T2_d, [36] sum g
constexpr int STEP_PARTIAL_STRIDE = 38;
constexpr int STEP_REDUCE_BLOCKS = 64;   // stage-A blocks of the fixed-order reduction over per-tile records
constexpr int STEP_COUNTER_DOUBLES = 32; // workspace slot of the tile counter (256 B)

struct StepParams {
  PassDesc pass[2];     // [0] = q (approximate GP), [1] = p (regulariser GP)
  int npass;
  int M, Mp, D, G;
  const double* X;
  int64_t N;
  // fused risk epilogue (enabled when partials != null)
  const double* y;      // [N] or null (no NLL term)
  const double* m_q;    // [N] mean of q (host framework evaluates the mean function)
  const double* m_p_in; // [N] regulariser mean when there is no pass 1 (prior mode)
  const double* v_p_in; // [N]
  int reg_kind;
  double renyi_alpha;
  double* partials;     // [gridDim.x][STEP_PARTIAL_STRIDE] (static tile order) / [STEP_REDUCE_BLOCKS][..] (dynamic)
  // dynamic tile order: CTAs draw tiles from a global counter (an SM's two co-resident CTAs do not progress at the same
  // rate, and SMs differ slightly: with a static stride the launch ends on the slowest CTA, ~10 % after the fastest).
  // The loss / gradient partial sums are then written PER TILE and reduced in a fixed order afterwards, so the result
  // does not depend on which CTA ran which tile (run-to-run bit reproducible).
  int* tile_counter;    // null = static stride; zeroed by the launcher
  double* tile_partials;  // [ntiles][2] (forward: nll, reg) or [ntiles][STEP_PARTIAL_STRIDE] (gradient mode)
  // gradient mode
  int grad;
  double* g_out;        // [N] dR/dv_q
  const double* g_in;   // [N] or null: externally supplied dLoss/dvar (gvi_gp_predict_grad_f64); the risk epilogue is skipped
  double* dm_out;       // [N] dR/dm_q
  double* bscratch;     // [gridDim.x][Mp*T] per-CTA b tile (L2 resident)
  int grad_pointwise_only;  // 1: only g_out / dm_out are needed (SVGP step: no a_i, no phase 3)
  int debug_skip;       // profiling aid (env GVI_DEBUG_SKIP): bit 0 skips phase 1, bit 1 skips phase 2 (results garbage)
  // column panels: the K tile holds Pc columns at a time (Pc == Mp: one panel).  With several panels the accumulators
  // of the groups that extend past the current panel are parked in `acc_scratch` (L2 resident) between sweeps and the
  // per-group squared norms go to `reds_scratch` instead of shared memory.
  int Pc;
  double* acc_scratch;  // [gridDim.x][Mp*T]  (multi-panel only)
  double* reds_scratch; // [gridDim.x][G*T]   (when the per-group norms do not fit in shared memory)
  int reds_in_smem;     // set by the launcher
  int two_cta;          // set by the launcher: 8-warp CTAs, two per SM
  int tile_points;      // forward launches: 0 = chosen from N (gvi_step_forward_tile_points), else 24 / 16 / 8
  int has_extra;        // the q factor carries an extra matrix E (SVGP): single panel required
  // gradient mode: a_i = A^-1 k_i of every tile is also written to a ring in the operand order of the weighted-Gram
  // (SYRK) kernel of grad.cu, together with g_i, so that S = sum_i g_i a_i a_i^T is accumulated from the a_i directly
  // (error ~ cond(A) eps; the earlier S = A^-1 (sum_i g_i k_i k_i^T) A^-1 amplified the rounding of C by cond(A)^2).
  double* a_ring;       // [ncb][a_npg][SYRK_BLK] or null
  int a_npg;            // 32-point groups per column block in the ring (this launch)
};

// ---- operand blocks of the weighted-Gram (SYRK) kernel (grad.cu), also written by the tile kernel -------------------
// One block = the DMMA fragments of 128 columns x 32 points: [kp (4)][tile (16)][lane (32)][2] doubles, each kp part
// padded by 4 doubles (bank spread for the generating warps), followed by the 32 weights g_i of the points.
constexpr int SYRK_CB = 128;  // columns per block = C tile edge
constexpr int SYRK_PC = 32;   // points per block
constexpr int SYRK_KPSTRIDE = 16 * 32 * 2 + 4;
constexpr int SYRK_STAGE = 4 * SYRK_KPSTRIDE;
constexpr int SYRK_BLK = SYRK_STAGE + SYRK_PC;  // doubles per (column block, point group)
// element (column c of the block, point i of the group)
__host__ __device__ __forceinline__ int syrk_frag_index(int c, int i) {
  return (i >> 3) * SYRK_KPSTRIDE + (((c >> 3) * 32) + ((c & 7) * 4 + (i & 3))) * 2 + ((i >> 2) & 1);
}

int gvi_step_tile_points(int Mp, int D, bool grad);
int gvi_step_forward_tile_points(int64_t N);  // 24, or 16 / 8 for launches of a few waves at most
// forward-only launches use column panels when the whole K tile does not fit: returns Pc (== Mp: single panel)
size_t gvi_step_scratch_doubles(int Mp);  // per-launch scratch for the panel path
int gvi_launch_step(const StepParams& P, cudaStream_t st);
int gvi_launch_finish(const double* partials, int nblocks, int64_t N, int grad, int D, const double* Fq, double* out,
                      cudaStream_t st);
// per-tile records -> STEP_REDUCE_BLOCKS records of STEP_PARTIAL_STRIDE (fixed order); returns the record count
int gvi_launch_tile_reduce(const double* tile_partials, int64_t ntiles, int grad, double* records, cudaStream_t st);
int gvi_fill_step_params(StepParams& P, const void* factor_q, const void* factor_p, int M, int D, const double* X,
                         const double* y, const double* m_q, const double* prior_mean_p, const double* m_p_in,
                         const double* v_p_in, int64_t N, int reg_kind, double renyi_alpha, double* var_q_out,
                         double* mean_p_out, double* var_p_out);
