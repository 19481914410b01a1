// predict.cu - the hot kernel: per tile of T points, build K(x,Z) in shared memory (never in HBM), multiply by the
// packed L^-1 on the FP64 tensor pipe (DMMA m8n8k4), and reduce the squared row norms -> predictive variance;
// the same K tile gives the mean (K alpha).  Up to two GPs (q = approximate GP, p = exact regulariser GP) are
// processed per tile and the Gaussian NLL + pGVI/GWI terms are reduced in the epilogue, so one launch is the whole
// "Gram + predict + risk" step and only O(N) bytes cross HBM.
//
// Reference arithmetic: sparse_posterior_kernel.py:63-96 in diagonal mode (src/kernels/base.py:132-147),
// src/gps/base/base.py:495-526 (exact posterior), negative_log_likelihood.py:41-55, regularisations/projected/*.
// Algebra: v = k(x,x) - k^T A^-1 k = k(x,x) - ||L^-1 k||^2 with A = L L^T;  b = L^-1 k is a dependency-free
// triangular GEMM against the explicit inverse (validated against cho_solve in tests; see DESIGN.md numerics).
#include "common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

struct PassDesc {
  const double* F;          // factor buffer
  const double* prior_mean; // [N] or null, added to K alpha
  const double* noise_add;  // device scalar log-noise or null: var += exp(*noise_add)
  double* mean_out;         // [N] or null
  double* var_out;          // [N] or null
};

struct StepParams {
  PassDesc pass[2];
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
  double* partials;     // [gridDim.x][2]
};

template <int NT, int DR>
__global__ void __launch_bounds__(THREADS, 1) gp_step_kernel(const StepParams P) {
  constexpr int T = 8 * NT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ksm = reinterpret_cast<double*>(smem_raw);            // [Mp*T] fragment order
  double* xs = Ksm + (size_t)P.Mp * T;                            // [T][DR] scaled x tile
  double* redm = xs + T * DR;                                     // [WARPS][T]
  double* reds = redm + WARPS * T;                                // [G][T] squared norms per 32-row group
  double* res = reds + (size_t)P.G * T;                           // [2 passes][T][2] (mean, var)
  int* counter = reinterpret_cast<int*>(res + 2 * T * 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ntiles = (P.N + T - 1) / T;
  double part_nll = 0.0, part_reg = 0.0;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * T;
    for (int ps = 0; ps < P.npass; ps++) {
      const PassDesc& pd = P.pass[ps];
      const FactorLayout f = factor_layout(P.M, P.D);
      const double* __restrict__ F = pd.F;
      const double amp2 = F[0];
      // ---- stage the scaled x tile -------------------------------------------------------------------
      for (int idx = tid; idx < T * DR; idx += THREADS) {
        int i = idx / DR, d = idx % DR;
        double v = 0.0;
        if (d < P.D && row0 + i < P.N) v = P.X[(row0 + i) * P.D + d] * F[f.sqrtw_off + d];
        xs[idx] = v;
      }
      if (tid == 0) *counter = 0;
      __syncthreads();
      // ---- phase 1: K tile (T x Mp) into shared memory in B-fragment order; mean partials ------------------
      {
        double macc[T];
#pragma unroll
        for (int i = 0; i < T; i++) macc[i] = 0.0;
        const double* __restrict__ Zs = F + f.zs_off;
        const double* __restrict__ alpha = F + f.alpha_off;
        for (int j = tid; j < P.Mp; j += THREADS) {
          double z[DR];
#pragma unroll
          for (int d = 0; d < DR; d++) z[d] = (d < P.D) ? Zs[(int64_t)j * P.D + d] : 0.0;
          const double aj = alpha[j];
          const double scale = (j < P.M) ? amp2 : 0.0;
          double* kcol = Ksm + (size_t)(j >> 3) * (NT * 64) + ((j >> 2) & 1) + 2 * (j & 3);
#pragma unroll
          for (int i = 0; i < T; i++) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < DR; d++) {
              double diff = xs[i * DR + d] - z[d];
              s = fma(diff, diff, s);
            }
            double k = scale * exp(-0.5 * s);
            kcol[(i >> 3) * 64 + (i & 7) * 8] = k;
            macc[i] = fma(k, aj, macc[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < T; i++) {
          double v = warp_sum(macc[i]);
          if (lane == 0) redm[warp * T + i] = v;
        }
      }
      __syncthreads();
      // ---- phase 2: b = L^-1 k on the tensor pipe, accumulate ||b||^2 ----------------------------------
      {
        const double2* __restrict__ Ks2 = reinterpret_cast<const double2*>(Ksm);
        const double* __restrict__ W = F + f.wpack_off;
        while (true) {
          int gi = 0;
          if (lane == 0) gi = atomicAdd(counter, 1);
          gi = __shfl_sync(0xffffffffu, gi, 0);
          if (gi >= P.G) break;
          const int g = P.G - 1 - gi;  // largest groups first
          const double2* __restrict__ Ag = reinterpret_cast<const double2*>(W + wpack_group_off(g)) + lane;
          const int nkp = 4 * (g + 1);
          double acc[4][NT][2];
#pragma unroll
          for (int mt = 0; mt < 4; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
          double2 a_cur[4], a_nxt[4];
#pragma unroll
          for (int mt = 0; mt < 4; mt++) a_cur[mt] = __ldcg(Ag + mt * 32);
#pragma unroll 1
          for (int kp = 0; kp < nkp; kp++) {
            const int kn = (kp + 1 < nkp) ? kp + 1 : kp;
#pragma unroll
            for (int mt = 0; mt < 4; mt++) a_nxt[mt] = __ldcg(Ag + (kn * 4 + mt) * 32);
            double2 b[NT];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) b[nt] = Ks2[(kp * NT + nt) * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
              for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a_cur[mt].x, b[nt].x);
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
              for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a_cur[mt].y, b[nt].y);
#pragma unroll
            for (int mt = 0; mt < 4; mt++) a_cur[mt] = a_nxt[mt];
          }
          // squared norms of this group's 32 rows: lanes with equal lane%4 hold the same two points of every
          // n-tile.  Stored per GROUP (not per warp) so the final sum has a fixed order whatever warp ran the group.
#pragma unroll
          for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              double v = 0.0;
#pragma unroll
              for (int mt = 0; mt < 4; mt++) v = fma(acc[mt][nt][e], acc[mt][nt][e], v);
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              if (lane < 4) reds[g * T + nt * 8 + lane * 2 + e] = v;
            }
        }
      }
      __syncthreads();
      // ---- tile epilogue: variance and mean of this pass --------------------------------------------
      if (tid < T) {
        double q = 0.0, mu = 0.0;
        for (int g = 0; g < P.G; g++) q += reds[g * T + tid];
#pragma unroll
        for (int w = 0; w < WARPS; w++) mu += redm[w * T + tid];
        double var = F[1] - q;
        if (pd.noise_add) var += exp(pd.noise_add[0]);
        const int64_t r = row0 + tid;
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
    // ---- fused risk epilogue -------------------------------------------------------------------------
    if (P.partials && tid < T) {
      const int64_t r = row0 + tid;
      if (r < P.N) {
        const double vq = res[(0 * T + tid) * 2 + 1];
        const double mq = P.m_q ? P.m_q[r] : res[(0 * T + tid) * 2 + 0];
        if (P.y) part_nll += gvi_nll_term(P.y[r], mq, vq);
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
        }
      }
    }
  }  // tiles
  if (P.partials && warp == 0) {
    // T <= 24 < 32: threads >= T hold zeros
    double a = warp_sum(part_nll), b = warp_sum(part_reg);
    if (lane == 0) {
      P.partials[2 * blockIdx.x + 0] = a;
      P.partials[2 * blockIdx.x + 1] = b;
    }
  }
}

// fixed-order final reduction of per-CTA partial sums -> out3 = {sum nll, sum reg, N}
__global__ void __launch_bounds__(256) finish_partials_kernel(const double* __restrict__ partials, int nblocks,
                                                              int64_t N, double* __restrict__ out3) {
  __shared__ double sa[256], sb[256];
  const int tid = threadIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = tid; i < nblocks; i += 256) {
    a += partials[2 * i];
    b += partials[2 * i + 1];
  }
  sa[tid] = a;
  sb[tid] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) {
      sa[tid] += sa[tid + s];
      sb[tid] += sb[tid + s];
    }
    __syncthreads();
  }
  if (tid == 0) {
    out3[0] = sa[0];
    out3[1] = sb[0];
    out3[2] = (double)N;
  }
}

template <int NT, int DR>
size_t step_smem_bytes(int Mp) {
  constexpr int T = 8 * NT;
  return ((size_t)Mp * T + T * DR + WARPS * T + (size_t)(Mp / 32) * T + 4 * T) * sizeof(double) + 16;
}

constexpr size_t SMEM_LIMIT = 227 * 1024;

template <int NT, int DR>
int launch_step_inst(const StepParams& P, int grid_cap, cudaStream_t st) {
  constexpr int T = 8 * NT;
  size_t smem = step_smem_bytes<NT, DR>(P.Mp);
  static bool configured = false;  // per instantiation
  if (!configured) {
    GVI_CUDA(cudaFuncSetAttribute(gp_step_kernel<NT, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)SMEM_LIMIT));
    configured = true;
  }
  int64_t ntiles = (P.N + T - 1) / T;
  int grid = (int)(ntiles < grid_cap ? ntiles : grid_cap);
  gp_step_kernel<NT, DR><<<grid, THREADS, smem, st>>>(P);
  GVI_CHECK_LAUNCH("gp_step_kernel");
  return grid;
}

template <int NT>
int launch_step_nt(const StepParams& P, int grid_cap, cudaStream_t st) {
  if (P.D <= 1) return launch_step_inst<NT, 1>(P, grid_cap, st);
  if (P.D <= 4) return launch_step_inst<NT, 4>(P, grid_cap, st);
  if (P.D <= 8) return launch_step_inst<NT, 8>(P, grid_cap, st);
  if (P.D <= 16) return launch_step_inst<NT, 16>(P, grid_cap, st);
  return launch_step_inst<NT, 32>(P, grid_cap, st);
}

// returns grid size (>0) or a negative error code
int launch_step(const StepParams& P, cudaStream_t st) {
  if (P.D > 32) {
    gvi_set_error("fused predict supports D <= 32 in this build (got D=%d)", P.D);
    return GVI_EUNSUPPORTED;
  }
  const int DRmax = 32;
  (void)DRmax;
  const int grid_cap = GVI_NUM_SMS;
  if (step_smem_bytes<3, 32>(P.Mp) <= SMEM_LIMIT) return launch_step_nt<3>(P, grid_cap, st);
  if (step_smem_bytes<2, 32>(P.Mp) <= SMEM_LIMIT) return launch_step_nt<2>(P, grid_cap, st);
  if (step_smem_bytes<1, 32>(P.Mp) <= SMEM_LIMIT) return launch_step_nt<1>(P, grid_cap, st);
  gvi_set_error("fused predict supports M <= 3584 in this build (got M=%d)", P.M);
  return GVI_EUNSUPPORTED;
}
}  // namespace

extern "C" int gvi_gp_predict_f64(const void* factor, int M, int D, const double* X, int64_t N,
                                  const double* prior_mean, const double* log_noise_add, double* mean_out,
                                  double* var_out, void* stream) {
  GVI_CHECK_ARG(M >= 1 && D >= 1 && N >= 0, "sizes");
  if (N == 0) return GVI_OK;
  GVI_CHECK_ARG(factor && X, "null pointer");
  StepParams P{};
  const FactorLayout f = factor_layout(M, D);
  P.pass[0] = PassDesc{(const double*)factor, prior_mean, log_noise_add, mean_out, var_out};
  P.npass = 1;
  P.M = M; P.Mp = f.Mp; P.D = D; P.G = f.G;
  P.X = X; P.N = N;
  P.partials = nullptr;
  int rc = launch_step(P, (cudaStream_t)stream);
  return rc < 0 ? rc : GVI_OK;
}

extern "C" size_t gvi_fused_step_workspace_bytes(int64_t N) {
  (void)N;
  return (size_t)GVI_NUM_SMS * 2 * sizeof(double);
}

extern "C" int gvi_fused_step_f64(const void* factor_q, const void* factor_p, int M, int D, const double* X,
                                  const double* y, const double* m_q, const double* prior_mean_p,
                                  const double* m_p_in, const double* v_p_in, int64_t N, int reg_kind,
                                  double renyi_alpha, double* out3, double* var_q_out, double* mean_p_out,
                                  double* var_p_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(M >= 1 && D >= 1 && N >= 1, "sizes");
  GVI_CHECK_ARG(factor_q && X && m_q && out3 && workspace, "null pointer");
  GVI_CHECK_ARG(reg_kind >= GVI_REG_NONE && reg_kind <= GVI_REG_SQUARED_DIFFERENCE, "reg_kind");
  GVI_CHECK_ARG(reg_kind == GVI_REG_NONE || factor_p || (m_p_in && v_p_in),
                "regulariser needs factor_p or (m_p_in, v_p_in)");
  cudaStream_t st = (cudaStream_t)stream;
  StepParams P{};
  const FactorLayout f = factor_layout(M, D);
  P.pass[0] = PassDesc{(const double*)factor_q, nullptr, nullptr, nullptr, var_q_out};
  P.npass = 1;
  if (factor_p) {
    P.pass[1] = PassDesc{(const double*)factor_p, prior_mean_p, nullptr, mean_p_out, var_p_out};
    P.npass = 2;
  }
  P.M = M; P.Mp = f.Mp; P.D = D; P.G = f.G;
  P.X = X; P.N = N;
  P.y = y; P.m_q = m_q; P.m_p_in = m_p_in; P.v_p_in = v_p_in;
  P.reg_kind = reg_kind; P.renyi_alpha = renyi_alpha;
  P.partials = (double*)workspace;
  int grid = launch_step(P, st);
  if (grid < 0) return grid;
  finish_partials_kernel<<<1, 256, 0, st>>>(P.partials, grid, N, out3);
  GVI_CHECK_LAUNCH("finish_partials_kernel");
  return GVI_OK;
}
