// select.cu - greedy conditional-variance inducing-point selection as a pivoted Cholesky (a13).
// Reference: src/inducing_points_selection.py:113-176.  Exact rules kept: d0 = diag + jitter, first pivot = argmax
// with the LOWEST index among ties; every later pivot = first entry of reversed(stable argsort(d)) that is not
// already chosen, i.e. the max d with the HIGHEST index among ties; g = np.round(k(X,x_j), 20) (multiply by 1e20,
// rint, divide); jitter added to g[j] only; d clipped at 0.  The C matrix [(M-1), N] stays in HBM (the reference
// keeps it on the host and re-uploads C[:m] every step); per step every point reads its column of C once:
// HBM-bound, 8*m*N bytes per step.  Sharded use: each rank runs update on its slice and ranks exchange one
// fixed-size candidate record per step (see gvigp.h).
#include "common.cuh"

namespace {
constexpr int STHREADS = 256;
constexpr int64_t FUSED_FINALIZE_MAX_POINTS = 500000;  // single-shard selector: one launch per pivot up to this N

// optional tail of select_update_kernel: the LAST CTA to finish reduces the per-block candidates, gathers the record of the
// next pivot and (single shard) makes it the next winner - one launch per pivot instead of three
struct FusedFinalize {
  unsigned int* ticket;  // null = not fused (select_finalize_kernel / select_resolve_kernel follow)
  int M_sel;
  double* record;
  double* trace_out;     // tr(K - Q) after this step, or null
  double* winner_next;   // where the next step reads its winner (single shard), or null
  int64_t* idx_out_next; // pivot position of the next step, or null
};

struct SelectWs {
  double* d;
  double* C;
  unsigned char* chosen;
  double* blockcand;   // [nblk][2]
  double* blocktrace;  // [nblk]
  double* record;
  double* winner;
  unsigned int* ticket;  // "last CTA done" counter of the fused update (zero between launches)
  int64_t nblk;
};

__host__ inline size_t align16(size_t x) { return (x + 15) / 16 * 16; }

__host__ SelectWs carve(void* ws, int64_t N, int M_sel, int D) {
  SelectWs s;
  unsigned char* p = (unsigned char*)ws;
  s.nblk = (N + STHREADS - 1) / STHREADS;
  if (s.nblk < 1) s.nblk = 1;
  int64_t rl = 2 + D + (M_sel - 1);
  s.d = (double*)p;            p += align16((size_t)N * 8);
  s.C = (double*)p;            p += align16((size_t)N * (size_t)(M_sel - 1) * 8);
  s.chosen = p;                p += align16((size_t)N);
  s.blockcand = (double*)p;    p += align16((size_t)s.nblk * 16);
  s.blocktrace = (double*)p;   p += align16((size_t)s.nblk * 8);
  s.record = (double*)p;       p += align16((size_t)rl * 8);
  s.winner = (double*)p;       p += align16((size_t)rl * 8);
  s.ticket = (unsigned int*)p; p += 16;
  return s;
}

__device__ __forceinline__ bool better(double v, double idx, double bv, double bidx, bool prefer_high) {
  return (v > bv) || (v == bv && (prefer_high ? idx > bidx : idx < bidx));
}

// block-wide (value, index) arg-max with an explicit tie rule; result valid in thread 0
__device__ __forceinline__ void block_argmax(double& v, double& idx, bool prefer_high, double* sv, double* si) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    double oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (better(ov, oi, v, idx, prefer_high)) { v = ov; idx = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = v; si[warp] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++)
      if (better(sv[w], si[w], v, idx, prefer_high)) { v = sv[w]; idx = si[w]; }
  }
}

__global__ void __launch_bounds__(STHREADS) select_init_kernel(int64_t N, int64_t offset,
                                                               const double* __restrict__ log_scaling,
                                                               const double* __restrict__ kdiag, double jitter,
                                                               double* __restrict__ d, unsigned char* __restrict__ chosen,
                                                               double* __restrict__ blockcand,
                                                               double* __restrict__ blocktrace,
                                                               unsigned int* __restrict__ ticket) {
  __shared__ double sv[STHREADS / 32], si[STHREADS / 32];
  const int64_t i = (int64_t)blockIdx.x * STHREADS + threadIdx.x;
  const double amp = kdiag ? 0.0 : exp(log_scaling[0]);
  double v = -1.0, idx = -1.0;
  if (i < N) {
    // k(x,x) + jitter, inducing_points_selection.py:124-132 (kdiag: any kernel, supplied by the caller)
    v = (kdiag ? kdiag[i] : amp * amp) + jitter;
    d[i] = v;
    chosen[i] = 0;
    idx = (double)(offset + i);
  }
  block_argmax(v, idx, /*prefer_high=*/false, sv, si);
  if (threadIdx.x == 0) {
    blockcand[2 * blockIdx.x] = v;
    blockcand[2 * blockIdx.x + 1] = idx;
    blocktrace[blockIdx.x] = 0.0;
    if (blockIdx.x == 0 && ticket) *ticket = 0u;
  }
}

// step m for the local slice: e = (round20(k(X,x_j)) [+jitter at j] - c_j . C[:m]) / sqrt(d_j); C[m] = e;
// d = max(d - e^2, 0); per-block candidate for the next pivot and per-block partial of tr(K - Q).
// `winner` carries no __restrict__: in the fused form the last CTA's tail writes the NEXT winner into the same buffer
// (fin.winner_next), after every CTA has consumed the current one and taken its ticket.
__global__ void __launch_bounds__(STHREADS) select_update_kernel(
    const double* __restrict__ X, int64_t N, int64_t offset, int D, int m, const double* winner,
    const double* __restrict__ log_scaling, const double* __restrict__ log_ls, int ls_len,
    const double* __restrict__ kcol, double jitter, double* __restrict__ d, double* C,
    unsigned char* __restrict__ chosen, double* __restrict__ blockcand, double* __restrict__ blocktrace,
    const FusedFinalize fin) {
  extern __shared__ double sm[];
  double* cj = sm;          // [m]
  double* xj = sm + m;      // [D]
  double* w = xj + D;       // [D]
  __shared__ double sv[STHREADS / 32], si[STHREADS / 32], st[STHREADS / 32];
  const int tid = threadIdx.x;
  for (int k = tid; k < m; k += STHREADS) cj[k] = winner[2 + D + k];
  if (!kcol)
    for (int k = tid; k < D; k += STHREADS) {
      xj[k] = winner[2 + k];
      w[k] = exp(log_ls[ls_len == 1 ? 0 : k]);
    }
  __syncthreads();
  const double dj = sqrt(winner[0]);
  const double jglob = winner[1];
  const double amp = kcol ? 0.0 : exp(log_scaling[0]);
  const double amp2 = amp * amp;
  const int64_t i = (int64_t)blockIdx.x * STHREADS + tid;
  double v = -1.0, idx = -1.0, tr = 0.0;
  if (i < N) {
    double g;
    if (kcol) {
      g = kcol[i];  // k(x_i, x_j) from the caller's kernel
    } else {
      double s = 0.0;
      for (int k = 0; k < D; k++) {
        double diff = X[i * D + k] - xj[k];
        s = fma(w[k], diff * diff, s);
      }
      g = amp2 * exp(-0.5 * s);
    }
    g = __ddiv_rn(rint(__dmul_rn(g, 1e20)), 1e20);  // np.round(g, 20)
    const bool is_pivot = ((double)(offset + i) == jglob);
    if (is_pivot) g += jitter;
    double dot = 0.0;
    double* Ci = C + i;
    int k = 0;
    for (; k + 8 <= m; k += 8) {
      double c0 = Ci[(int64_t)(k + 0) * N], c1 = Ci[(int64_t)(k + 1) * N], c2 = Ci[(int64_t)(k + 2) * N],
             c3 = Ci[(int64_t)(k + 3) * N], c4 = Ci[(int64_t)(k + 4) * N], c5 = Ci[(int64_t)(k + 5) * N],
             c6 = Ci[(int64_t)(k + 6) * N], c7 = Ci[(int64_t)(k + 7) * N];
      dot = fma(cj[k + 0], c0, dot); dot = fma(cj[k + 1], c1, dot); dot = fma(cj[k + 2], c2, dot);
      dot = fma(cj[k + 3], c3, dot); dot = fma(cj[k + 4], c4, dot); dot = fma(cj[k + 5], c5, dot);
      dot = fma(cj[k + 6], c6, dot); dot = fma(cj[k + 7], c7, dot);
    }
    for (; k < m; k++) dot = fma(cj[k], Ci[(int64_t)k * N], dot);
    const double e = (g - dot) / dj;
    Ci[(int64_t)m * N] = e;
    double dn = d[i] - e * e;
    dn = dn > 0.0 ? dn : 0.0;  // jnp.clip(di, 0, None)
    d[i] = dn;
    tr = dn;
    if (is_pivot) chosen[i] = 1;
    if (!chosen[i]) {
      v = dn;
      idx = (double)(offset + i);
    }
  }
  tr = warp_sum(tr);
  if ((tid & 31) == 0) st[tid >> 5] = tr;
  block_argmax(v, idx, /*prefer_high=*/true, sv, si);  // contains a __syncthreads
  if (tid == 0) {
    double t = 0.0;
    for (int q = 0; q < STHREADS / 32; q++) t += st[q];
    blocktrace[blockIdx.x] = t;
    blockcand[2 * blockIdx.x] = v;
    blockcand[2 * blockIdx.x + 1] = idx;
  }
  if (!fin.ticket) return;
  // ---- fused finalize: every thread publishes its C / d writes, the CTA takes a ticket, the last one carries on --------
  __shared__ bool is_last;
  __shared__ double win_v, win_i;
  __syncthreads();  // the CTA's C / d / candidate writes happen-before thread 0's fence (cumulative at device scope)
  if (tid == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(fin.ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *fin.ticket = 0u;  // ready for the next launch
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int64_t nblk = gridDim.x;
  v = -1.0;
  idx = -1.0;
  tr = 0.0;
  for (int64_t b = tid; b < nblk; b += STHREADS) {
    const double bv = __ldcg(blockcand + 2 * b), bi = __ldcg(blockcand + 2 * b + 1);
    if (better(bv, bi, v, idx, true)) { v = bv; idx = bi; }
    tr += __ldcg(blocktrace + b);
  }
  tr = warp_sum(tr);
  if ((tid & 31) == 0) st[tid >> 5] = tr;
  block_argmax(v, idx, /*prefer_high=*/true, sv, si);  // contains a __syncthreads
  if (tid == 0) {
    win_v = v;
    win_i = idx;
    double t = 0.0;
    for (int q = 0; q < STHREADS / 32; q++) t += st[q];
    if (fin.trace_out) *fin.trace_out = t;
    if (fin.idx_out_next) *fin.idx_out_next = (int64_t)idx;
  }
  __syncthreads();
  const double wv = win_v, wi = win_i;
  const int rl = 2 + D + (fin.M_sel - 1);
  double* dst2 = fin.winner_next;
  if (tid == 0) {
    fin.record[0] = wv;
    fin.record[1] = wi;
    if (dst2) { dst2[0] = wv; dst2[1] = wi; }
  }
  if (wi < 0.0) {  // no eligible local point
    for (int k = 2 + tid; k < rl; k += STHREADS) {
      fin.record[k] = 0.0;
      if (dst2) dst2[k] = 0.0;
    }
    return;
  }
  const int64_t il = (int64_t)wi - offset;
  for (int k = tid; k < D; k += STHREADS) {
    const double xv = X[il * D + k];
    fin.record[2 + k] = xv;
    if (dst2) dst2[2 + k] = xv;
  }
  for (int k = tid; k < fin.M_sel - 1; k += STHREADS) {
    const double cv = (k <= m) ? __ldcg(C + (int64_t)k * N + il) : 0.0;
    fin.record[2 + D + k] = cv;
    if (dst2) dst2[2 + D + k] = cv;
  }
}

// one CTA: reduce the per-block candidates, then gather the local record (d, idx, x_j, c_j[0..mrows))
__global__ void __launch_bounds__(1024) select_finalize_kernel(
    const double* __restrict__ X, int64_t N, int64_t offset, int D, int M_sel, int mrows, int prefer_high,
    const double* __restrict__ blockcand, const double* __restrict__ blocktrace, int64_t nblk,
    const double* __restrict__ C, double* __restrict__ record, double* __restrict__ trace_out,
    int64_t* __restrict__ idx_out) {
  __shared__ double sv[32], si[32], stv[32];
  __shared__ double win_v, win_i;
  const int tid = threadIdx.x;
  double v = -1.0, idx = -1.0, tr = 0.0;
  for (int64_t b = tid; b < nblk; b += 1024) {
    double bv = blockcand[2 * b], bi = blockcand[2 * b + 1];
    if (better(bv, bi, v, idx, prefer_high)) { v = bv; idx = bi; }
    tr += blocktrace[b];
  }
  tr = warp_sum(tr);
  if ((tid & 31) == 0) stv[tid >> 5] = tr;
  block_argmax(v, idx, prefer_high, sv, si);
  if (tid == 0) {
    win_v = v;
    win_i = idx;
    double t = 0.0;
    for (int q = 0; q < 32; q++) t += stv[q];
    if (trace_out) *trace_out = t;
    if (idx_out) *idx_out = (int64_t)idx;
  }
  __syncthreads();
  const double wv = win_v, wi = win_i;
  const int rl = 2 + D + (M_sel - 1);
  if (tid == 0) { record[0] = wv; record[1] = wi; }
  if (wi < 0.0) {  // no eligible local point
    for (int k = 2 + tid; k < rl; k += 1024) record[k] = 0.0;
    return;
  }
  const int64_t il = (int64_t)wi - offset;
  for (int k = tid; k < D; k += 1024) record[2 + k] = X[il * D + k];
  for (int k = tid; k < M_sel - 1; k += 1024) record[2 + D + k] = (k < mrows) ? C[(int64_t)k * N + il] : 0.0;
}

// pick the winner among R gathered records: first pivot -> max d, lowest index; later -> max d, highest index
__global__ void __launch_bounds__(256) select_resolve_kernel(const double* __restrict__ records, int R, int rl,
                                                             int first, double* __restrict__ winner,
                                                             int64_t* __restrict__ idx_out) {
  __shared__ int best_s;
  if (threadIdx.x == 0) {
    int best = 0;
    for (int r = 1; r < R; r++)
      if (better(records[(int64_t)r * rl], records[(int64_t)r * rl + 1], records[(int64_t)best * rl],
                 records[(int64_t)best * rl + 1], !first))
        best = r;
    best_s = best;
    if (idx_out) *idx_out = (int64_t)records[(int64_t)best * rl + 1];
  }
  __syncthreads();
  const double* src = records + (int64_t)best_s * rl;
  for (int k = threadIdx.x; k < rl; k += 256) winner[k] = src[k];
}
}  // namespace

extern "C" int64_t gvi_select_record_len(int D, int M_sel) { return 2 + (int64_t)D + (M_sel - 1); }

extern "C" size_t gvi_select_workspace_bytes(int64_t N, int M_sel, int D) {
  if (N < 1 || M_sel < 2 || D < 1) return 0;
  int64_t nblk = (N + STHREADS - 1) / STHREADS;
  int64_t rl = 2 + D + (M_sel - 1);
  return align16((size_t)N * 8) + align16((size_t)N * (size_t)(M_sel - 1) * 8) + align16((size_t)N) +
         align16((size_t)nblk * 16) + align16((size_t)nblk * 8) + 2 * align16((size_t)rl * 8) + 16 + 64;
}

static int select_init_impl(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel,
                            const double* log_scaling, const double* kdiag, double jitter, double* record_out,
                            void* workspace, void* stream) {
  GVI_CHECK_ARG(N_local >= 1 && D >= 1, "sizes");
  GVI_CHECK_ARG(M_sel > 1, "Must have at least 2 inducing points");
  GVI_CHECK_ARG(x_local && (log_scaling || kdiag) && record_out && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  SelectWs s = carve(workspace, N_local, M_sel, D);
  select_init_kernel<<<(unsigned)s.nblk, STHREADS, 0, st>>>(N_local, offset, log_scaling, kdiag, jitter, s.d, s.chosen,
                                                            s.blockcand, s.blocktrace, s.ticket);
  GVI_CHECK_LAUNCH("select_init_kernel");
  select_finalize_kernel<<<1, 1024, 0, st>>>(x_local, N_local, offset, D, M_sel, 0, /*prefer_high=*/0, s.blockcand,
                                             s.blocktrace, s.nblk, s.C, record_out, nullptr, nullptr);
  GVI_CHECK_LAUNCH("select_finalize_kernel");
  return GVI_OK;
}

extern "C" int gvi_select_init_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel,
                                   const double* log_scaling, double jitter, double* record_out, void* workspace,
                                   void* stream) {
  GVI_CHECK_ARG(log_scaling, "null pointer");
  return select_init_impl(x_local, N_local, offset, D, M_sel, log_scaling, nullptr, jitter, record_out, workspace,
                          stream);
}
// any kernel: k(x_i, x_i) supplied by the caller
extern "C" int gvi_select_init_gram_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel,
                                        const double* kdiag, double jitter, double* record_out, void* workspace,
                                        void* stream) {
  GVI_CHECK_ARG(kdiag, "null pointer");
  return select_init_impl(x_local, N_local, offset, D, M_sel, nullptr, kdiag, jitter, record_out, workspace, stream);
}

extern "C" int gvi_select_resolve_f64(const double* records, int R, int D, int M_sel, int first, double* winner_out,
                                      int64_t* idx_out_m, void* stream) {
  GVI_CHECK_ARG(records && winner_out && R >= 1, "records");
  int rl = (int)gvi_select_record_len(D, M_sel);
  select_resolve_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(records, R, rl, first, winner_out, idx_out_m);
  GVI_CHECK_LAUNCH("select_resolve_kernel");
  return GVI_OK;
}

static int select_update_impl(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel, int m,
                              const double* winner, const double* log_scaling, const double* log_ls, int ls_len,
                              const double* kcol, double jitter, double* record_out, double* trace_partial_out,
                              void* workspace, void* stream, int64_t* fused_idx_out_next = nullptr,
                              double* fused_winner_next = nullptr, bool fused = false);

extern "C" int gvi_select_update_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel, int m,
                                     const double* winner, const double* log_scaling, const double* log_ls,
                                     int ls_len, double jitter, double* record_out, double* trace_partial_out,
                                     void* workspace, void* stream) {
  GVI_CHECK_ARG(log_scaling && log_ls, "null pointer");
  GVI_CHECK_ARG(ls_len == 1 || ls_len == D, "log_lengthscales must have 1 or D entries");
  return select_update_impl(x_local, N_local, offset, D, M_sel, m, winner, log_scaling, log_ls, ls_len, nullptr, jitter,
                            record_out, trace_partial_out, workspace, stream);
}
// any kernel: kcol[i] = k(x_i, x_j) for the winning pivot j, supplied by the caller
extern "C" int gvi_select_update_gram_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel,
                                          int m, const double* winner, const double* kcol, double jitter,
                                          double* record_out, double* trace_partial_out, void* workspace,
                                          void* stream) {
  GVI_CHECK_ARG(kcol, "null pointer");
  return select_update_impl(x_local, N_local, offset, D, M_sel, m, winner, nullptr, nullptr, 1, kcol, jitter,
                            record_out, trace_partial_out, workspace, stream);
}

static int select_update_impl(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel, int m,
                              const double* winner, const double* log_scaling, const double* log_ls, int ls_len,
                              const double* kcol, double jitter, double* record_out, double* trace_partial_out,
                              void* workspace, void* stream, int64_t* fused_idx_out_next, double* fused_winner_next,
                              bool fused) {
  GVI_CHECK_ARG(N_local >= 1 && D >= 1 && M_sel > 1 && m >= 0 && m < M_sel - 1, "sizes");
  GVI_CHECK_ARG(x_local && winner && record_out && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  SelectWs s = carve(workspace, N_local, M_sel, D);
  size_t smem = (size_t)(m + 2 * D) * sizeof(double);
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      GVI_CUDA(cudaFuncSetAttribute(select_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
  }
  FusedFinalize fin{};
  if (fused) fin = FusedFinalize{s.ticket, M_sel, record_out, trace_partial_out, fused_winner_next, fused_idx_out_next};
  select_update_kernel<<<(unsigned)s.nblk, STHREADS, smem, st>>>(x_local, N_local, offset, D, m, winner, log_scaling,
                                                                 log_ls, ls_len, kcol, jitter, s.d, s.C, s.chosen,
                                                                 s.blockcand, s.blocktrace, fin);
  GVI_CHECK_LAUNCH("select_update_kernel");
  if (fused) return GVI_OK;  // the last CTA of the launch did the reduction, the record and (single shard) the winner
  select_finalize_kernel<<<1, 1024, 0, st>>>(x_local, N_local, offset, D, M_sel, m + 1, /*prefer_high=*/1,
                                             s.blockcand, s.blocktrace, s.nblk, s.C, record_out, trace_partial_out,
                                             nullptr);
  GVI_CHECK_LAUNCH("select_finalize_kernel");
  return GVI_OK;
}

extern "C" int gvi_cond_var_select_f64(const double* x_perm, int64_t N, int D, int M_sel, const double* log_scaling,
                                       const double* log_ls, int ls_len, double jitter, int64_t* idx_out,
                                       double* trace_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(M_sel > 1, "Must have at least 2 inducing points");
  GVI_CHECK_ARG(N >= 1 && D >= 1, "sizes");
  GVI_CHECK_ARG(x_perm && log_scaling && log_ls && idx_out && workspace, "null pointer");
  GVI_CHECK_ARG(ls_len == 1 || ls_len == D, "log_lengthscales must have 1 or D entries");
  SelectWs s = carve(workspace, N, M_sel, D);
  int rc = gvi_select_init_f64(x_perm, N, 0, D, M_sel, log_scaling, jitter, s.record, workspace, stream);
  if (rc) return rc;
  // single shard: the local record is the winner.
  if (N > FUSED_FINALIZE_MAX_POINTS) {
    // large N: three launches per pivot (update, finalize, resolve).  The fused tail below costs every CTA a fence and a
    // ticket round trip (~2 us of its lifetime, ~2 % of the step at N = 1e6: 175.7 vs 172.5 ms for M = 512), more than the
    // two small launches it saves once a step runs for several waves of CTAs.
    for (int m = 0; m < M_sel - 1; m++) {
      if ((rc = gvi_select_resolve_f64(s.record, 1, D, M_sel, m == 0, s.winner, idx_out + m, stream))) return rc;
      if ((rc = gvi_select_update_f64(x_perm, N, 0, D, M_sel, m, s.winner, log_scaling, log_ls, ls_len, jitter,
                                      s.record, trace_out ? trace_out + m : nullptr, workspace, stream)))
        return rc;
    }
    return gvi_select_resolve_f64(s.record, 1, D, M_sel, 0, s.winner, idx_out + (M_sel - 1), stream);
  }
  // small N (the step is latency bound: N = 1e5, M = 512 went 25.2 -> 22.5 ms): ONE launch per pivot - the last CTA of
  // update m reduces the per-block candidates, gathers the record of pivot m+1, copies it to `winner`, writes idx_out[m+1].
  if ((rc = gvi_select_resolve_f64(s.record, 1, D, M_sel, 1, s.winner, idx_out, stream))) return rc;
  for (int m = 0; m < M_sel - 1; m++) {
    if ((rc = select_update_impl(x_perm, N, 0, D, M_sel, m, s.winner, log_scaling, log_ls, ls_len, nullptr, jitter,
                                 s.record, trace_out ? trace_out + m : nullptr, workspace, stream, idx_out + m + 1,
                                 s.winner, true)))
      return rc;
  }
  return GVI_OK;
}
