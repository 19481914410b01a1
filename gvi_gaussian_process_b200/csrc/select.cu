// select.cu - greedy conditional-variance inducing-point selection as a pivoted Cholesky (a13).
// Reference: src/inducing_points_selection.py:113-176.  Exact rules kept: d0 = diag + jitter, first pivot = argmax
// with the LOWEST index among ties; every later pivot = first entry of reversed(stable argsort(d)) that is not
// already chosen, i.e. the max d with the HIGHEST index among ties; g = np.round(k(X,x_j), 20) (multiply by 1e20,
// rint, divide); jitter added to g[j] only; d clipped at 0.  The C matrix [(M-1), N] stays in HBM (the reference
// keeps it on the host and re-uploads C[:m] every step); per step every point reads its column of C once:
// HBM-bound, 8*m*N bytes per step.  Sharded use: each rank runs update on its slice and ranks exchange one
// fixed-size candidate record per step (see gvigp.h).
#include "comm.cuh"

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

// step m for ONE point i of a slice: e = (round20(k(x_i,x_j)) [+jitter at j] - c_j . C[:m,i]) / sqrt(d_j); C[m,i] = e;
// d_i = max(d_i - e^2, 0) (inducing_points_selection.py:141-160); (v, idx) = the point's candidacy for the next pivot
__device__ __forceinline__ void update_point(const double* __restrict__ X, int64_t N, int64_t offset, int D, int m,
                                             const double* cj, const double* xj, const double* w, double dj,
                                             double jglob, double amp2, const double* __restrict__ kcol,
                                             double jitter, double* __restrict__ d, double* C,
                                             unsigned char* __restrict__ chosen, int64_t i, double& v, double& idx,
                                             double& tr) {
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
  if (i < N)
    update_point(X, N, offset, D, m, cj, xj, w, dj, jglob, amp2, kcol, jitter, d, C, chosen, i, v, idx, tr);
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
    static std::atomic<unsigned long long> configured{0};
    if (gvi_first_use_on_this_device(configured))
      GVI_CUDA(cudaFuncSetAttribute(select_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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

// =====================================================================================================================
// Sharded selection with the exchange INSIDE the kernels (SURVEY.md section 8(e) "Greedy selector: one exchange step per
// pivot"; reference loop: src/inducing_points_selection.py:137-172).
//
// Every rank owns a contiguous slice of the permuted points and a MAILBOX of 2 slots x R records.  After update step m
// a one-CTA finalize kernel reduces the slice's per-block candidates and writes the slice's candidate record
//   [d, global index, local tr(K-Q), tag, x_j[D], c_j[0..m]]
// straight into slot (m+1)&1 of EVERY rank's mailbox (plain stores on peer-mapped pointers over NVLink / NVSwitch, the
// tag word last, behind a system-scope fence).  The next update kernel starts by waiting for the R tags of its own
// mailbox, resolves the winner with the reference tie rule (every CTA redundantly: R <= 16 comparisons) and carries on.
// No host code, no collective call and no second broadcast sit between two pivots.  Two slots suffice: a rank can only
// send its record for step m+1 after it has received every rank's record for step m, which those ranks sent after they
// had finished reading step m-1.  Tags are signalling-NaN bit patterns carrying a counter that grows over the
// communicator's lifetime: no arithmetic produces them, so stale payload can never be taken for a tag and the windows
// never need clearing.
// Transport fall-back (no peer mapping): the finalize kernel writes the record to a local send buffer and
// ncclAllGather (enqueued from C on the same stream) fills the mailbox slot - same kernels, stream-ordered.
// The same kernels also run R VIRTUAL shards on one device (gvi_cond_var_select_virtual_shards_f64): one launch covers
// all shards' slices, the mailboxes are local - the exchange protocol is exercised bit for bit without a second GPU and
// without kernels that wait on one another (a launch only ever waits on tags written by an EARLIER launch there).
namespace {
constexpr int SEL_HDR = 4;
constexpr int SEL_MAX_LOCAL = GVI_COMM_MAX_RANKS;  // virtual shards per launch
constexpr unsigned long long SEL_TAG_BASE = 0xFFF4000000000000ull;  // sign + all-ones exponent + signalling payload
constexpr long long SEL_WAIT_NS = 20ll * 1000 * 1000 * 1000;         // a peer that never answers: give up, do not hang

struct ShardView {
  const double* X;
  int64_t N, offset;
  double* d;
  double* C;
  unsigned char* chosen;
  double* blockcand;
  double* blocktrace;
  int64_t blk0, nblk;   // this shard's CTAs in the launch: [blk0, blk0 + nblk)
  double* mailbox;      // [2][R][rl_max]
  int64_t* idx_out;     // [M_sel]
  double* trace_out;    // [M_sel - 1] or null
  int* status;          // sticky: != 0 after a wait timed out
};
struct ShardSet {
  int n;
  ShardView s[SEL_MAX_LOCAL];
};
struct TargetSet {  // where a finalize CTA sends its record: n destinations, record position `pos` in each
  int n;
  double* base[GVI_COMM_MAX_RANKS];
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const double* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// wait until the R records of `slot` carry `tag`, then pick the winner (first pivot: max d, LOWEST index - argmax,
// inducing_points_selection.py:133; later: max d, HIGHEST index among the not yet chosen - reversed stable argsort,
// :161-166).  Returns the winning rank in *best_s (shared) or -1 after a time-out; also sums the ranks' tr(K-Q) partials in
// rank order.  Called by all threads of a CTA.
__device__ __forceinline__ int wait_and_resolve(const double* slot, int R, int64_t stride, unsigned long long tag,
                                                bool first, int* status, int* best_s, double* trace_s) {
  if (threadIdx.x == 0) *best_s = 0;
  __syncthreads();
  if ((int)threadIdx.x < R) {
    const double* tagword = slot + (int64_t)threadIdx.x * stride + 3;
    if (ld_volatile_u64(tagword) != tag) {
      const long long t0 = globaltimer_ns();
      int spins = 0;
      while (ld_volatile_u64(tagword) != tag) {
        if (((++spins) & 1023) == 0 &&
            (*(volatile int*)status != 0 || globaltimer_ns() - t0 > SEL_WAIT_NS)) {
          atomicExch(status, 1);
          *best_s = -1;
          break;
        }
      }
    }
    __threadfence();  // the payload reads below are ordered after the tag read
  }
  __syncthreads();
  if (*best_s < 0) return -1;
  if (threadIdx.x == 0) {
    int best = 0;
    double tr = __ldcg(slot + 2);
    for (int r = 1; r < R; r++) {
      const double* rec = slot + (int64_t)r * stride;
      if (better(__ldcg(rec), __ldcg(rec + 1), __ldcg(slot + (int64_t)best * stride),
                 __ldcg(slot + (int64_t)best * stride + 1), !first))
        best = r;
      tr += __ldcg(rec + 2);
    }
    *best_s = best;
    *trace_s = tr;
  }
  __syncthreads();
  return *best_s;
}

__global__ void __launch_bounds__(STHREADS) select_update_sharded_kernel(
    const ShardSet S, int D, int m, int R, int64_t stride, int64_t slot_off, unsigned long long tag,
    const double* __restrict__ log_scaling, const double* __restrict__ log_ls, int ls_len, double jitter) {
  extern __shared__ double sm[];
  double* cj = sm;      // [m]
  double* xj = sm + m;  // [D]
  double* w = xj + D;   // [D]
  __shared__ double sv[STHREADS / 32], si[STHREADS / 32], st[STHREADS / 32];
  __shared__ int best_s;
  __shared__ double trace_s;
  const int tid = threadIdx.x;
  int s = 0;
  while (s + 1 < S.n && (int64_t)blockIdx.x >= S.s[s + 1].blk0) s++;
  const ShardView& sh = S.s[s];
  if (*(volatile int*)sh.status != 0) return;
  const double* slot = sh.mailbox + slot_off;
  const int best = wait_and_resolve(slot, R, stride, tag, m == 0, sh.status, &best_s, &trace_s);
  if (best < 0) return;
  const double* winner = slot + (int64_t)best * stride;
  for (int k = tid; k < m; k += STHREADS) cj[k] = __ldcg(winner + SEL_HDR + D + k);
  for (int k = tid; k < D; k += STHREADS) {
    xj[k] = __ldcg(winner + SEL_HDR + k);
    w[k] = exp(log_ls[ls_len == 1 ? 0 : k]);
  }
  __syncthreads();
  const double dj = sqrt(__ldcg(winner));
  const double jglob = __ldcg(winner + 1);
  const int64_t lb = (int64_t)blockIdx.x - sh.blk0;
  if (lb == 0 && tid == 0) {
    sh.idx_out[m] = (int64_t)jglob;
    if (m > 0 && sh.trace_out) sh.trace_out[m - 1] = trace_s;
  }
  const double amp = exp(log_scaling[0]);
  const double amp2 = amp * amp;
  const int64_t i = lb * STHREADS + tid;
  double v = -1.0, idx = -1.0, tr = 0.0;
  if (i < sh.N)
    update_point(sh.X, sh.N, sh.offset, D, m, cj, xj, w, dj, jglob, amp2, nullptr, jitter, sh.d, sh.C, sh.chosen, i, v,
                 idx, tr);
  tr = warp_sum(tr);
  if ((tid & 31) == 0) st[tid >> 5] = tr;
  block_argmax(v, idx, /*prefer_high=*/true, sv, si);  // contains a __syncthreads
  if (tid == 0) {
    double t = 0.0;
    for (int q = 0; q < STHREADS / 32; q++) t += st[q];
    sh.blocktrace[lb] = t;
    sh.blockcand[2 * lb] = v;
    sh.blockcand[2 * lb + 1] = idx;
  }
}

// one CTA per local shard: reduce the per-block candidates, gather the shard's record and send it to every destination
__global__ void __launch_bounds__(1024) select_finalize_sharded_kernel(const ShardSet S, int D, int mrows,
                                                                       int prefer_high, const TargetSet T, int pos0,
                                                                       int64_t stride, int64_t slot_off,
                                                                       unsigned long long tag) {
  __shared__ double sv[32], si[32], stv[32];
  __shared__ double win_v, win_i, win_t;
  const ShardView& sh = S.s[blockIdx.x];
  const int tid = threadIdx.x;
  double v = -1.0, idx = -1.0, tr = 0.0;
  for (int64_t b = tid; b < sh.nblk; b += 1024) {
    double bv = sh.blockcand[2 * b], bi = sh.blockcand[2 * b + 1];
    if (better(bv, bi, v, idx, prefer_high)) { v = bv; idx = bi; }
    tr += sh.blocktrace[b];
  }
  tr = warp_sum(tr);
  if ((tid & 31) == 0) stv[tid >> 5] = tr;
  block_argmax(v, idx, prefer_high, sv, si);
  if (tid == 0) {
    win_v = v;
    win_i = idx;
    double t = 0.0;
    for (int q = 0; q < 32; q++) t += stv[q];
    win_t = t;
  }
  __syncthreads();
  const double wv = win_v, wi = win_i, wt = win_t;
  const int64_t pos = (int64_t)(pos0 + (int)blockIdx.x) * stride + slot_off;
  const int64_t il = (wi < 0.0) ? -1 : (int64_t)wi - sh.offset;
  const int payload = D + mrows;
  for (int k = tid; k < payload; k += 1024) {
    double val = 0.0;
    if (il >= 0) val = (k < D) ? sh.X[il * D + k] : sh.C[(int64_t)(k - D) * sh.N + il];
    for (int t = 0; t < T.n; t++) T.base[t][pos + SEL_HDR + k] = val;
  }
  if (tid < 3) {
    const double h = tid == 0 ? wv : (tid == 1 ? wi : wt);
    for (int t = 0; t < T.n; t++) T.base[t][pos + tid] = h;
  }
  __threadfence_system();  // every thread's peer stores are performed before ...
  __syncthreads();
  if (tid < T.n) {         // ... the tag becomes visible (one thread per destination)
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(T.base[tid] + pos + 3), "l"(tag) : "memory");
  }
}

// the last pivot: nothing left to update, only the winner is read
__global__ void __launch_bounds__(32) select_last_resolve_kernel(const ShardSet S, int M_sel, int R, int64_t stride,
                                                                 int64_t slot_off, unsigned long long tag) {
  __shared__ int best_s;
  __shared__ double trace_s;
  const ShardView& sh = S.s[blockIdx.x];
  if (*(volatile int*)sh.status != 0) return;
  const double* slot = sh.mailbox + slot_off;
  const int best = wait_and_resolve(slot, R, stride, tag, false, sh.status, &best_s, &trace_s);
  if (best < 0) return;
  if (threadIdx.x == 0) {
    sh.idx_out[M_sel - 1] = (int64_t)__ldcg(slot + (int64_t)best * stride + 1);
    if (sh.trace_out) sh.trace_out[M_sel - 2] = trace_s;
  }
}

__global__ void select_prepare_kernel(int64_t* idx_out, int M_sel, int* status) {
  for (int k = threadIdx.x; k < M_sel; k += blockDim.x) idx_out[k] = -1;
  if (threadIdx.x == 0) *status = 0;
}

struct ShardWs {
  double* d;
  double* C;
  unsigned char* chosen;
  double* blockcand;
  double* blocktrace;
  int* status;
  int64_t nblk;
};
__host__ size_t shard_ws_bytes(int64_t N, int M_sel) {
  int64_t nblk = (N + STHREADS - 1) / STHREADS;
  if (nblk < 1) nblk = 1;
  return align16((size_t)N * 8) + align16((size_t)N * (size_t)(M_sel - 1) * 8) + align16((size_t)N) +
         align16((size_t)nblk * 16) + align16((size_t)nblk * 8) + 16;
}
__host__ ShardWs carve_shard(void* ws, int64_t N, int M_sel) {
  ShardWs s;
  unsigned char* p = (unsigned char*)ws;
  s.nblk = (N + STHREADS - 1) / STHREADS;
  if (s.nblk < 1) s.nblk = 1;
  s.d = (double*)p;          p += align16((size_t)N * 8);
  s.C = (double*)p;          p += align16((size_t)N * (size_t)(M_sel - 1) * 8);
  s.chosen = p;              p += align16((size_t)N);
  s.blockcand = (double*)p;  p += align16((size_t)s.nblk * 16);
  s.blocktrace = (double*)p; p += align16((size_t)s.nblk * 8);
  s.status = (int*)p;
  return s;
}

// the pivot loop shared by the multi-GPU and the virtual-shard entry points.  `gather`: NCCL transport (T holds the send
// buffer; after every finalize the records are all-gathered into the mailbox slot); otherwise T holds the R mailboxes.
int run_sharded_selection(const ShardSet& S, int R, int pos0, const TargetSet& T, gvi_comm* gather, int D, int M_sel,
                          const double* log_scaling, const double* log_ls, int ls_len, double jitter,
                          unsigned long long tag0, cudaStream_t st) {
  const int64_t rl_max = SEL_HDR + D + (M_sel - 1);
  int64_t total_blocks = 0;
  for (int s = 0; s < S.n; s++) total_blocks += S.s[s].nblk;
  GVI_CHECK_ARG(total_blocks < (1ll << 31), "too many points for one launch");
  static std::atomic<unsigned long long> configured{0};
  if (gvi_first_use_on_this_device(configured))
    GVI_CUDA(cudaFuncSetAttribute(select_update_sharded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int s = 0; s < S.n; s++) {
    const ShardView& sh = S.s[s];
    select_prepare_kernel<<<1, 256, 0, st>>>(sh.idx_out, M_sel, sh.status);
    GVI_CHECK_LAUNCH("select_prepare_kernel");
    select_init_kernel<<<(unsigned)sh.nblk, STHREADS, 0, st>>>(sh.N, sh.offset, log_scaling, nullptr, jitter, sh.d,
                                                               sh.chosen, sh.blockcand, sh.blocktrace, nullptr);
    GVI_CHECK_LAUNCH("select_init_kernel");
  }
  // step -1: the first candidates (argmax of d0: lowest index among ties)
  for (int m = 0; m < M_sel; m++) {
    // records for step m: count = SEL_HDR + D + m doubles, slot m & 1
    const int64_t count = SEL_HDR + D + m;
    const int64_t stride = gather ? count : rl_max;
    const int64_t slot_off = (int64_t)(m & 1) * R * rl_max;
    const unsigned long long tag = SEL_TAG_BASE | ((tag0 + (unsigned long long)m) & 0x0000FFFFFFFFFFFFull);
    select_finalize_sharded_kernel<<<S.n, 1024, 0, st>>>(S, D, /*mrows=*/m, /*prefer_high=*/m > 0, T,
                                                          gather ? 0 : pos0, stride, gather ? 0 : slot_off, tag);
    GVI_CHECK_LAUNCH("select_finalize_sharded_kernel");
    if (gather) {
      int rc = gvi_comm_allgather_f64(gather, T.base[0], S.s[0].mailbox + slot_off, count, st);
      if (rc) return rc;
    }
    if (m == M_sel - 1) {
      select_last_resolve_kernel<<<S.n, 32, 0, st>>>(S, M_sel, R, stride, slot_off, tag);
      GVI_CHECK_LAUNCH("select_last_resolve_kernel");
      break;
    }
    const size_t smem = (size_t)(m + 2 * D) * sizeof(double);
    select_update_sharded_kernel<<<(unsigned)total_blocks, STHREADS, smem, st>>>(S, D, m, R, stride, slot_off, tag,
                                                                                log_scaling, log_ls, ls_len, jitter);
    GVI_CHECK_LAUNCH("select_update_sharded_kernel");
  }
  return GVI_OK;
}
}  // namespace

extern "C" size_t gvi_select_sharded_workspace_bytes(int64_t N_local, int M_sel, int D) {
  if (N_local < 0 || M_sel < 2 || D < 1) return 0;
  return shard_ws_bytes(N_local, M_sel) + 64;
}

extern "C" int gvi_cond_var_select_sharded_f64(const double* x_local, int64_t N_local, int64_t offset, int D, int M_sel,
                                               const double* log_scaling, const double* log_ls, int ls_len,
                                               double jitter, int64_t* idx_out, double* trace_out, void* workspace,
                                               gvi_comm* comm, void* stream) {
  GVI_CHECK_ARG(M_sel > 1, "Must have at least 2 inducing points");
  GVI_CHECK_ARG(N_local >= 0 && D >= 1 && offset >= 0, "sizes");
  GVI_CHECK_ARG((x_local || N_local == 0) && log_scaling && log_ls && idx_out && workspace && comm, "null pointer");
  GVI_CHECK_ARG(ls_len == 1 || ls_len == D, "log_lengthscales must have 1 or D entries");
  const int R = comm->world;
  const int64_t rl_max = SEL_HDR + D + (M_sel - 1);
  if ((size_t)(2 * R * rl_max) * 8 > comm->window_bytes || (size_t)rl_max * 8 > comm->scratch_bytes) {
    gvi_set_error("selector records of %lld doubles x %d ranks do not fit the communicator window", (long long)rl_max, R);
    return GVI_EUNSUPPORTED;
  }
  ShardWs w = carve_shard(workspace, N_local, M_sel);
  ShardSet S{};
  S.n = 1;
  S.s[0] = ShardView{x_local, N_local, offset, w.d, w.C, w.chosen, w.blockcand, w.blocktrace, 0, w.nblk,
                     (double*)comm->window, idx_out, trace_out, w.status};
  TargetSet T{};
  gvi_comm* gather = nullptr;
  if (comm->peer_ok) {
    T.n = R;
    for (int r = 0; r < R; r++) T.base[r] = (double*)comm->peer_window[r];
  } else {
    T.n = 1;
    T.base[0] = (double*)comm->scratch;
    gather = comm;
  }
  const unsigned long long tag0 = comm->next_tag;
  comm->next_tag += (unsigned long long)M_sel;
  return run_sharded_selection(S, R, comm->rank, T, gather, D, M_sel, log_scaling, log_ls, ls_len, jitter, tag0,
                               (cudaStream_t)stream);
}

// R virtual shards on ONE device: bounds_h[R+1] (host array, ascending, bounds_h[0] = 0, bounds_h[R] = N; empty shards
// allowed).  idx_out is [R][M_sel] (every shard's view of the pivots - they must agree), trace_out [R][M_sel-1] or NULL.
extern "C" size_t gvi_select_virtual_shards_workspace_bytes(int64_t N, int M_sel, int D, int R) {
  if (N < 1 || M_sel < 2 || D < 1 || R < 1 || R > SEL_MAX_LOCAL) return 0;
  // an upper bound that does not depend on where the bounds fall
  const int64_t rl_max = SEL_HDR + D + (M_sel - 1);
  return shard_ws_bytes(N, M_sel) + (size_t)R * (shard_ws_bytes(0, M_sel) + 64 * (size_t)M_sel + 4096) +
         (size_t)R * align16((size_t)(2 * R * rl_max) * 8) + 256;
}

extern "C" int gvi_cond_var_select_virtual_shards_f64(const double* x_perm, int64_t N, int D, int M_sel,
                                                      const double* log_scaling, const double* log_ls, int ls_len,
                                                      double jitter, int R, const int64_t* bounds_h, int64_t* idx_out,
                                                      double* trace_out, void* workspace, void* stream) {
  GVI_CHECK_ARG(M_sel > 1, "Must have at least 2 inducing points");
  GVI_CHECK_ARG(N >= 1 && D >= 1 && R >= 1 && R <= SEL_MAX_LOCAL, "sizes");
  GVI_CHECK_ARG(x_perm && log_scaling && log_ls && idx_out && workspace && bounds_h, "null pointer");
  GVI_CHECK_ARG(ls_len == 1 || ls_len == D, "log_lengthscales must have 1 or D entries");
  GVI_CHECK_ARG(bounds_h[0] == 0 && bounds_h[R] == N, "bounds must cover [0, N)");
  for (int r = 0; r < R; r++) GVI_CHECK_ARG(bounds_h[r] <= bounds_h[r + 1], "bounds must ascend");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rl_max = SEL_HDR + D + (M_sel - 1);
  unsigned char* p = (unsigned char*)workspace;
  ShardSet S{};
  S.n = R;
  TargetSet T{};
  T.n = R;
  int64_t blk = 0;
  for (int r = 0; r < R; r++) {
    const int64_t n = bounds_h[r + 1] - bounds_h[r];
    ShardWs w = carve_shard(p, n, M_sel);
    p += (shard_ws_bytes(n, M_sel) + 15) / 16 * 16;
    double* mb = (double*)p;
    p += align16((size_t)(2 * R * rl_max) * 8);
    GVI_CUDA(cudaMemsetAsync(mb, 0, (size_t)(2 * R * rl_max) * 8, st));
    S.s[r] = ShardView{x_perm + bounds_h[r] * D, n, bounds_h[r], w.d, w.C, w.chosen, w.blockcand, w.blocktrace, blk,
                       w.nblk, mb, idx_out + (int64_t)r * M_sel, trace_out ? trace_out + (int64_t)r * (M_sel - 1) : nullptr,
                       w.status};
    T.base[r] = mb;
    blk += w.nblk;
  }
  return run_sharded_selection(S, R, 0, T, nullptr, D, M_sel, log_scaling, log_ls, ls_len, jitter, 1ull, st);
}
