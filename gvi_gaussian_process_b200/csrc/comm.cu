// comm.cu - gvi_comm: the multi-GPU side of the C ABI (SURVEY.md section 8(b) "gvi_comm_init/destroy", 8(e)).
// One rank per process and GPU.  A communicator owns
//   - an ncclComm_t (NCCL is bound at run time with dlopen: a single-GPU host never needs libnccl), used for bootstrap,
//     for the all-reduce of the risk / gradient sums and as the fall-back transport of the selector's exchange;
//   - a PEER WINDOW: a small device buffer per rank that every other rank maps with CUDA IPC over NVLink / NVSwitch.
//     The greedy selector's per-pivot exchange writes one candidate record straight into every peer's window from the
//     kernel that produced it (st.global on mapped peer pointers + a tag word), so a pivot costs one NVLink store
//     latency instead of a host-issued collective (select.cu).
// The reference has no multi-device path at all (pure single-process JAX); the split it would need is SURVEY.md 8(e).
#include <dlfcn.h>
#include <nccl.h>  // types only: every NCCL entry point is resolved with dlsym
#include <cstdlib>
#include <cstring>

#include "comm.cuh"

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

// libnccl.so.2 already mapped by the host framework (torch ships its own) is found by soname; otherwise the system one.
int load_nccl() {
  if (g_nccl.handle) return GVI_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    gvi_set_error("cannot load libnccl.so.2: %s", dlerror());
    return GVI_ECOMM;
  }
  NcclApi a;
  a.handle = h;
#define GVI_SYM(field, name)                                        \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));    \
  if (!a.field) {                                                   \
    gvi_set_error("libnccl has no symbol %s", name);                \
    return GVI_ECOMM;                                               \
  }
  GVI_SYM(GetUniqueId, "ncclGetUniqueId")
  GVI_SYM(CommInitRank, "ncclCommInitRank")
  GVI_SYM(CommDestroy, "ncclCommDestroy")
  GVI_SYM(AllGather, "ncclAllGather")
  GVI_SYM(AllReduce, "ncclAllReduce")
  GVI_SYM(GetErrorString, "ncclGetErrorString")
#undef GVI_SYM
  g_nccl = a;
  return GVI_OK;
}

#define GVI_NCCL(call)                                                              \
  do {                                                                              \
    ncclResult_t r__ = (call);                                                      \
    if (r__ != ncclSuccess) {                                                       \
      gvi_set_error("NCCL error in %s: %s", #call, g_nccl.GetErrorString(r__));     \
      return GVI_ECOMM;                                                             \
    }                                                                               \
  } while (0)

// out[i] = sum over ranks, in rank order, of gathered[r*count + i]: every rank computes the same bits, whatever
// algorithm the transport picked (SURVEY.md "Reduction determinism")
__global__ void __launch_bounds__(256) rank_order_sum_kernel(const double* __restrict__ gathered, int R, int64_t count,
                                                             double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= count) return;
  double s = gathered[i];
  for (int r = 1; r < R; r++) s += gathered[(int64_t)r * count + i];
  out[i] = s;
}
}  // namespace

extern "C" int gvi_comm_unique_id_h(void* id128_h) {
  GVI_CHECK_ARG(id128_h, "null pointer");
  static_assert(sizeof(ncclUniqueId) == GVI_COMM_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  GVI_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128_h, &id, sizeof(id));
  return GVI_OK;
}

extern "C" int gvi_comm_destroy(gvi_comm* c) {
  if (!c) return GVI_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; r++)
    if (r != c->rank && c->peer_window[r]) cudaIpcCloseMemHandle(c->peer_window[r]);
  if (c->nccl) g_nccl.CommDestroy((ncclComm_t)c->nccl);
  if (c->window) cudaFree(c->window);
  if (c->scratch) cudaFree(c->scratch);
  if (prev >= 0) cudaSetDevice(prev);
  delete c;
  return GVI_OK;
}

extern "C" int gvi_comm_init(gvi_comm** comm_out, int rank, int world, const void* id128_h, int flags) {
  GVI_CHECK_ARG(comm_out, "null pointer");
  GVI_CHECK_ARG(world >= 1 && world <= GVI_COMM_MAX_RANKS && rank >= 0 && rank < world, "rank / world size");
  GVI_CHECK_ARG(world == 1 || id128_h, "a unique id is needed for world > 1");
  *comm_out = nullptr;
  gvi_comm* c = new gvi_comm();
  c->rank = rank;
  c->world = world;
  c->window_bytes = GVI_COMM_WINDOW_BYTES;
  c->scratch_bytes = GVI_COMM_SCRATCH_BYTES;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc(&c->window, c->window_bytes);
  if (e == cudaSuccess) e = cudaMemset(c->window, 0, c->window_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&c->scratch, c->scratch_bytes);
  if (e != cudaSuccess) {
    gvi_set_error("CUDA error allocating the communicator buffers: %s", cudaGetErrorString(e));
    gvi_comm_destroy(c);
    return GVI_ECUDA;
  }
  c->peer_window[rank] = c->window;
  if (world == 1) {
    c->peer_ok = 1;
    *comm_out = c;
    return GVI_OK;
  }
  int rc = load_nccl();
  if (rc) {
    gvi_comm_destroy(c);
    return rc;
  }
  ncclUniqueId id;
  memcpy(&id, id128_h, sizeof(id));
  ncclComm_t nc = nullptr;
  ncclResult_t nr = g_nccl.CommInitRank(&nc, world, id, rank);
  if (nr != ncclSuccess) {
    gvi_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(nr));
    gvi_comm_destroy(c);
    return GVI_ECOMM;
  }
  c->nccl = nc;
  // ---- peer window: exchange the IPC handles (and a "my export worked" byte) through the NCCL communicator ---------
  struct Export {
    cudaIpcMemHandle_t handle;
    int ok;
    int pad[15];
  };
  static_assert(sizeof(Export) == 128, "export record");
  Export mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = (!(flags & GVI_COMM_NO_PEER_WINDOW) && cudaIpcGetMemHandle(&mine.handle, c->window) == cudaSuccess) ? 1 : 0;
  cudaGetLastError();
  Export all[GVI_COMM_MAX_RANKS];
  unsigned char* dev = (unsigned char*)c->scratch;  // [world + 1] export records
  e = cudaMemcpy(dev + (size_t)world * sizeof(Export), &mine, sizeof(Export), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    nr = g_nccl.AllGather(dev + (size_t)world * sizeof(Export), dev, sizeof(Export), ncclUint8, nc, (cudaStream_t)0);
    if (nr != ncclSuccess) {
      gvi_set_error("ncclAllGather (bootstrap) failed: %s", g_nccl.GetErrorString(nr));
      gvi_comm_destroy(c);
      return GVI_ECOMM;
    }
    e = cudaStreamSynchronize((cudaStream_t)0);
  }
  if (e == cudaSuccess) e = cudaMemcpy(all, dev, (size_t)world * sizeof(Export), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    gvi_set_error("CUDA error during the communicator bootstrap: %s", cudaGetErrorString(e));
    gvi_comm_destroy(c);
    return GVI_ECUDA;
  }
  int ok = 1;
  for (int r = 0; r < world; r++) ok &= all[r].ok;
  if (ok) {
    for (int r = 0; r < world && ok; r++) {
      if (r == rank) continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      } else {
        c->peer_window[r] = p;
      }
    }
  }
  // every rank must agree on the transport: one more (tiny) exchange of "all my mappings worked"
  int* flag_dev = (int*)c->scratch;
  int mine_ok = ok;
  int oks[GVI_COMM_MAX_RANKS];
  e = cudaMemcpy(flag_dev + world, &mine_ok, sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    nr = g_nccl.AllGather(flag_dev + world, flag_dev, sizeof(int), ncclUint8, nc, (cudaStream_t)0);
    if (nr != ncclSuccess) {
      gvi_set_error("ncclAllGather (bootstrap) failed: %s", g_nccl.GetErrorString(nr));
      gvi_comm_destroy(c);
      return GVI_ECOMM;
    }
    e = cudaStreamSynchronize((cudaStream_t)0);
  }
  if (e == cudaSuccess) e = cudaMemcpy(oks, flag_dev, (size_t)world * sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    gvi_set_error("CUDA error during the communicator bootstrap: %s", cudaGetErrorString(e));
    gvi_comm_destroy(c);
    return GVI_ECUDA;
  }
  for (int r = 0; r < world; r++) ok &= oks[r];
  c->peer_ok = ok;
  *comm_out = c;
  return GVI_OK;
}

extern "C" int gvi_comm_rank(const gvi_comm* c) { return c ? c->rank : -1; }
extern "C" int gvi_comm_size(const gvi_comm* c) { return c ? c->world : -1; }
extern "C" int gvi_comm_has_peer_window(const gvi_comm* c) { return c ? c->peer_ok : 0; }

extern "C" int gvi_comm_allgather_f64(gvi_comm* c, const double* send, double* recv, int64_t count, void* stream) {
  GVI_CHECK_ARG(c && send && recv && count >= 0, "arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->world == 1) {
    if (send != recv) GVI_CUDA(cudaMemcpyAsync(recv, send, (size_t)count * 8, cudaMemcpyDeviceToDevice, st));
    return GVI_OK;
  }
  GVI_NCCL(g_nccl.AllGather(send, recv, (size_t)count, ncclFloat64, (ncclComm_t)c->nccl, st));
  return GVI_OK;
}

extern "C" int gvi_comm_allreduce_sum_f64(gvi_comm* c, double* buf, int64_t count, void* stream) {
  GVI_CHECK_ARG(c && buf && count >= 0, "arguments");
  if (c->world == 1 || count == 0) return GVI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if ((size_t)count * 8 * (size_t)c->world <= c->scratch_bytes) {
    // small vectors (the [sum NLL, sum D0, N] triple, the D+4 gradient vector): gather, then sum in rank order
    GVI_NCCL(g_nccl.AllGather(buf, c->scratch, (size_t)count, ncclFloat64, (ncclComm_t)c->nccl, st));
    rank_order_sum_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>((const double*)c->scratch, c->world, count,
                                                                          buf);
    GVI_CHECK_LAUNCH("rank_order_sum_kernel");
    return GVI_OK;
  }
  GVI_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)c->nccl, st));
  return GVI_OK;
}
