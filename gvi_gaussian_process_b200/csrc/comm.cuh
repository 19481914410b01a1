// comm.cuh - the communicator object behind gvi_comm_* (comm.cu) as the kernels' launchers see it (select.cu).
#pragma once
#include "common.cuh"

constexpr int GVI_COMM_MAX_RANKS = 16;
constexpr size_t GVI_COMM_WINDOW_BYTES = 4u << 20;   // per-rank peer window (selector mailboxes: 2 slots x R records)
constexpr size_t GVI_COMM_SCRATCH_BYTES = 1u << 20;  // private staging (gathered small vectors, send records)

struct gvi_comm {
  int rank = 0, world = 1, device = 0;
  int peer_ok = 0;        // every rank mapped every other rank's window (else: NCCL transport)
  void* nccl = nullptr;   // ncclComm_t
  void* window = nullptr; // this rank's window
  void* peer_window[GVI_COMM_MAX_RANKS] = {};  // [r] = rank r's window as mapped into this process ([rank] = window)
  size_t window_bytes = 0;
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // tags written into the window grow monotonically over the communicator's lifetime, so a window never has to be
  // cleared between collective calls (all ranks make the same calls in the same order)
  uint64_t next_tag = 1;
};
