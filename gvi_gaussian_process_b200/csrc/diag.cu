// diag.cu - measurement aids of the library itself (no reference counterpart): the FP64 tensor-pipe issue rate of the
// device the caller is on, so that bench.py divides by a peak measured on the box the number comes from
// (MEASURED_PEAKS.json carries no FP64 entry; profiles/fp64_peak_r01.txt is where the 37.1 TFLOP/s constant came from).
#include "common.cuh"

namespace {
// 8 warps per CTA, 2 CTAs per SM, 16 independent m8n8k4 accumulator pairs per warp: the shape that reached the highest
// DMMA issue rate in tools/fp64_peak.cu (64 FMA / clk / SM)
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    c[i][0] = threadIdx.x * 1e-3;
    c[i][1] = i;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

// Runs for about `target_ms` milliseconds on `stream`, SYNCHRONISES it, and writes the sustained rate (TFLOP/s, 2 flop
// per multiply-add) to the HOST double *tflops_out_h.  scratch: device buffer of >= gvi_fp64_tensor_peak_scratch_bytes().
extern "C" size_t gvi_fp64_tensor_peak_scratch_bytes(void) { return (size_t)2 * GVI_NUM_SMS * 256 * sizeof(double); }

extern "C" int gvi_fp64_tensor_peak_tflops(double target_ms, double* tflops_out_h, void* scratch, void* stream) {
  GVI_CHECK_ARG(tflops_out_h && scratch && target_ms > 0.0 && target_ms <= 2000.0, "arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = 2 * GVI_NUM_SMS, threads = 256;
  cudaEvent_t e0, e1;
  GVI_CUDA(cudaEventCreate(&e0));
  GVI_CUDA(cudaEventCreate(&e1));
  int iters = 4000;
  float ms = 0.f;
  for (int pass = 0; pass < 3; pass++) {  // warm-up / calibration, then the timed run
    cudaEventRecord(e0, st);
    dmma_peak_kernel<<<blocks, threads, 0, st>>>((double*)scratch, iters, 1.0000001, 1e-9);
    cudaError_t le = cudaGetLastError();
    cudaEventRecord(e1, st);
    cudaError_t se = cudaEventSynchronize(e1);
    gvi_count_launch();
    if (le != cudaSuccess || se != cudaSuccess) {
      gvi_set_error("CUDA error in dmma_peak_kernel: %s", cudaGetErrorString(le != cudaSuccess ? le : se));
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      return GVI_ECUDA;
    }
    cudaEventElapsedTime(&ms, e0, e1);
    if (pass < 2 && ms > 0.f) {
      double scale = target_ms / ms;
      if (scale > 64.0) scale = 64.0;
      iters = (int)(iters * scale) + 1;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // one m8n8k4 = 8*8*4 multiply-adds per warp
  const double flops = 2.0 * 256.0 * 16.0 * (double)iters * (threads / 32) * blocks;
  *tflops_out_h = flops / (ms * 1e-3) / 1e12;
  return GVI_OK;
}
