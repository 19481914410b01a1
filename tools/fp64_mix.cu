// Do DFMA (FP64 pipe) and DMMA (tensor pipe) overlap on sm_100a?  Half the warps of every CTA run DFMA chains, the
// other half DMMA chains; compare with each alone.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: all warps DFMA, 1: all warps DMMA, 2: warps 0-3,8-11 DFMA / warps 4-7,12-15 DMMA (every sub-partition gets both)
__global__ void mix(double* out, int iters, double a, double b, int mode) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = mode == 1 || (mode == 2 && ((warp >> 2) & 1));  // warp%4 = sub-partition: both kinds on each
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3 + i;
  if (do_mma) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma884(acc[2 * i], acc[2 * i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = 512, blocks = 148, iters = 20000; float ms;
  for (int mode = 0; mode < 3; mode++) {
    mix<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9, mode); cudaDeviceSynchronize();
    cudaEventRecord(e0); mix<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, mode); cudaEventRecord(e1);
    cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1);
    double warps = blocks * threads / 32.0;
    double fl_fma = 2.0 * 16 * iters * 32, fl_mma = 2.0 * 256 * 8 * iters;  // per warp
    double fl = mode == 0 ? warps * fl_fma : mode == 1 ? warps * fl_mma : warps / 2 * (fl_fma + fl_mma);
    printf("mode %d (%s): %.3f ms  %.2f TFLOP/s total\n", mode, mode == 0 ? "DFMA only" : mode == 1 ? "DMMA only" : "half DFMA + half DMMA", ms, fl / ms * 1e-9);
  }
  return 0;
}
