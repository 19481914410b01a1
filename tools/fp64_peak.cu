// Microbenchmark: FP64 DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 f64) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NT>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__global__ void dmma1688_kernel(double* out, int iters, double av, double bv) {
  double c[4][4]; double a[4] = {av, av + 1, av + 2, av + 3}; double b[2] = {bv, bv + 1};
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x * 1e-3 + i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void exp_kernel(double* out, int iters, double a) {
  double x[4] = {a * threadIdx.x, a * 2, a * 3, a * 4}; double s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) { s += exp(x[i]); x[i] -= 1e-7; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  printf("dev %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, clk);
  double* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = p.multiProcessorCount; float ms;
  for (int wpb : {4, 8, 16, 32}) {
    int threads = wpb * 32; int blocks = sms * (wpb <= 16 ? 2 : 1); int iters = 20000;
    dfma_kernel<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * (double)threads * blocks;
    printf("DFMA  warps/blk %2d blocks %d: %.3f ms  %.2f TFLOP/s\n", wpb, blocks, ms, fl / ms * 1e-9);
  }
  for (int wpb : {4, 8, 16}) {
    int threads = wpb * 32; int blocks = sms; int iters = 20000;
    dmma_kernel<8><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 256 * 8 * iters * (double)wpb * blocks;
    printf("DMMA884 x8 warps/blk %2d: %.3f ms  %.2f TFLOP/s\n", wpb, ms, fl / ms * 1e-9);
    cudaEventRecord(e0); dmma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 256 * 16 * iters * (double)wpb * blocks;
    printf("DMMA884 x16 warps/blk %2d: %.3f ms  %.2f TFLOP/s\n", wpb, ms, fl / ms * 1e-9);
    cudaEventRecord(e0); dmma1688_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 1024 * 4 * iters * (double)wpb * blocks;
    printf("DMMA1688 x4 warps/blk %2d: %.3f ms  %.2f TFLOP/s\n", wpb, ms, fl / ms * 1e-9);
  }
  {
    int threads = 512, blocks = sms * 2, iters = 4000;
    exp_kernel<<<blocks, threads>>>(out, 10, -1e-3); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); exp_kernel<<<blocks, threads>>>(out, iters, -1e-3); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    double n = 4.0 * iters * threads * blocks;
    printf("exp(double): %.3f ms  %.2f Gexp/s\n", ms, n / ms * 1e-6);
  }
  return 0;
}
