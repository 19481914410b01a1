// gram_lab.cu - throw-away variants of the materialised-Gram tile kernel, timed side by side (one gpurun call):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gvi_gaussian_process_b200/csrc tools/lab/gram_lab.cu -o tools/lab/gram_lab
#include <cstdio>
#include <vector>
#include "common.cuh"
void gvi_set_error(const char*, ...) {}
void gvi_count_launch() {}

__device__ const double2 TAB[32] = {
    {1.0, 0.0}, {1.0218971486541166, 5.109225028973443e-17}, {1.0442737824274138, 8.551889705537965e-17},
    {1.0671404006768237, -7.899853966841582e-17}, {1.0905077326652577, -3.046782079812471e-17},
    {1.1143867425958924, 1.0410278456845571e-16}, {1.1387886347566916, 8.912812676025407e-17},
    {1.1637248587775775, 3.8292048369240935e-17}, {1.189207115002721, 3.982015231465646e-17},
    {1.215247359980469, -7.71263069268148e-17}, {1.241857812073484, 4.658027591836936e-17},
    {1.2690509571917332, 2.667932131342186e-18}, {1.2968395546510096, 2.5382502794888315e-17},
    {1.3252366431597413, -2.858731210038861e-17}, {1.3542555469368927, 7.700948379802989e-17},
    {1.383909881963832, -6.770511658794786e-17}, {1.4142135623730951, -9.667293313452913e-17},
    {1.4451808069770467, -3.023758134993987e-17}, {1.4768261459394993, -3.4839945568927958e-17},
    {1.5091644275934228, -1.016455327754295e-16}, {1.5422108254079407, 7.94983480969762e-17},
    {1.5759808451078865, -1.0136916471278304e-17}, {1.6104903319492543, 2.470719256979788e-17},
    {1.645755478153965, -1.0125679913674773e-16}, {1.681792830507429, 8.19901002058149e-17},
    {1.718619298122478, -1.851380418263111e-17}, {1.7562521603732995, 2.960140695448873e-17},
    {1.7947090750031072, 1.822745842791208e-17}, {1.8340080864093424, 3.283107224245627e-17},
    {1.8741676341103, -6.122763413004142e-17}, {1.9152065613971474, -1.0619946056195963e-16},
    {1.9571441241754002, 8.960767791036668e-17}};

template <int DP, int CT, int RT, int MODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) lab_kernel(const double* __restrict__ x1, int64_t m1,
                                                          const double* __restrict__ x2, int64_t m2, int D,
                                                          double* __restrict__ out, int64_t ld, int rblocks) {
  constexpr int RB = 64;
  __shared__ __align__(16) double xs[RB][DP];
  __shared__ double2 tab[32][8];
  const int tid = threadIdx.x;
  const int64_t col0 = (int64_t)blockIdx.x * (THREADS * CT) + (int64_t)tid * CT;
  if (tid < 256) tab[tid >> 3][tid & 7] = TAB[tid >> 3];
  double z[CT][DP];
#pragma unroll
  for (int c = 0; c < CT; c++)
#pragma unroll
    for (int d = 0; d < DP; d++) z[c][d] = (col0 + c < m2 && d < D) ? 0.25 * x2[(col0 + c) * D + d] : 0.0;
  const double2* tb = &tab[0][tid & 7];
  for (int blk = 0; blk < rblocks; blk++) {
    const int64_t row0 = ((int64_t)blockIdx.y * rblocks + blk) * RB;
    if (row0 >= m1) break;
    __syncthreads();
    for (int idx = tid; idx < RB * DP; idx += THREADS) {
      const int r = idx / DP, d = idx % DP;
      xs[r][d] = (row0 + r < m1 && d < D) ? 0.25 * x1[(row0 + r) * D + d] : 0.0;
    }
    __syncthreads();
    for (int r0 = 0; r0 < RB; r0 += RT) {
      double acc[RT][CT];
#pragma unroll
      for (int rr = 0; rr < RT; rr++)
#pragma unroll
        for (int c = 0; c < CT; c++) acc[rr][c] = 0.0;
      if (MODE != 2) {
#pragma unroll
        for (int d = 0; d < DP; d += 2) {
#pragma unroll
          for (int rr = 0; rr < RT; rr++) {
            const double2 xv = *reinterpret_cast<const double2*>(&xs[r0 + rr][d]);
#pragma unroll
            for (int c = 0; c < CT; c++) {
              const double d0 = xv.x - z[c][d];
              acc[rr][c] = fma(d0, d0, acc[rr][c]);
              const double d1 = xv.y - z[c][d + 1];
              acc[rr][c] = fma(d1, d1, acc[rr][c]);
            }
          }
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < RT; rr++)
#pragma unroll
          for (int c = 0; c < CT; c++) acc[rr][c] = fabs(xs[r0 + rr][0] - z[c][0]);
      }
      double v[RT][CT];
#pragma unroll
      for (int rr = 0; rr < RT; rr++)
#pragma unroll
        for (int c = 0; c < CT; c++) v[rr][c] = (MODE == 1) ? acc[rr][c] : gvi_exp_neg_t32<false>(acc[rr][c], tb, 8);
      double* dst = out + (row0 + r0) * ld + col0;
#pragma unroll
      for (int rr = 0; rr < RT; rr++) {
        if (CT == 2) {
          __stcs(reinterpret_cast<double2*>(dst + rr * ld), make_double2(v[rr][0], v[rr][CT - 1]));
        } else if (CT == 4) {
          __stcs(reinterpret_cast<double2*>(dst + rr * ld), make_double2(v[rr][0], v[rr][1]));
          __stcs(reinterpret_cast<double2*>(dst + rr * ld + 2), make_double2(v[rr][CT - 2], v[rr][CT - 1]));
        } else {
          __stcs(dst + rr * ld, v[rr][0]);
        }
      }
    }
  }
}

template <int DP, int CT, int RT, int MODE, int THREADS, int MINB>
void run(const char* name, const double* x, int64_t n, const double* z, int64_t m, int D, double* out, int rblocks) {
  dim3 grid((unsigned)((m + THREADS * CT - 1) / (THREADS * CT)), (unsigned)((n + 64 * rblocks - 1) / (64 * rblocks)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 2; i++) lab_kernel<DP, CT, RT, MODE, THREADS, MINB><<<grid, THREADS>>>(x, n, z, m, D, out, m, rblocks);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) lab_kernel<DP, CT, RT, MODE, THREADS, MINB><<<grid, THREADS>>>(x, n, z, m, D, out, m, rblocks);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  printf("%-44s %7.3f ms  %6.0f GB/s  (%s)\n", name, ms, 8.0 * n * m / ms * 1e-6, cudaGetErrorString(e));
}

int main() {
  const int64_t n = 400000, m = 1024;
  const int D = 8;
  std::vector<double> hx(n * D), hz(m * D);
  for (size_t i = 0; i < hx.size(); i++) hx[i] = (double)((i * 2654435761u) % 2000) / 500.0 - 2.0;
  for (size_t i = 0; i < hz.size(); i++) hz[i] = (double)((i * 40503u) % 2000) / 500.0 - 2.0;
  double *x, *z, *out;
  cudaMalloc(&x, hx.size() * 8);
  cudaMalloc(&z, hz.size() * 8);
  cudaMalloc(&out, n * m * 8);
  cudaMemcpy(x, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(z, hz.data(), hz.size() * 8, cudaMemcpyHostToDevice);
  cudaMemset(out, 0, n * m * 8);
  run<8, 2, 4, 0, 256, 3>("full  CT2 RT4 256thr x3", x, n, z, m, D, out, 4);
  run<8, 2, 4, 1, 256, 3>("dist  CT2 RT4 256thr x3 (no exp)", x, n, z, m, D, out, 4);
  run<8, 2, 4, 2, 256, 3>("exp   CT2 RT4 256thr x3 (no distance)", x, n, z, m, D, out, 4);
  run<8, 2, 8, 0, 256, 2>("full  CT2 RT8 256thr x2", x, n, z, m, D, out, 4);
  run<8, 2, 8, 0, 256, 3>("full  CT2 RT8 256thr x3", x, n, z, m, D, out, 4);
  run<8, 2, 2, 0, 256, 4>("full  CT2 RT2 256thr x4", x, n, z, m, D, out, 4);
  run<8, 2, 4, 0, 256, 4>("full  CT2 RT4 256thr x4 (64 regs)", x, n, z, m, D, out, 4);
  run<8, 2, 4, 0, 128, 6>("full  CT2 RT4 128thr x6", x, n, z, m, D, out, 4);
  run<8, 2, 4, 0, 512, 1>("full  CT2 RT4 512thr x1", x, n, z, m, D, out, 4);
  run<8, 4, 2, 0, 256, 2>("full  CT4 RT2 256thr x2", x, n, z, m, D, out, 4);
  run<8, 4, 4, 0, 256, 2>("full  CT4 RT4 256thr x2", x, n, z, m, D, out, 4);
  run<8, 1, 8, 0, 256, 4>("full  CT1 RT8 256thr x4", x, n, z, m, D, out, 4);
  run<8, 2, 4, 0, 256, 3>("full  CT2 RT4 256thr x3 rblocks16", x, n, z, m, D, out, 16);
  return 0;
}
