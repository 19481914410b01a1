// chol_lab.cu - phase timing of the cooperative factorisation (CTA 0's %globaltimer stamps):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGVI_COOP_STAMPS -I gvi_gaussian_process_b200/csrc tools/lab/chol_lab.cu -o tools/lab/chol_lab
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <map>
#include <string>
#include <cstdarg>
#include "../../gvi_gaussian_process_b200/csrc/chol_coop.cu"
void gvi_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); }
void gvi_count_launch() {}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 1024;
  std::vector<double> h((size_t)M * M);
  // SPD: exp(-|i-j|/50) + I
  for (int i = 0; i < M; i++)
    for (int j = 0; j < M; j++) h[(size_t)i * M + j] = exp(-abs(i - j) / 50.0) + (i == j ? 1.0 : 0.0);
  double *A, *X, *P;
  int* info;
  long long* stamps;
  cudaMalloc(&A, (size_t)M * M * 8);
  cudaMalloc(&X, (size_t)M * M * 8);
  cudaMalloc(&P, gvi_chol_coop_scratch_doubles(M) * 8);
  cudaMalloc(&info, 4);
  cudaMalloc(&stamps, 8 * 20000);
  for (int rep = 0; rep < 3; rep++) {
    cudaMemcpy(A, h.data(), (size_t)M * M * 8, cudaMemcpyHostToDevice);
    cudaMemset(stamps, 0, 8 * 20000);
    int big = 0x7fffffff;
    cudaMemcpy(info, &big, 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int rc = gvi_launch_chol_inverse_coop(A, X, P, M, M, info, 0, stamps);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("M=%d rep %d: rc %d %s %.3f ms\n", M, rep, rc, cudaGetErrorString(e), ms);
  }
  std::vector<long long> s(20000);
  cudaMemcpy(s.data(), stamps, 8 * 20000, cudaMemcpyDeviceToHost);
  const long long n = s[0];
  std::map<std::pair<int, int>, std::pair<long long, int>> agg;  // (from id, to id) -> (ns, count)
  for (long long i = 1; i < n; i++) {
    auto key = std::make_pair((int)s[2 * (i - 1) + 2], (int)s[2 * i + 2]);
    agg[key].first += s[2 * i + 3] - s[2 * (i - 1) + 3];
    agg[key].second++;
  }
  const char* names[] = {"start", "diag0 done", "sync0 done", "panel done", "panel sync done", "update(0) done", "update+diag done",
                         "update sync done", "inv products done", "inv sync1 done", "inv reduce done", "inv sync2 done"};
  auto nm = [&](int id) -> std::string { return id < 12 ? names[id] : "diag#" + std::to_string(id); };
  for (auto& kv : agg)
    printf("  %-20s -> %-20s : n=%4d total %8.1f us  avg %6.2f us\n", nm(kv.first.first).c_str(), nm(kv.first.second).c_str(),
           kv.second.second, kv.second.first / 1e3, kv.second.first / 1e3 / kv.second.second);
  return 0;
}
