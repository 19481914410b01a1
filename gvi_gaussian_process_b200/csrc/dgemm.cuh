// dgemm.cuh - the in-tree FP64 tensor-pipe GEMM / GEMV (dgemm.cu) as the other translation units call it.
#pragma once
#include "common.cuh"

constexpr int GVI_GEMM_K_FULL = 0;
constexpr int GVI_GEMM_K_UPTO_N = 1;  // op(B)[k][n] == 0 for k > n (e.g. B = L^-1 read transposed): k loop stops at the tile
constexpr int GVI_GEMM_K_FROM_N = 2;  // op(B)[k][n] == 0 for k < n (e.g. B = L^-1): k loop starts at the tile

// Row-major C[m,n] = op(A) op(B) + beta C on the caller's stream.  ta: op(A) = A^T with A stored [k][m] (lda >= m), else A
// stored [m][k]; tb: op(B) = B^T with B stored [n][k], else B stored [k][n].  symmetric: m == n and the result is known
// to be symmetric - only the tiles on or below the diagonal are computed, the others are written as mirror images.
int gvi_dgemm_rm(bool ta, bool tb, int m, int n, int k, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double beta, double* C, int64_t ldc, int kmode, bool symmetric, cudaStream_t st);
// y[m] = op(A) x + beta y for row-major A ([m][k], or [k][m] when ta)
int gvi_dgemv_rm(bool ta, int m, int k, const double* A, int64_t lda, const double* x, double beta, double* y,
                 cudaStream_t st);
