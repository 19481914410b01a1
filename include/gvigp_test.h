/*
 * gvigp_test.h - entry points that exist ONLY in libgvigp_test.so: the same sources as libgvigp.so compiled with
 * -DGVI_TEST_HOOKS, loaded by tests/ and tools/ and never by the product path.  That build also honours the profiling
 * environment switches GVI_DEBUG_SKIP / GVI_DEBUG_GRID / GVI_ONE_CTA / GVI_STATIC_TILES (csrc/predict.cu); the product
 * library exports none of this and never reads the environment.
 */
#ifndef GVIGP_TEST_H
#define GVIGP_TEST_H
#include "gvigp.h"
#ifdef __cplusplus
extern "C" {
#endif

/* test hook: out[i] = the branch-free exp used inside the fused kernels (<= 1 ulp; tests/test_gpu_parity.py) */
int gvi_test_exp_f64(const double* x, int64_t n, double* out, void* stream);
/* test hook: cap the chunk length of the gradient step's a_i ring at `waves` x 3552 points (0 = default, up to 1 GB),
 * so that small problems exercise the multi-chunk path; affects gvi_fused_step_grad_workspace_bytes too.  Returns the
 * previous value. */
int gvi_test_set_grad_chunk_waves(int waves);

/* test hooks: the in-tree FP64 tensor-pipe GEMM / GEMV (csrc/dgemm.cu) on their own.  Row-major C[m,n] = op(A) op(B) +
 * beta C; kmode 0 = full k range, 1 = op(B)[k][n] is zero for k > n, 2 = zero for k < n; symmetric: only the tiles on or
 * below the diagonal are multiplied, the others mirrored.  y[m] = op(A) x + beta y. */
int gvi_test_dgemm_f64(int ta, int tb, int m, int n, int k, const double* A, int64_t lda, const double* B, int64_t ldb,
                       double beta, double* C, int64_t ldc, int kmode, int symmetric, void* stream);
int gvi_test_dgemv_f64(int ta, int m, int k, const double* A, int64_t lda, const double* x, double beta, double* y,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GVIGP_TEST_H */
