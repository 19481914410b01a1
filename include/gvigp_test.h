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

#ifdef __cplusplus
}
#endif
#endif /* GVIGP_TEST_H */
