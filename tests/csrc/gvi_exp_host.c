/* TEST INFRASTRUCTURE.  Runs the DEVICE source of gvi_exp (the text between the markers in csrc/common.cuh, pasted
 * in by tests/test_host_logic.py as gvi_exp_body.inc) on the host: the function is nothing but IEEE-754 double
 * fma / add / compare and an integer add into the exponent field, so the host evaluation is bit-for-bit what the GPU
 * computes for every non-NaN argument.  Reports the largest error in ulps against expl() over a dense sweep of the
 * argument range the kernels use (exponents -0.5 * s <= 0), the special values, and NaN propagation. */
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define __device__
#define __forceinline__ static inline
static inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
static inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __hiloint2double(int hi, int lo) {
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo;
  double d; memcpy(&d, &u, 8); return d;
}

#include "gvi_exp_body.inc"

static double ulps(double got, double x) {
  long double ref = expl((long double)x);
  double refd = (double)ref;
  double ulp = nextafter(refd, INFINITY) - refd;
  return (double)(fabsl((long double)got - ref) / (long double)ulp);
}

int main(void) {
  double worst = 0.0, worst_x = 0.0;
  long count = 0;
  /* dense uniform sweep of [-708, 0] and a log-spaced sweep towards 0 (small |x| is where most kernel values live) */
  for (long i = 0; i <= 4000000; i++) {
    double x = -708.0 * (double)i / 4000000.0;
    double e = ulps(gvi_exp(x), x);
    if (e > worst) { worst = e; worst_x = x; }
    count++;
  }
  for (int k = 0; k < 2000000; k++) {
    double x = -ldexp(1.0 + (double)(k % 1000) / 1000.0, -(k / 1000) % 60);
    if (x < -708.0) continue;
    double e = ulps(gvi_exp(x), x);
    if (e > worst) { worst = e; worst_x = x; }
    count++;
  }
  uint64_t s = 88172645463325252ULL; /* xorshift: random arguments */
  for (int k = 0; k < 2000000; k++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    double x = -708.0 * (double)(s >> 11) / 9007199254740992.0;
    double e = ulps(gvi_exp(x), x);
    if (e > worst) { worst = e; worst_x = x; }
    count++;
  }
  printf("max_ulp %.4f at %.17g over %ld arguments\n", worst, worst_x, count);
  printf("exp0 %d\n", gvi_exp(0.0) == 1.0 && gvi_exp(-0.0) == 1.0);
  printf("below %d\n", gvi_exp(-708.0000001) == 0.0 && gvi_exp(-1e300) == 0.0 && gvi_exp(-INFINITY) == 0.0);
  printf("edge %d\n", gvi_exp(-708.0) > 0.0 && ulps(gvi_exp(-708.0), -708.0) <= 1.0);
  double n = gvi_exp(NAN);
  printf("nan %d\n", n != n);
  return 0;
}
