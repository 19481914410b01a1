/* TEST INFRASTRUCTURE.  Runs the DEVICE source of gvi_log (table + function text of csrc/common.cuh, pasted in by
 * tests/test_host_logic.py as gvi_log_body.inc) on the host: IEEE-754 double fma / add / multiply, integer work on the
 * bit pattern and one table lookup, so the host evaluation is what the GPU computes.  Reports the largest error in ulps
 * against logl() over dense sweeps (the whole double range, [0.5, 2], the neighbourhood of 1) and the special values. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define __device__
#define __forceinline__ static inline
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }

#include "gvi_log_body.inc"

static double worst = 0.0, worst_x = 1.0;
static long count = 0;
static void check(double x) {
  long double ref = logl((long double)x);
  double got = gvi_log(x);
  double refd = (double)ref;
  double ulp = refd == 0.0 ? 4.9e-324 : fabs(nextafter(refd, INFINITY) - refd);
  if (refd != 0.0) {
    double lower = fabs(refd - nextafter(refd, -INFINITY));
    if (lower < ulp) ulp = lower;
  }
  double e = (double)(fabsl((long double)got - ref) / (long double)ulp);
  if (refd == 0.0) e = got == 0.0 ? 0.0 : 1e9;
  if (e > worst) { worst = e; worst_x = x; }
  count++;
}

int main(void) {
  for (long i = 0; i <= 3000000; i++) check(0.5 + 1.5 * (double)i / 3000000.0);            /* [0.5, 2] */
  for (int k = 1; k < 53; k++)                                                               /* 1 +- 2^-k (1 + u) */
    for (int j = 0; j < 20000; j++) {
      double d = ldexp(1.0 + (double)j / 20000.0, -k);
      check(1.0 + d);
      check(1.0 - 0.5 * d);
    }
  for (int e = -1022; e <= 1023; e++)                                                        /* every binade */
    for (int j = 0; j < 1500; j++) check(ldexp(1.0 + (double)j / 1500.0, e));
  uint64_t s = 88172645463325252ULL;                                                         /* random positive normals */
  for (int k = 0; k < 2000000; k++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    uint64_t bits = (s >> 1) % 0x7fe0000000000000ULL + 0x0010000000000000ULL;
    double x; memcpy(&x, &bits, 8);
    check(x);
  }
  /* every interval boundary of the table (bit patterns OFF + i 2^45, both sides) */
  for (int i = 0; i <= 128; i++)
    for (int d = -3; d <= 3; d++) {
      uint64_t bits = 0x3fe6100000000000ULL + ((uint64_t)i << 45) + (uint64_t)(int64_t)d;
      double x; memcpy(&x, &bits, 8);
      check(x);
    }
  printf("max_ulp %.4f at %.17g over %ld arguments\n", worst, worst_x, count);
  printf("one %d\n", gvi_log(1.0) == 0.0 && !signbit(gvi_log(1.0)));
  printf("zero %d\n", gvi_log(0.0) == -INFINITY && gvi_log(-0.0) == -INFINITY);
  printf("negative %d\n", isnan(gvi_log(-1.0)) && isnan(gvi_log(-1e-300)) && isnan(gvi_log(-INFINITY)));
  printf("inf %d\n", gvi_log(INFINITY) == INFINITY);
  printf("nan %d\n", isnan(gvi_log(NAN)));
  printf("subnormal %d\n", fabs(gvi_log(4.9e-324) - log(4.9e-324)) <= 1e-13 && gvi_log(2.2250738585072014e-308) == log(2.2250738585072014e-308));
  return 0;
}
