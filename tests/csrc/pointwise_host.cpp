// TEST INFRASTRUCTURE.  Compiles the DEVICE source of the pointwise loss terms (gvi_nll_term, gvi_nll_grad,
// gvi_reg_term, gvi_reg_grad: the text of csrc/common.cuh pasted in by tests/test_host_logic.py as
// pointwise_body.inc) for the host, so the formulas the kernels execute can be held against the reference's golden
// values and their hand-derived derivatives against central differences without a GPU.
// stdin: one "kind alpha y mp cp mq cq" per line; stdout: "nll dnll_dm dnll_dv d0 dd0_dmq dd0_dcq" per line.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "gvigp.h"
#define __device__
#define __forceinline__ inline
using std::exp;
using std::fma;
using std::log;
using std::sqrt;
static inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; std::memcpy(&d, &u, 8); return d; }

#include "pointwise_body.inc"

int main() {
  int kind;
  double alpha, y, mp, cp, mq, cq;
  while (std::scanf("%d %lf %lf %lf %lf %lf %lf", &kind, &alpha, &y, &mp, &cp, &mq, &cq) == 7) {
    if (kind == -1) {  // "-1 0 x 0 0 0 0": gvi_log(x) alone (17 significant digits + the bit pattern)
      const double v = gvi_log(y);
      std::printf("%.17g %lld 0 0 0 0\n", v, __double_as_longlong(v));
      continue;
    }
    double dm, dv, dmq, dcq;
    gvi_nll_grad(y, mq, cq, dm, dv);
    gvi_reg_grad(kind, alpha, mp, cp, mq, cq, dmq, dcq);
    std::printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", gvi_nll_term(y, mq, cq), dm, dv,
                gvi_reg_term(kind, alpha, mp, cp, mq, cq), dmq, dcq);
  }
  return 0;
}
