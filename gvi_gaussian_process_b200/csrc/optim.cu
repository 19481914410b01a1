// optim.cu - SURVEY.md section 8(f)-4: device-resident optimiser step for the training loops of
// experiments/shared/trainer.py:42-100.  The reference resolves optax (third-party, 0.1.7 pinned in poetry.lock)
// transformations (experiments/shared/resolvers/optimiser.py:6-16: adam / adabelief / rmsprop with optax defaults);
// their published update rules are restated here as ONE elementwise kernel over the flat parameter vector, so that
// parameters, gradients and optimiser state never leave HBM and a step costs one launch (40 B / parameter of traffic).
//   adam      : mu = b1 mu + (1-b1) g;  nu = b2 nu + (1-b2) g^2;  p -= lr mu^ / (sqrt(nu^ + eps_root) + eps)
//   adabelief : mu as above;  nu = b2 nu + (1-b2) (g - mu_prev)^2 + eps_root;  p -= lr mu^ / (sqrt(nu^) + eps)
//   rmsprop   : nu = b2 nu + (1-b2) g^2;  p -= lr g / sqrt(nu + eps)                      (b2 = decay)
// with the bias corrections mu^ = mu / (1 - b1^t), nu^ = nu / (1 - b2^t), t = 1-based step count.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) optimiser_step_kernel(int kind, double* __restrict__ p,
                                                             const double* __restrict__ g, double* __restrict__ mu,
                                                             double* __restrict__ nu, int64_t n, double lr, double b1,
                                                             double b2, double eps, double eps_root, double c1,
                                                             double c2) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double gi = g[i];
    double step;
    if (kind == GVI_OPT_RMSPROP) {
      const double v = b2 * nu[i] + (1.0 - b2) * gi * gi;
      nu[i] = v;
      step = gi * rsqrt(v + eps);
    } else {
      const double m_prev = mu[i];
      const double m = b1 * m_prev + (1.0 - b1) * gi;
      mu[i] = m;
      double v;
      if (kind == GVI_OPT_ADABELIEF) {
        const double e = gi - m_prev;
        v = b2 * nu[i] + (1.0 - b2) * e * e + eps_root;
        nu[i] = v;
        step = (m / c1) / (sqrt(v / c2) + eps);
      } else {
        v = b2 * nu[i] + (1.0 - b2) * gi * gi;
        nu[i] = v;
        step = (m / c1) / (sqrt(v / c2 + eps_root) + eps);
      }
    }
    p[i] = fma(-lr, step, p[i]);
  }
}
}  // namespace

extern "C" int gvi_optimiser_step_f64(int kind, double* param, const double* grad, double* mu, double* nu, int64_t n,
                                      int64_t count, double learning_rate, double b1, double b2, double eps,
                                      double eps_root, void* stream) {
  GVI_CHECK_ARG(kind >= GVI_OPT_ADAM && kind <= GVI_OPT_RMSPROP, "optimiser kind");
  GVI_CHECK_ARG(n >= 0 && count >= 1, "sizes");
  if (n == 0) return GVI_OK;
  GVI_CHECK_ARG(param && grad && nu && (mu || kind == GVI_OPT_RMSPROP), "null pointer");
  const double c1 = 1.0 - pow(b1, (double)count), c2 = 1.0 - pow(b2, (double)count);
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)(want < GVI_NUM_SMS * 8 ? want : GVI_NUM_SMS * 8);
  optimiser_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kind, param, grad, mu, nu, n, learning_rate, b1, b2,
                                                                  eps, eps_root, c1, c2);
  GVI_CHECK_LAUNCH("optimiser_step_kernel");
  return GVI_OK;
}
