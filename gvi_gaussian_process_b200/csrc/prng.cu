// prng.cu - the counter-mode hash behind jax.random (threefry2x32, jax 0.4.13 with jax_threefry_partitionable=False; an
// un-vendored third-party dependency of the reference, pinned in its poetry.lock:1078-1079) for the permutation of
// ConditionalVarianceInducingPointsSelector.compute_inducing_points (src/inducing_points_selection.py:115-119).
// random_bits(key, 32, (n,)) = threefry_2x32(key, arange(n)): the counter array (padded to even length with a 0) is
// split into halves x0 = counts[:h], x1 = counts[h:], hashed pairwise, and the result is concat(x0', x1')[:n].
// At N = 1e7 the host restatement (NumPy) of two such rounds plus two stable argsorts costs seconds; on the device the
// hash is one elementwise launch and the stable key sort is the host framework's radix sort.
#include "common.cuh"

namespace {
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__global__ void __launch_bounds__(256) threefry_bits_kernel(uint32_t k0, uint32_t k1, int64_t n, int64_t h,
                                                            int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= h) return;
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t a = (uint32_t)i + ks[0];
  uint32_t b = ((h + i < n) ? (uint32_t)(h + i) : 0u) + ks[1];
  const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
#pragma unroll
  for (int g = 0; g < 5; g++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      a += b;
      b = rotl32(b, rot[g & 1][r]);
      b ^= a;
    }
    a += ks[(g + 1) % 3];
    b += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  out[i] = (int64_t)a;
  if (h + i < n) out[h + i] = (int64_t)b;
}
}  // namespace

extern "C" int gvi_threefry_random_bits(uint32_t key0, uint32_t key1, int64_t n, int64_t* out, void* stream) {
  GVI_CHECK_ARG(n >= 0 && n < (1ll << 32), "n");
  GVI_CHECK_ARG(out || n == 0, "null pointer");
  if (n == 0) return GVI_OK;
  const int64_t h = (n + 1) / 2;
  threefry_bits_kernel<<<(unsigned)((h + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key0, key1, n, h, out);
  GVI_CHECK_LAUNCH("threefry_bits_kernel");
  return GVI_OK;
}
