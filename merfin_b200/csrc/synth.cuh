// Counter-hash random numbers of the synthetic read databases (table.cu, joint_stream.cu): the same draws
// whether the database is built as a resident table or streamed through the semi-join.
#pragma once
#include <cstdint>

// splitmix-style counter hash -> uniform in [0,1)
__device__ __forceinline__ uint64_t mfn_rng(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// Poisson(lambda) by inversion for small lambda, rounded normal (Box-Muller) above 64.
static __device__ uint32_t mfn_poisson(double lambda, uint64_t key, uint64_t seed) {
  uint64_t s = mfn_rng(key ^ seed);
  if (lambda < 64.0) {
    double u = (double)(s >> 11) * (1.0 / 9007199254740992.0);
    double p = exp(-lambda), cdf = p;
    uint32_t x = 0;
    while (u > cdf && x < 1000) { x++; p *= lambda / x; cdf += p; }
    return x;
  }
  double u1 = ((double)(s >> 11) + 1.0) * (1.0 / 9007199254740993.0);
  double u2 = (double)(mfn_rng(s) >> 11) * (1.0 / 9007199254740992.0);
  double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  double v = lambda + sqrt(lambda) * z + 0.5;
  return v < 0.0 ? 0u : (uint32_t)v;
}

