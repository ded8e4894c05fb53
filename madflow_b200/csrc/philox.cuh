// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw, SC'11).
// The reference draws its uniforms with TensorFlow's stateful generator inside vegasflow
// (utilities.py:89 sets the seed); a counter-based stream replaces it so that every event's
// random point depends only on (seed, iteration, global event index) -- never on the GPU count,
// the chunking or the launch shape.
//   key = (seed lo, seed hi); counter = (event lo, event hi, iteration, j); block j -> dims 2j, 2j+1
#pragma once
#include "mf_complex.cuh"

namespace mf {

struct U4 {
  uint32_t x, y, z, w;
};

MF_DEV uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

MF_DEV U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// 53 random bits -> [0,1)
MF_DEV double u01(uint32_t a, uint32_t b) {
  const uint64_t v = ((uint64_t)a << 32) | (uint64_t)b;
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// two uniforms (dimensions 2j and 2j+1) of one event
MF_DEV void philox_pair(uint64_t seed, uint32_t iteration, uint64_t event, uint32_t j, double& u0, double& u1) {
  const U4 r = philox4x32_10(U4{(uint32_t)event, (uint32_t)(event >> 32), iteration, j}, (uint32_t)seed,
                             (uint32_t)(seed >> 32));
  u0 = u01(r.x, r.y);
  u1 = u01(r.z, r.w);
}

}  // namespace mf
