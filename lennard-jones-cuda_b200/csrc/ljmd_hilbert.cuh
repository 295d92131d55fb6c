// Hilbert-curve index of a cell (x, y, z), `bits` bits per axis (Skilling's transpose algorithm, AIP Conf. Proc.
// 707 (2004) 381): consecutive indices are face-adjacent cells, so a run of particles sorted by it is a compact
// cloud — which is what the warp frames of the Newton-3 kernel and the RDF box pruning need.  Host and device.
#pragma once
#include <stdint.h>

namespace ljmd {

#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t hilbert3(uint32_t x, uint32_t y, uint32_t z, int bits) {
  uint32_t X[3] = {x, y, z};
  const uint32_t M = 1u << (bits - 1);
  // inverse undo excess work
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    const uint32_t P = Q - 1;
    for (int i = 0; i < 3; ++i) {
      if (X[i] & Q) {
        X[0] ^= P;
      } else {
        const uint32_t t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
  // Gray encode
  for (int i = 1; i < 3; ++i) X[i] ^= X[i - 1];
  uint32_t t = 0;
  for (uint32_t Q = M; Q > 1; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1;
  for (int i = 0; i < 3; ++i) X[i] ^= t;
  // interleave: bit b of X[0] is the most significant of its triple
  uint32_t key = 0;
  for (int b = bits - 1; b >= 0; --b)
    for (int i = 0; i < 3; ++i) key = (key << 1) | ((X[i] >> b) & 1u);
  return key;
}

}  // namespace ljmd
