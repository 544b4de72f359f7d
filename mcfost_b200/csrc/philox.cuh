// Philox4x32-10 per-packet streams (replaces random_numbers.f90 / SPRNG, see
// include/mcfost_b200.h: mcb_run_params.seed).  Stream definition:
//   key = (seed_lo, seed_hi); counter = (block, packet_lo, packet_hi, call_index)
//   packet = (chunk-1) * 2^40 + index_in_chunk; each block -> four draws
//   rand = (w >> 8) * 2^-24 (fp32-exact, in [0,1)), consumed in order w0..w3 --
//   every sprng() call site on this path assigns to a Fortran `real`.  A packet's
//   draw sequence is a pure function of (seed, call_index, packet), independent
//   of thread / GPU count.
#pragma once
#include <cstdint>

namespace mcb {

struct Rng {
  uint32_t k0, k1;        // key
  uint32_t c1, c2, c3;    // packet_lo, packet_hi, call_index
  uint32_t blk;           // next block index
  uint32_t w1, w2, w3;    // unread words of the current block
  int      left;          // how many of w1..w3 are unread

  __device__ __forceinline__ void seed(uint64_t s, uint32_t call_index, uint64_t packet) {
    k0 = (uint32_t)s; k1 = (uint32_t)(s >> 32);
    c1 = (uint32_t)packet; c2 = (uint32_t)(packet >> 32); c3 = call_index;
    blk = 0; left = 0; w1 = w2 = w3 = 0;
  }

  // `rand = sprng(stream(id))` with rand a Fortran `real`
  __device__ __forceinline__ float nextf() {
    uint32_t w;
    if (left > 0) {
      w = w1; w1 = w2; w2 = w3; --left;
    } else {
      const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
      uint32_t a0 = blk, a1 = c1, a2 = c2, a3 = c3, ka = k0, kb = k1;
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, a0), lo0 = M0 * a0;
        uint32_t hi1 = __umulhi(M1, a2), lo1 = M1 * a2;
        uint32_t n0 = hi1 ^ a1 ^ ka, n2 = hi0 ^ a3 ^ kb;
        a0 = n0; a1 = lo1; a2 = n2; a3 = lo0;
        ka += W0; kb += W1;
      }
      ++blk;
      w = a0; w1 = a1; w2 = a2; w3 = a3; left = 3;
    }
    return (float)(w >> 8) * (1.0f / 16777216.0f);
  }
};

}  // namespace mcb
