// Philox4x32-10 per-packet streams (replaces random_numbers.f90 / SPRNG, see
// include/mcfost_b200.h: mcb_run_params.seed).  Stream definition (identical in
// oracle/philox.h):
//   key = (seed_lo, seed_hi); counter = (block, packet_lo, packet_hi, call_index)
//   packet = (chunk-1) * 2^40 + index_in_chunk; each block -> four draws
//   rand = (w >> 8) * 2^-24 (fp32-exact, in [0,1)) -- every sprng() call site on
//   this path assigns to a Fortran `real`.  Blocks are assigned per EVENT:
//       emission                                   blocks 0, 1
//       flight e = 1, 2, ... (tau, interaction-type draw)      block 2e
//       interaction ending flight e (angles / re-emission)     block 2e + 1
//   so a packet's RNG state is (packet id, event counter): 12 bytes of shared
//   memory, and any thread can continue any packet after a regrouping step.
#pragma once
#include <cstdint>

namespace mcb {

__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

// one Philox4x32-10 block; deliberately not inlined (it is called from every phase and
// the unrolled rounds would otherwise be duplicated ~10x in the instruction stream)
__device__ __noinline__ uint4 philox_block(uint32_t k0, uint32_t k1, uint32_t blk, uint32_t c1, uint32_t c2, uint32_t c3) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t a0 = blk, a1 = c1, a2 = c2, a3 = c3;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, a0), lo0 = M0 * a0;
    const uint32_t hi1 = __umulhi(M1, a2), lo1 = M1 * a2;
    const uint32_t n0 = hi1 ^ a1 ^ k0, n2 = hi0 ^ a3 ^ k1;
    a0 = n0; a1 = lo1; a2 = n2; a3 = lo0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(a0, a1, a2, a3);
}

// two blocks of the same packet at once: the two chains are independent, so the scheduler interleaves
// them (the interaction block of flight e and the flight block of e+1 are always needed together)
__device__ __noinline__ void philox_block2(uint32_t k0, uint32_t k1, uint32_t blkA, uint32_t blkB, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint4& outA, uint4& outB) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t a0 = blkA, a1 = c1, a2 = c2, a3 = c3, b0 = blkB, b1 = c1, b2 = c2, b3 = c3;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t ah0 = __umulhi(M0, a0), al0 = M0 * a0, ah1 = __umulhi(M1, a2), al1 = M1 * a2;
    const uint32_t bh0 = __umulhi(M0, b0), bl0 = M0 * b0, bh1 = __umulhi(M1, b2), bl1 = M1 * b2;
    const uint32_t an0 = ah1 ^ a1 ^ k0, an2 = ah0 ^ a3 ^ k1, bn0 = bh1 ^ b1 ^ k0, bn2 = bh0 ^ b3 ^ k1;
    a0 = an0; a1 = al1; a2 = an2; a3 = al0;
    b0 = bn0; b1 = bl1; b2 = bn2; b3 = bl0;
    k0 += W0; k1 += W1;
  }
  outA = make_uint4(a0, a1, a2, a3); outB = make_uint4(b0, b1, b2, b3);
}

}  // namespace mcb
