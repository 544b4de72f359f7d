/*
 * ORACLE (test infrastructure only -- never linked into the product path).
 *
 * Philox4x32-10 counter-based RNG (Salmon et al., SC'11, "Parallel random
 * numbers: as easy as 1, 2, 3"; Random123 v1.09).  Replaces the reference's
 * SPRNG 2.0b 64-bit LCG streams (random_numbers.f90:19-30, one stream per
 * OpenMP thread, dust_transfer.f90:125-128): sequence parity with SPRNG is a
 * stated non-goal (BASELINE.json north_star), so the oracle and the CUDA path
 * both use this generator, keyed per packet, which makes packet trajectories
 * reproducible and independent of thread / GPU count.
 *
 * Stream definition (shared with mcfost_b200/csrc/philox.cuh):
 *   key     = (seed_lo, seed_hi)
 *   counter = (block, packet_lo, packet_hi, call_index)
 *   packet  = (chunk-1) * 2^40 + index_in_chunk (0-based)
 *   each 128-bit block yields four uniform draws, consumed in order w0..w3:
 *       rand = (w >> 8) * 2^-24          (exactly representable in fp32, in [0,1))
 *   Blocks are assigned per EVENT so that a packet's RNG state is just (packet,
 *   event counter) -- what the GPU keeps per packet in shared memory:
 *       emission (wavelength, source, position, direction)   blocks 0, 1
 *       flight e = 1, 2, ...   (tau, interaction-type draw)    block  2e
 *       interaction ending flight e (angles / re-emission)    block  2e + 1
 *   Unused words of a block are discarded.
 *   Every sprng() call site on this path assigns the result to a Fortran `real`
 *   (fp32), so one 24-bit draw per call carries all the bits the reference keeps
 *   near 1; the one difference is that rand can never round up to exactly 1.0
 *   (the reference guards that case explicitly, dust_transfer.f90:1209).
 */
#ifndef ORACLE_PHILOX_H
#define ORACLE_PHILOX_H
#include <stdint.h>

static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct PacketRng {
  /* recorded stream (deterministic unit tests): if rec != 0, values are
     replayed cyclically instead of generated */
  const double *rec;
  int64_t n_rec, i_rec;
  uint32_t key[2];
  uint32_t ctr[4];     /* ctr[0] = next block index */
  uint32_t buf[4];
  int      n_buf;      /* unread words left in buf */
  int64_t  n_draws;
} PacketRng;

static inline void rng_seed_packet(PacketRng *g, uint64_t seed, uint32_t call_index, uint64_t packet) {
  g->key[0] = (uint32_t)seed; g->key[1] = (uint32_t)(seed >> 32);
  g->ctr[0] = 0; g->ctr[1] = (uint32_t)packet; g->ctr[2] = (uint32_t)(packet >> 32); g->ctr[3] = call_index;
  g->n_buf = 0; g->n_draws = 0;
}

/* jump to the first word of a given block (start of an event) */
static inline void rng_set_block(PacketRng *g, uint32_t block) {
  if (g->rec) return;
  g->ctr[0] = block; g->n_buf = 0;
}

/* the analogue of sprng(stream(id)): a value in [0,1) (24-bit resolution) */
static inline double rng_next(PacketRng *g) {
  g->n_draws++;
  if (g->rec) { double v = g->rec[g->i_rec % g->n_rec]; g->i_rec++; return v; }
  if (g->n_buf == 0) {
    philox4x32_10(g->ctr, g->key, g->buf);
    g->ctr[0]++;
    g->n_buf = 4;
  }
  uint32_t w = g->buf[4 - g->n_buf];
  g->n_buf--;
  return (double)(w >> 8) * (1.0 / 16777216.0);
}
#endif
