// Philox4x32-10 counter-based generator and the canonical dropout mask of the GRU layers (nn.Dropout at reference
// models.py:246/276/700, training mode).  The keep decision of element (b, t, col) of a [B][T][256] layer output is
//     16-bit draw (t & 7) of philox4x32_10(counter = ((b*256 + col) << 32) | (t >> 3), key = seed)  <  keep_threshold16
// (draw k = bits 16*(k&1).. of word k>>1; P(keep) is quantised to 2^-16), so a thread that walks t for a fixed (b, col) -- as
// the persistent GRU kernels do, forwards or backwards -- needs one Philox call per 8 steps, and the backward kernel regenerates
// exactly the forward's mask from (seed, p) with no mask tensor in HBM.
#pragma once
#include <stdint.h>

__device__ __forceinline__ void slu_philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  c[0] = hi1 ^ c[1] ^ k0;
  c[1] = lo1;
  c[2] = hi0 ^ c[3] ^ k1;
  c[3] = lo0;
}

__device__ __forceinline__ void slu_philox4x32_10(uint64_t ctr, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    slu_philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

// the eight 16-bit draws (4 words) of time group tg = t >> 3 for element (b, col)
__device__ __forceinline__ void slu_gru_mask_draws(int b, int col, int tg, uint64_t seed, uint32_t (&out)[4]) {
  slu_philox4x32_10(((uint64_t)(uint32_t)(b * 256 + col) << 32) | (uint32_t)tg, seed, out);
}
// draw k = t & 7 of a group, selected WITHOUT dynamic register-array indexing (that would put the array in local memory)
__device__ __forceinline__ uint32_t slu_gru_mask_draw16(const uint32_t (&w)[4], int t) {
  const int k = t & 7;
  const uint32_t word = k < 4 ? (k < 2 ? w[0] : w[1]) : (k < 6 ? w[2] : w[3]);
  return (word >> (16 * (k & 1))) & 0xffffu;
}

// P(keep) = threshold / 2^32 (32-bit draws: slu_dropout_mask)
static inline uint32_t slu_keep_threshold(float p) {
  const double th = (1.0 - (double)p) * 4294967296.0;
  return th >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)th;
}
// P(keep) = threshold / 2^16 (16-bit draws: the GRU-layer mask); 0 is reserved for "no dropout", so p -> 1 keeps 1 / 65536
static inline uint32_t slu_keep_threshold16(float p) {
  const double th = (1.0 - (double)p) * 65536.0 + 0.5;
  const uint32_t t = th >= 65536.0 ? 65536u : (uint32_t)th;
  return t == 0u ? 1u : t;
}
