// K2: warp-split Poseidon leaf sponge for sm_100a.
// Replaces the per-leaf PoseidonHash::hash_or_noop loop of plonky2's MerkleTree::new, reached
// from PolynomialBatch::from_values inside starky::prover::prove
// (reference call sites /root/reference/src/aggregate_proof.rs:59,105,138,169,212).
//
// A leaf is one strictly sequential chain of ceil(C/8) permutations (C = 3k..97k columns) and there are only
// N = 2048..32768 leaves, so one thread per leaf cannot fill 148 SMs.  Here the 12-word state of 32 leaves is split
// across WPS warps of one block: warp w owns words [w*W, w*W+W), W = 12/WPS, of the 32 leaves held by its lanes
// (lane = leaf, so every column read is one coalesced 256-byte row of the column-major LDE).
//   * the S-box of a partial round runs only in the warp that owns word 0 -- the other warps skip it instead of
//     idling masked lanes, which is what made the lane-split variant issue-bound;
//   * per round the warps publish their words to a double-buffered shared tile (one __syncthreads per round) that is
//     stored twice ([24][32]) so that the row "w + i" of the circulant needs no modulo and every MDS coefficient
//     CIRC[i] is a compile-time immediate;
//   * the MDS layer splits words into 32-bit halves and accumulates 32x6-bit products with IMAD.WIDE, reducing once
//     per output word; the state stays lazy (any u64) between layers.
#pragma once
#include "poseidon.cuh"

// position (coset-major, see ntt.cu) -> plonky2 leaf index:  J*n + k  ->  J*n + bitrev_n(k)
__device__ __forceinline__ uint32_t leaf_index_of(uint32_t pos, unsigned log_block) {
  uint32_t mask = (1u << log_block) - 1;
  return (pos & ~mask) | bitrev32(pos & mask, log_block);
}

template <int WPS>
__global__ void __launch_bounds__(32 * WPS) leaf_sponge_ws_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                                  uint32_t n_leaves, unsigned log_block,
                                                                  u64* __restrict__ digests) {
  constexpr int W = 12 / WPS;
  static_assert(W * WPS == 12, "WPS must divide 12");
  __shared__ __align__(16) u64 xch[2][24][32];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned w0 = wid * W;
  const uint32_t pos_raw = blockIdx.x * 32 + lane;
  const bool live = pos_raw < n_leaves;
  const uint32_t pos = live ? pos_raw : n_leaves - 1;
  const u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

  u64 s[W], nx[W];
#pragma unroll
  for (int k = 0; k < W; k++) { s[k] = 0; nx[k] = 0; }
  const uint32_t n_chunks = (leaf_len + 7) / 8;
  auto fetch = [&](uint32_t chunk) {
#pragma unroll
    for (int k = 0; k < W; k++) {
      uint32_t c = chunk * 8 + w0 + k;
      if (w0 + k < 8 && c < leaf_len) nx[k] = cols[(size_t)c * n_leaves + pos];
    }
  };
  fetch(0);
  unsigned xb = 0;
  for (uint32_t m = 0; m < n_chunks; m++) {
    const unsigned take = min(8u, leaf_len - 8 * m);
#pragma unroll
    for (int k = 0; k < W; k++)
      if (w0 + k < take) s[k] = nx[k];
    if (m + 1 < n_chunks) fetch(m + 1);

    // Between layers s[k] holds the *pre-S-box* value of the coming round (MDS output + that round's constant):
    // the constant addition is free because it initialises the MDS half accumulators.
#pragma unroll
    for (int k = 0; k < W; k++) s[k] = gl_add_lazy_canon(s[k], c_poseidon_rc[w0 + k]);

    // publish v[k] (the S-box layer output of round rd) and form this warp's MDS rows + constants of round rd + 1
    auto linear_layer = [&](const u64 (&v)[W], int rd) {
      u64* dst = &xch[xb][w0][lane];
#pragma unroll
      for (int k = 0; k < W; k++) { dst[32 * k] = v[k]; dst[32 * (k + 12)] = v[k]; }
      u32 al0[W], al1[W], ah0[W], ah1[W];
#pragma unroll
      for (int k = 0; k < W; k++) {
        const u64 c = c_poseidon_rc[12 * (rd + 1) + w0 + k];
        al0[k] = (u32)c; al1[k] = 0; ah0[k] = (u32)(c >> 32); ah1[k] = 0;
      }
      __syncthreads();
      const u64* src = &xch[xb][w0][lane];
#pragma unroll
      for (int j = 0; j < W + 11; j++) {   // t = word (w0 + j) mod 12
        const u64 t = src[32 * j];
        const u32 lo = (u32)t, hi = (u32)(t >> 32);
#pragma unroll
        for (int k = 0; k < W; k++) {
          if (j - k >= 0 && j - k < 12) {
            mac32(al0[k], al1[k], lo, CIRC[j - k]);
            mac32(ah0[k], ah1[k], hi, CIRC[j - k]);
          }
        }
        if (j == 0 && wid == 0) { mac32(al0[0], al1[0], lo, 8u); mac32(ah0[0], ah1[0], hi, 8u); }   // DIAG[0] = 8
      }
#pragma unroll
      for (int k = 0; k < W; k++) s[k] = mds_recombine(al0[k], al1[k], ah0[k], ah1[k]);
      xb ^= 1;
    };
    auto full_round = [&](int rd) {
      u64 v[W];
#pragma unroll
      for (int k = 0; k < W; k++) v[k] = poseidon_sbox(s[k]);
      linear_layer(v, rd);
    };
#pragma unroll 1
    for (int rd = 0; rd < 4; rd++) full_round(rd);
#pragma unroll 1
    for (int rd = 4; rd < 26; rd++) {
      u64 v[W];
#pragma unroll
      for (int k = 0; k < W; k++) v[k] = s[k];
      if (wid == 0) v[0] = poseidon_sbox(v[0]);
      linear_layer(v, rd);
    }
#pragma unroll 1
    for (int rd = 26; rd < 30; rd++) full_round(rd);
  }
  if (live) {
    u64* d = digests + 4ull * leaf_index_of(pos, log_block);
#pragma unroll
    for (int k = 0; k < W; k++)
      if (w0 + k < 4) d[w0 + k] = gl_canon(s[k]);
  }
}
