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

// ---------------------------------------------------------------------------------------------------------
// WPS = 12 specialisation (one word per warp) for the latency-bound shapes (N = 2048..4096 leaves: every leaf is in
// flight at once, so wall time = permutations per leaf x time of one permutation of a 32-leaf group).
// The 22 partial rounds are the critical path: word 0 goes through x^7 every round while the other eleven words are
// only mixed.  Two named barriers per partial round split the MDS layer so that only the word-0 term waits for the
// S-box:
//   A: words 1..11 (known as soon as the previous layer ends) are published; every warp accumulates its eleven-term
//      partial row while warp 0 is still inside the S-box (its own row's partial sum interleaves with the S-box);
//   B: warp 0 publishes y0 = x0^7 and only *arrives*; the others wait, add c_j0 * y0 and reduce.
// Row 0 of the exchange tile is kept zero during partial rounds so that all twelve row reads keep their immediate
// coefficients.
// ---------------------------------------------------------------------------------------------------------
// Row of the MDS layer: sum_i c[i] * t[i] + rc as half accumulators, CH IMAD.WIDE chains per half.  (Measured on B200:
// a dependent IMAD.WIDE accumulate costs ~9 cycles, yet CH = 2, 3 were not faster than CH = 1 -- the exchange through
// shared memory, 144 LDS.64 per round and 32-leaf group = 40 % of the LSU data pipe, is what the round waits for.)
template <int CH, class Coef>
__device__ __forceinline__ void mds_row(const u64 (&t)[12], Coef coef, u64 rc, int k0, u32& al0, u32& al1, u32& ah0, u32& ah1) {
  u32 l0[CH], l1[CH], h0[CH], h1[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { l0[c] = 0; l1[c] = 0; h0[c] = 0; h1[c] = 0; }
  l0[0] = (u32)rc; h0[0] = (u32)(rc >> 32);
#pragma unroll
  for (int k = 0; k < 12; k++) {
    if (k >= k0) {
      mac32(l0[k % CH], l1[k % CH], (u32)t[k], coef(k));
      mac32(h0[k % CH], h1[k % CH], (u32)(t[k] >> 32), coef(k));
    }
  }
  u64 al = ((u64)l1[0] << 32) | l0[0], ah = ((u64)h1[0] << 32) | h0[0];
#pragma unroll
  for (int c = 1; c < CH; c++) { al += ((u64)l1[c] << 32) | l0[c]; ah += ((u64)h1[c] << 32) | h0[c]; }
  al0 = (u32)al; al1 = (u32)(al >> 32); ah0 = (u32)ah; ah1 = (u32)(ah >> 32);
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

static __constant__ u32 c_poseidon_circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

// LAYOUT 0: 12 warps, word = warp id.   LAYOUT 1: 12 warps, word = 11 - warp id (the arbiter favours high warp ids).
// LAYOUT 2: 16 warps, word 0 on warp 11 alone in its SM sub-partition (warps 3, 7, 15 and 14 exit at once).
template <int LAYOUT, int CH = 1>
__global__ void __launch_bounds__(LAYOUT == 2 ? 512 : 384) leaf_sponge_w12_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                              uint32_t n_leaves, unsigned log_block,
                                                              u64* __restrict__ digests) {
  __shared__ __align__(16) u64 xch[2][24][32];
  __shared__ __align__(16) u64 ysl[2][32];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned wid;   // the state word this warp owns
  if (LAYOUT == 0) wid = warp;
  else if (LAYOUT == 1) wid = 11 - warp;
  else {
    if (((warp & 3) == 3 && warp != 11) || warp == 14) return;
    wid = warp == 11 ? 0 : warp - (warp >> 2) + 1;
  }
  const uint32_t pos_raw = blockIdx.x * 32 + lane;
  const bool live = pos_raw < n_leaves;
  const uint32_t pos = live ? pos_raw : n_leaves - 1;
  const u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  const u32 c_y0 = c_poseidon_circ[(12 - wid) % 12] + (wid == 0 ? 8u : 0u);   // coefficient of word 0 in row `wid`

  u64 s = 0, nx = 0;
  const uint32_t n_chunks = (leaf_len + 7) / 8;
  if (wid < 8 && wid < leaf_len) nx = cols[(size_t)wid * n_leaves + pos];
  unsigned xb = 0;

  // twelve-term row: acc = rc(next round) + sum_i CIRC[i] * word(wid + i)
  auto row_accumulate = [&](int rd, u32& al0, u32& al1, u32& ah0, u32& ah1) {
    const u64 c = c_poseidon_rc[12 * (rd + 1) + wid];
    u64 t[12];
#pragma unroll
    for (int i = 0; i < 12; i++) t[i] = xch[xb][wid + i][lane];
    mds_row<CH>(t, [&](int i) { return CIRC[i]; }, c, 0, al0, al1, ah0, ah1);
  };

  for (uint32_t m = 0; m < n_chunks; m++) {
    const unsigned take = min(8u, leaf_len - 8 * m);
    if (wid < take) s = nx;
    if (m + 1 < n_chunks) {
      const uint32_t c = (m + 1) * 8 + wid;
      if (wid < 8 && c < leaf_len) nx = cols[(size_t)c * n_leaves + pos];
    }
    s = gl_add_lazy_canon(s, c_poseidon_rc[wid]);

    auto full_round = [&](int rd) {
      const u64 v = poseidon_sbox(s);
      xch[xb][wid][lane] = v; xch[xb][wid + 12][lane] = v;
      named_bar_sync(1, 384);
      u32 al0, al1, ah0, ah1;
      row_accumulate(rd, al0, al1, ah0, ah1);
      if (wid == 0) {   // DIAG[0] = 8
        const u64 t = xch[xb][0][lane];
        mac32(al0, al1, (u32)t, 8u); mac32(ah0, ah1, (u32)(t >> 32), 8u);
      }
      s = mds_recombine(al0, al1, ah0, ah1);
      xb ^= 1;
    };
#pragma unroll 1
    for (int rd = 0; rd < 4; rd++) full_round(rd);
#pragma unroll 1
    for (int rd = 4; rd < 26; rd++) {
      u32 al0, al1, ah0, ah1;
      u64 y0;
      if (wid == 0) {
        xch[xb][0][lane] = 0; xch[xb][12][lane] = 0;
        const u64 x2 = gl_mul_lazy(s, s);                 // first S-box level runs before the others have published
        named_bar_sync(1, 384);
        row_accumulate(rd, al0, al1, ah0, ah1);           // interleaves with the rest of the S-box
        const u64 x4 = gl_mul_lazy(x2, x2), x3 = gl_mul_lazy(x2, s);
        y0 = gl_mul_lazy(x3, x4);
        ysl[xb][lane] = y0;
        named_bar_arrive(2, 384);
      } else {
        xch[xb][wid][lane] = s; xch[xb][wid + 12][lane] = s;
        named_bar_sync(1, 384);
        row_accumulate(rd, al0, al1, ah0, ah1);
        named_bar_sync(2, 384);
        y0 = ysl[xb][lane];
      }
      mac32(al0, al1, (u32)y0, c_y0);
      mac32(ah0, ah1, (u32)(y0 >> 32), c_y0);
      s = mds_recombine(al0, al1, ah0, ah1);
      xb ^= 1;
    }
#pragma unroll 1
    for (int rd = 26; rd < 30; rd++) full_round(rd);
  }
  if (live && wid < 4) digests[4ull * leaf_index_of(pos, log_block) + wid] = gl_canon(s);
}

// ---------------------------------------------------------------------------------------------------------
// Sparse partial rounds (poseidon_fast.h, derived by tools/gen_poseidon_fast.py from MDS + round constants; the
// factorisation plonky2 ships as FAST_PARTIAL_*).  In the kernel above a partial round is an all-to-all through shared
// memory (144 LDS.64 per 32-leaf group and round, 40 % of the LSU pipe) and costs ~700 cycles.  In sparse form a partial
// round is   y = x0^7 + a_r;   x0' = 25 y + sum_i w^_r[i] x_i;   x_i' = x_i + v_r[i] y   -- one broadcast of y and one
// reduction, and with  sum_i w^_r[i] x_i(r) = sum_i w^_r[i] x_i(r-1) + U[r] y(r-1)  the reduction is known a round ahead,
// so the warp that owns word 0 never waits for the others:
//   13 warps per 32 leaves.  warps 0..11 own one word each (word = (warp + 9) % 12, so word 0 sits on warp 3), warp 12
//   ("R") reduces.  Round r, slot q = r % 3 (triple-buffered slots and barrier ids: every reuse is ordered by a chain of
//   barrier completions):
//     word 0 : y = x0^7 + a_r -> ybuf[q]; wait F_q; arrive B_q and Y_q; x0 = 25 y + Ebuf[q]
//     word i : wait B_q; x_i += v_r[i] y; e = w^_{r+2}[i] x_i -> ebuf[(r+2) % 3][i]; arrive G_{(r+2) % 3}
//     R      : wait G_q; s = sum_i ebuf[q][i]; wait Y_{(r-1) % 3}; Ebuf[q] = s + U[r] y(r-1) (as two half sums); arrive F_q
//   (letting word 0's warp multiply U[r] y(r-1) itself removes R from its loop but measured slower: that warp is
//   issue-bound, 21 more instructions cost more than the wait)
//   The dense 11x11 "initial" matrix of the sparse form is folded into the linear layer of full round 3 (D3ROT, K3).
// Barrier ids: A = 1 (full rounds, 12 warps), B = 2..4, G = 5..7, F = 8..10, Y = 11..13.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 dp2a_lo(u32 a, u32 b, u32 c) {
  u32 d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// limb accumulators (each < 2^28) -> lazy u64:  a0 + a1 2^16 + a2 2^32 + a3 2^48
__device__ __forceinline__ u64 limbs_recombine(u32 a0, u32 a1, u32 a2, u32 a3) {
  const u64 al = (u64)a0 + ((u64)a1 << 16), ah = (u64)a2 + ((u64)a3 << 16);     // both < 2^45
  return mds_recombine((u32)al, (u32)(al >> 32), (u32)ah, (u32)(ah >> 32));
}

// one MDS row on dp2a for a thread that holds the twelve words rotated (t[i] = word (row + i) mod 12): six word pairs x
// four 16-bit limbs, coefficients (CIRC[2p], CIRC[2p+1]) as immediates; `diag0`: row 0 adds 8 * word 0
__device__ __forceinline__ u64 mds_row_dp2a(const u64 (&t)[12], u64 rc, bool diag0) {
  constexpr u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  const u32 lo = (u32)rc, hi = (u32)(rc >> 32);
  u32 a0 = lo & 0xFFFFu, a1 = lo >> 16, a2 = hi & 0xFFFFu, a3 = hi >> 16;
#pragma unroll
  for (int p = 0; p < 6; p++) {
    const u32 alo = (u32)t[2 * p], ahi = (u32)(t[2 * p] >> 32), blo = (u32)t[2 * p + 1], bhi = (u32)(t[2 * p + 1] >> 32);
    const u32 q0 = __byte_perm(alo, blo, 0x5410), q1 = __byte_perm(alo, blo, 0x7632);
    const u32 q2 = __byte_perm(ahi, bhi, 0x5410), q3 = __byte_perm(ahi, bhi, 0x7632);
    const u32 b = CIRC[2 * p] | (CIRC[2 * p + 1] << 8);
    a0 = dp2a_lo(q0, b, a0); a1 = dp2a_lo(q1, b, a1); a2 = dp2a_lo(q2, b, a2); a3 = dp2a_lo(q3, b, a3);
    if (p == 0 && diag0) { a0 = dp2a_lo(q0, 8u, a0); a1 = dp2a_lo(q1, 8u, a1); a2 = dp2a_lo(q2, 8u, a2); a3 = dp2a_lo(q3, 8u, a3); }
  }
  return limbs_recombine(a0, a1, a2, a3);
}

#include <type_traits>

#include "poseidon_fast.h"
// sum of 64x64 products as lo + hi 2^64 + top 2^128
struct Acc192 { u64 lo, hi; u32 top; };
__device__ __forceinline__ void acc192_mul(Acc192& a, u64 x, u64 y) {
  const u64 pl = x * y, ph = __umul64hi(x, y);
  asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
      : "+l"(a.lo), "+l"(a.hi), "+r"(a.top) : "l"(pl), "l"(ph));
}
// (lo, hi, top) mod p -> lazy u64.   2^128 = -2^32 (mod p):  top * 2^128 = top * (p - 2^32) = top * (2^64 - 2^33 + 1)
__device__ __forceinline__ u64 acc192_reduce(const Acc192& a) {
  u64 lo = a.lo, hi = a.hi;
  // t = top * (2^64 - 2^33 + 1) = (top - borrow) : (top - top 2^33), added to hi:lo; a wrap of 2^128 (at most one) is again
  // worth 2^64 - 2^33 + 1, and after a wrap hi is tiny, so the second addition cannot wrap
  asm("{\n\t"
      ".reg .u64 t, sb, tlo, thi, c2, s2, ulo, uhi;\n\t"
      "cvt.u64.u32 t, %2;\n\t"
      "shl.b64 sb, t, 33;\n\t"
      "sub.cc.u64 tlo, t, sb;\n\t"
      "subc.u64 thi, t, 0;\n\t"
      "add.cc.u64 %0, %0, tlo;\n\t"
      "addc.cc.u64 %1, %1, thi;\n\t"
      "addc.u64 c2, 0, 0;\n\t"
      "shl.b64 s2, c2, 33;\n\t"
      "sub.cc.u64 ulo, c2, s2;\n\t"
      "subc.u64 uhi, c2, 0;\n\t"
      "add.cc.u64 %0, %0, ulo;\n\t"
      "addc.u64 %1, %1, uhi;\n\t"
      "}"
      : "+l"(lo), "+l"(hi) : "r"(a.top));
  // fold the 128-bit value (x3:x2:x1:x0) exactly like gl_mul_lazy
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 nc, nb;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\t"
      "subc.u32 %1, %4, 0;\n\t"
      "add.cc.u32 %1, %1, %3;\n\t"
      "addc.u32 nc, 0, 0;\n\t"
      "neg.s32 nc, nc;\n\t"
      "sub.cc.u32 %0, %0, %5;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 nb, 0, 0;\n\t"
      "add.cc.u32 %0, %0, nc;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "sub.cc.u32 %0, %0, nb;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"((u32)hi), "r"((u32)(hi >> 32)));
  return ((u64)r1 << 32) | r0;
}

static __constant__ u32 c_opaque_zero = 0;     // a zero ptxas cannot see through
static __constant__ u64 c_fast_post[22] = POSEIDON_FAST_POST;
static __constant__ u64 c_fast_what[22 * 11] = POSEIDON_FAST_WHAT;
static __constant__ u64 c_fast_vs[22 * 11] = POSEIDON_FAST_VS;
static __constant__ u64 c_fast_d3rot[12 * 12] = POSEIDON_FAST_D3ROT;
static __constant__ u64 c_fast_k3[12] = POSEIDON_FAST_K3;
static __constant__ u64 c_fast_u[22] = POSEIDON_FAST_U;
static __constant__ u64 c_fast_first[12] = POSEIDON_FAST_FIRST;

#ifdef SP_TRACE
__device__ long long* g_sp_trace;      // [64] clock64 stamps of block 0 / lane 0 of the word-0 warp, chunk 1 (lab only)
#define SP_STAMP(i) do { if (blockIdx.x == 0 && lane == 0 && m == 1) g_sp_trace[(i)] = clock64(); } while (0)
#else
#define SP_STAMP(i) do { } while (0)
#endif
// VAR (lab variants, tools/perf/poseidon_lab.cu; the product launches the one that measured fastest):
//   bit 0  the warp that owns word 0 sits alone in its SM sub-partition: 16 warps, warps 7, 11 and 15 (the ones that would
//          share scheduler 3 with warp 3) exit at once; its S-box chain then issues without competing for the scheduler or
//          the half-rate IMAD.WIDE pipe
//   bit 1  round 3's dense row (D3ROT, K3) lives in registers for the whole kernel instead of twelve indexed constant loads
//   bit 2  full-round MDS rows on dp2a (16-bit limbs, 24 IDP.2A per row) instead of 24 IMAD.WIDE
template <int VAR>
__global__ void __launch_bounds__((VAR & 1) ? 512 : 416) leaf_sponge_sp_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                             uint32_t n_leaves, unsigned log_block,
                                                             u64* __restrict__ digests,
                                                             const u64* __restrict__ state_in = nullptr,
                                                             u64* __restrict__ state_out = nullptr) {
  // state_in / state_out ([12][n_leaves], position order): the sponge fed in column order, one slab of columns per launch
  // (capi.cu: the leaf hashing of slab k runs while slab k+1 crosses PCIe); leaf_len % 8 == 0 on every launch but the last
  __shared__ __align__(16) u64 xch[2][24][32];
  __shared__ __align__(16) u64 ybuf[3][32];
  __shared__ __align__(16) u64 ebuf[3][12][32];
  __shared__ __align__(16) u64 Ebuf[3][64];   // per lane: (sum of low halves, sum of high halves)
  const unsigned lane = threadIdx.x & 31, warp_raw = threadIdx.x >> 5;
  unsigned warp = warp_raw;                   // logical warp 0..12
  if (VAR & 1) {
    if (warp_raw == 7 || warp_raw == 11 || warp_raw == 15) return;
    // physical 0..6 -> 0..6 (3 = word 0), 8..10 -> 7..9, 12..14 -> 10..12
    warp = warp_raw < 7 ? warp_raw : (warp_raw < 11 ? warp_raw - 1 : warp_raw - 2);
  }
  const bool reducer = warp == 12;
  const unsigned wid = reducer ? 0 : (warp + 9) % 12;      // the state word this warp owns
  const bool crit = !reducer && wid == 0;
  const uint32_t pos_raw = blockIdx.x * 32 + lane;
  const bool live = pos_raw < n_leaves;
  const uint32_t pos = live ? pos_raw : n_leaves - 1;
  const u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

  u64 s = 0, nx = 0;
  if (state_in && !reducer) s = state_in[(size_t)wid * n_leaves + pos];
  const uint32_t n_chunks = (leaf_len + 7) / 8;
  if (!reducer && wid < 8 && wid < leaf_len) nx = cols[(size_t)wid * n_leaves + pos];
  unsigned xb = 0;
  u64 d3row[12], k3v = 0;
  if (VAR & 2) {
#pragma unroll
    for (int i = 0; i < 12; i++) d3row[i] = c_fast_d3rot[12 * wid + i];
    k3v = c_fast_k3[wid];
  }

  for (uint32_t m = 0; m < n_chunks; m++) {
    if (!reducer) {
      const unsigned take = min(8u, leaf_len - 8 * m);
      if (wid < take) s = nx;
      if (m + 1 < n_chunks) {
        const uint32_t c = (m + 1) * 8 + wid;
        if (wid < 8 && c < leaf_len) nx = cols[(size_t)c * n_leaves + pos];
      }
      s = gl_add_lazy_canon(s, c_poseidon_rc[wid]);
      if (crit) SP_STAMP(0);

      // full round rd: publish x^7, read the twelve words rotated, small-coefficient MDS row + constant `next`
      auto full_round = [&](u64 next) {
        const u64 v = poseidon_sbox(s);
        xch[xb][wid][lane] = v; xch[xb][wid + 12][lane] = v;
        if (crit) SP_STAMP(50);
        named_bar_sync(1, 384);
        if (crit) SP_STAMP(51);
        u64 t[12];
#pragma unroll
        for (int i = 0; i < 12; i++) t[i] = xch[xb][wid + i][lane];
        if (VAR & 4) s = mds_row_dp2a(t, next, wid == 0);
        else {
          u32 al0, al1, ah0, ah1;
          mds_row<1>(t, [&](int i) { return CIRC[i]; }, next, 0, al0, al1, ah0, ah1);
          if (wid == 0) { mac32(al0, al1, (u32)t[0], 8u); mac32(ah0, ah1, (u32)(t[0] >> 32), 8u); }   // DIAG[0] = 8
          s = mds_recombine(al0, al1, ah0, ah1);
        }
        xb ^= 1;
        if (crit) SP_STAMP(52);
      };
#pragma unroll 1
      for (int rd = 0; rd < 3; rd++) { full_round(c_poseidon_rc[12 * (rd + 1) + wid]); if (crit) SP_STAMP(1 + rd); }
      if (crit) {
        full_round(c_fast_first[0]);        // word 0 passes the initial matrix unchanged
        SP_STAMP(4);
      } else {
        // round 3 with the dense initial matrix folded in: x_j = sum_i D3ROT[j][i] v_{(j+i) % 12} + K3[j]
        const u64 v = poseidon_sbox(s);
        xch[xb][wid][lane] = v; xch[xb][wid + 12][lane] = v;
        named_bar_sync(1, 384);
        // twelve unreduced 128-bit products into one 192-bit accumulator, one reduction (instead of twelve)
        Acc192 acc = {(VAR & 2) ? k3v : c_fast_k3[wid], 0, 0};
#pragma unroll
        for (int i = 0; i < 12; i++) acc192_mul(acc, (VAR & 2) ? d3row[i] : c_fast_d3rot[12 * wid + i], xch[xb][wid + i][lane]);
        s = gl_canon(acc192_reduce(acc));
        xb ^= 1;
      }
    }
    // ---- 22 sparse partial rounds (bodies take the slot q = r % 3 as a compile-time constant) ----
    if (crit) {
      u64 x0 = s;
      auto body = [&](int r, auto qc) {
        constexpr int q = decltype(qc)::value;
        const u64 post = c_fast_post[r];
        const u64 x2 = gl_mul_lazy(x0, x0);
        const u64 x4 = gl_mul_lazy(x2, x2), x3 = gl_mul_lazy(x2, x0);
        const u64 x7 = gl_mul_lazy(x3, x4);
        if (r < 8) SP_STAMP(8 + 4 * r);       // S-box issued
        // R needs y(r-1) for the look-ahead term and finishes Ebuf[q] ~200 cycles after it was published: waiting here,
        // behind the issue of the whole S-box, costs nothing unless R is late
        // the barrier id depends on x7 through a zero ptxas cannot fold: otherwise the S-box (pure register code) sinks
        // below the volatile bar.sync and the wait for R no longer overlaps it
        named_bar_sync(8 + q + ((u32)x7 & c_opaque_zero), 64);
        if (r < 8) SP_STAMP(9 + 4 * r);       // past the barrier
        const ulonglong2 E = *(const ulonglong2*)&Ebuf[q][2 * lane];
        const u64 y = gl_add_lazy_canon(x7, post);
        ybuf[q][lane] = y;
        named_bar_arrive(2 + q, 384);
        named_bar_arrive(11 + q, 64);
        u32 al0 = (u32)E.x, al1 = (u32)(E.x >> 32), ah0 = (u32)E.y, ah1 = (u32)(E.y >> 32);   // half sums, < 2^37
        mac32(al0, al1, (u32)y, 25u);
        mac32(ah0, ah1, (u32)(y >> 32), 25u);
        x0 = mds_recombine(al0, al1, ah0, ah1);
        if (r < 8) SP_STAMP(10 + 4 * r);      // x0 of the next round issued
      };
#pragma unroll 1
      for (int r = 0; r < 21; r += 3) {
        body(r, std::integral_constant<int, 0>()); body(r + 1, std::integral_constant<int, 1>()); body(r + 2, std::integral_constant<int, 2>());
      }
      body(21, std::integral_constant<int, 0>());
      s = x0;
      SP_STAMP(5);
    } else if (!reducer) {
      u64 x = s;                               // lazy
      const u64* what = c_fast_what + (wid - 1);
      const u64* vsp = c_fast_vs + (wid - 1);
      ebuf[0][wid][lane] = gl_mul_lazy(what[0], x);
      named_bar_arrive(5 + 0, 384);
      ebuf[1][wid][lane] = gl_mul_lazy(what[11], x);
      named_bar_arrive(5 + 1, 384);
      auto body = [&](int r, auto qc) {
        constexpr int q = decltype(qc)::value, q2 = (q + 2) % 3;
        const u64 vs = vsp[r * 11];
        const u64 wn = what[(r + 2 < 22 ? r + 2 : 0) * 11];
        named_bar_sync(2 + q, 384);
        const u64 y = ybuf[q][lane];
        x = gl_mad_lazy(vs, y, x);
        if (r + 2 < 22) {
          ebuf[q2][wid][lane] = gl_mul_lazy(wn, x);
          named_bar_arrive(5 + q2, 384);
        }
      };
#pragma unroll 1
      for (int r = 0; r < 21; r += 3) {
        body(r, std::integral_constant<int, 0>()); body(r + 1, std::integral_constant<int, 1>()); body(r + 2, std::integral_constant<int, 2>());
      }
      body(21, std::integral_constant<int, 0>());
      s = x;
    } else {
      // R: Ebuf[q] = (sum of low halves, sum of high halves) of the eleven products and of U[r] * y(r-1); the owner of
      // word 0 adds 25 * y on top and reduces once (every half sum stays below 2^37).
      auto body = [&](int r, auto qc) {
        constexpr int q = decltype(qc)::value, qp = (q + 2) % 3;
        const u64 u = c_fast_u[r];
        named_bar_sync(5 + q, 384);
        u64 e[12];
#pragma unroll
        for (int i = 1; i < 12; i++) e[i] = ebuf[q][i][lane];
        u64 al = 0, ah = 0;
#pragma unroll
        for (int i = 1; i < 12; i++) { al += (u32)e[i]; ah += e[i] >> 32; }
        if (r >= 1) {
          named_bar_sync(11 + qp, 64);
          const u64 uy = gl_mul_lazy(u, ybuf[qp][lane]);
          al += (u32)uy; ah += uy >> 32;
        }
        *(ulonglong2*)&Ebuf[q][2 * lane] = make_ulonglong2(al, ah);
        named_bar_arrive(8 + q, 64);
      };
#pragma unroll 1
      for (int r = 0; r < 21; r += 3) {
        body(r, std::integral_constant<int, 0>()); body(r + 1, std::integral_constant<int, 1>()); body(r + 2, std::integral_constant<int, 2>());
      }
      body(21, std::integral_constant<int, 0>());
      named_bar_sync(11 + 21 % 3, 64);          // consume the last Y arrival so that the ids start clean next time
    }
    // ---- last four full rounds ----
    if (!reducer) {
      s = gl_add_lazy_canon(s, c_poseidon_rc[12 * 26 + wid]);
#pragma unroll 1
      for (int rd = 26; rd < 30; rd++) {
        const u64 v = poseidon_sbox(s);
        xch[xb][wid][lane] = v; xch[xb][wid + 12][lane] = v;
        named_bar_sync(1, 384);
        u64 t[12];
#pragma unroll
        for (int i = 0; i < 12; i++) t[i] = xch[xb][wid + i][lane];
        if (VAR & 4) s = mds_row_dp2a(t, c_poseidon_rc[12 * (rd + 1) + wid], wid == 0);
        else {
          u32 al0, al1, ah0, ah1;
          mds_row<1>(t, [&](int i) { return CIRC[i]; }, c_poseidon_rc[12 * (rd + 1) + wid], 0, al0, al1, ah0, ah1);
          if (wid == 0) { mac32(al0, al1, (u32)t[0], 8u); mac32(ah0, ah1, (u32)(t[0] >> 32), 8u); }
          s = mds_recombine(al0, al1, ah0, ah1);
        }
        xb ^= 1;
      }
    }
  }
#ifdef SP_TRACE
  // (the chunk loop variable is out of scope here; the end-of-permutation stamp is taken inside the loop)
#endif
  if (state_out) {
    if (live && !reducer) state_out[(size_t)wid * n_leaves + pos] = s;
  } else if (live && !reducer && wid < 4) digests[4ull * leaf_index_of(pos, log_block) + wid] = gl_canon(s);
}

// ---------------------------------------------------------------------------------------------------------
// Throughput shapes (N = 32768 leaves: FinalExp, ECCAgg).  ncu on the dense-MDS kernels: sm__pipe_fmaheavy_cycles_active
// 81 % -- every IMAD-class instruction runs on the "heavy" half of the FMA pipe and IMAD.WIDE takes two passes, so the
// 288 IMAD.WIDE of a dense MDS layer are what bounds the machine, not issue slots (42 %) or the ALU pipe (26 %).
// In sparse form a partial round is 22 full multiplies (~265 heavy-pipe passes instead of ~600) with twenty-odd
// independent chains, so one thread can own a whole state: no shared memory, no barriers, lane = leaf.
//   x += FIRST, x[1..] = INIT x[1..]      folded into round 3's linear layer (D3, K3), sums of 128-bit products
//   per round: y = x0^7 + a_r;   x0 = 25 y + sum_i w^_r[i] x_i (one 192-bit accumulator, one reduction);   x_i += v_r[i] y
// ---------------------------------------------------------------------------------------------------------
static __constant__ u64 c_fast_d3[12 * 12] = POSEIDON_FAST_D3;

// full round on a register-resident state: s[i] holds the pre-S-box value (constant already added); `next` = the
// constants of the coming round (folded into the MDS accumulators)
__device__ __forceinline__ void st_full_round(u64 (&s)[12], const u64* __restrict__ next) {
  const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u32 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { const u64 v = poseidon_sbox(s[i]); lo[i] = (u32)v; hi[i] = (u32)(v >> 32); }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    const u64 c = next[r];
    u32 al0 = (u32)c, al1 = 0, ah0 = (u32)(c >> 32), ah1 = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { mac32(al0, al1, lo[(i + r) % 12], C[i]); mac32(ah0, ah1, hi[(i + r) % 12], C[i]); }
    if (r == 0) { mac32(al0, al1, lo[0], 8u); mac32(ah0, ah1, hi[0], 8u); }
    s[r] = mds_recombine(al0, al1, ah0, ah1);
  }
}

__global__ void __launch_bounds__(64) leaf_sponge_st_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                            uint32_t n_leaves, unsigned log_block,
                                                            u64* __restrict__ digests) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n_leaves) return;
  u64 s[12], nx[8];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  const uint32_t n_chunks = (leaf_len + 7) / 8;
  const u64* p = cols + pos;
#pragma unroll
  for (int i = 0; i < 8; i++) nx[i] = ((uint32_t)i < leaf_len) ? p[(size_t)i * n_leaves] : 0;
  for (uint32_t m = 0; m < n_chunks; m++) {
    const unsigned take = min(8u, leaf_len - 8 * m);
#pragma unroll
    for (int i = 0; i < 8; i++) if ((unsigned)i < take) s[i] = nx[i];
    if (m + 1 < n_chunks) {
      const u64* q = p + (size_t)(m + 1) * 8 * n_leaves;
#pragma unroll
      for (int i = 0; i < 8; i++) if ((m + 1) * 8 + i < leaf_len) nx[i] = q[(size_t)i * n_leaves];
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_lazy_canon(s[i], c_poseidon_rc[i]);
#pragma unroll 1
    for (int rd = 0; rd < 3; rd++) st_full_round(s, c_poseidon_rc + 12 * (rd + 1));
    {  // round 3 with FIRST and the dense initial matrix folded in: x0 = (M v)_0 + FIRST[0], x_j = sum_c D3[j][c] v_c + K3[j]
      u64 v[12];
#pragma unroll
      for (int i = 0; i < 12; i++) v[i] = poseidon_sbox(s[i]);
      const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
      const u64 c0 = c_fast_first[0];
      u32 al0 = (u32)c0, al1 = 0, ah0 = (u32)(c0 >> 32), ah1 = 0;
#pragma unroll
      for (int i = 0; i < 12; i++) { mac32(al0, al1, (u32)v[i], C[i]); mac32(ah0, ah1, (u32)(v[i] >> 32), C[i]); }
      mac32(al0, al1, (u32)v[0], 8u); mac32(ah0, ah1, (u32)(v[0] >> 32), 8u);
      s[0] = mds_recombine(al0, al1, ah0, ah1);
#pragma unroll 1
      for (int j = 1; j < 12; j++) {
        Acc192 a = {c_fast_k3[j], 0, 0};
#pragma unroll
        for (int c = 0; c < 12; c++) acc192_mul(a, c_fast_d3[12 * j + c], v[c]);
        const u64 r = acc192_reduce(a);
        // s[j] with a runtime j: keep the state in registers
#pragma unroll
        for (int k = 1; k < 12; k++) if (k == j) s[k] = r;
      }
    }
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
      const u64 y = gl_add_lazy_canon(poseidon_sbox(s[0]), c_fast_post[r]);
      const u64* wh = c_fast_what + 11 * r;
      const u64* vs = c_fast_vs + 11 * r;
      Acc192 a = {0, 0, 0};
#pragma unroll
      for (int i = 0; i < 11; i++) acc192_mul(a, wh[i], s[i + 1]);
      const u64 d = gl_canon(acc192_reduce(a));
      u32 al0 = (u32)d, al1 = 0, ah0 = (u32)(d >> 32), ah1 = 0;
      mac32(al0, al1, (u32)y, 25u);
      mac32(ah0, ah1, (u32)(y >> 32), 25u);
      s[0] = mds_recombine(al0, al1, ah0, ah1);
#pragma unroll
      for (int i = 0; i < 11; i++) s[i + 1] = gl_mad_lazy(vs[i], y, s[i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_lazy_canon(s[i], c_poseidon_rc[12 * 26 + i]);
#pragma unroll 1
    for (int rd = 26; rd < 30; rd++) st_full_round(s, c_poseidon_rc + 12 * (rd + 1));
  }
  u64* dg = digests + 4ull * leaf_index_of(pos, log_block);
#pragma unroll
  for (int i = 0; i < 4; i++) dg[i] = gl_canon(s[i]);
}

// ---------------------------------------------------------------------------------------------------------
// Dense MDS on the integer dot-product instruction.  Measured on B200 (tools/perf/pipe_lab.cu): IMAD.WIDE.U32 issues at
// 32 lanes/clk/SM -- half the rate of IMAD, IDP.2A and IDP.4A (64) -- and ncu shows the throughput-bound shapes at 81 %
// of sm__pipe_fmaheavy_cycles_active.  The MDS coefficients are <= 41, so a row is computed on 16-bit limbs with
//     dp2a.lo.u32.u32  acc, (limb_k of word j | limb_k of word j+1 << 16), (c_j | c_{j+1} << 8), acc
// : 7 word pairs x 4 limbs = 28 IDP per row (one pass each) instead of 24 IMAD.WIDE (two passes each), every limb sum
// < 2^27.  The pair packing (28 PRMT on the ALU pipe) is shared by the three rows a thread owns, which is why this
// variant exists for the three-words-per-thread layout only.
// ---------------------------------------------------------------------------------------------------------
// VAR != 0 (lab only, tools/perf/poseidon_lab dp): the warp that owns words 0..2 -- the only one with an S-box in the 22
// partial rounds -- differs from block to block (1: by block index, 2: by arrival order on the SM).  The idea was that
// warp k of every block sits on sub-partition k, so one scheduler would issue 22 x ~75 instructions per permutation
// more than the others.  Measured (profiles/r2_poseidon_lab_dp_roles.txt): both are 2-10 % SLOWER -- the hardware already
// staggers the warp slots of successive blocks over the sub-partitions, and a software rotation undoes part of that.
static __device__ unsigned dp_role_counter[256];
template <int VAR>
__global__ void __launch_bounds__(128) leaf_sponge_dp_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                             uint32_t n_leaves, unsigned log_block,
                                                             u64* __restrict__ digests,
                                                             const u64* __restrict__ state_in = nullptr,
                                                             u64* __restrict__ state_out = nullptr) {
  constexpr int W = 3, NW = W + 11, NP = NW / 2;
  __shared__ __align__(16) u64 xch[2][24][32];
  const unsigned lane = threadIdx.x & 31;
  unsigned rot = 0;
  if (VAR == 1) rot = blockIdx.x + blockIdx.x / 148;           // lab: assumes round-robin placement over 148 SMs
  if (VAR == 2) {                                               // arrival order on this SM: balanced whatever the block scheduler does
    __shared__ unsigned s_rot;
    if (threadIdx.x == 0) { unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); s_rot = atomicAdd(&dp_role_counter[sm & 255], 1u); }
    __syncthreads();
    rot = s_rot;
  }
  const unsigned wid = ((threadIdx.x >> 5) + rot) & 3;
  const unsigned w0 = wid * W;
  const uint32_t pos_raw = blockIdx.x * 32 + lane;
  const bool live = pos_raw < n_leaves;
  const uint32_t pos = live ? pos_raw : n_leaves - 1;
  constexpr u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

  u64 s[W], nx[W];
#pragma unroll
  for (int k = 0; k < W; k++) { s[k] = state_in ? state_in[(size_t)(w0 + k) * n_leaves + pos] : 0; nx[k] = 0; }
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
#pragma unroll
    for (int k = 0; k < W; k++) s[k] = gl_add_lazy_canon(s[k], c_poseidon_rc[w0 + k]);

    auto linear_layer = [&](const u64 (&v)[W], int rd) {
      u64* dst = &xch[xb][w0][lane];
#pragma unroll
      for (int k = 0; k < W; k++) { dst[32 * k] = v[k]; dst[32 * (k + 12)] = v[k]; }
      u64 rc[W];
#pragma unroll
      for (int k = 0; k < W; k++) rc[k] = c_poseidon_rc[12 * (rd + 1) + w0 + k];
      __syncthreads();
      const u64* src = &xch[xb][w0][lane];
      u64 t[NW];
#pragma unroll
      for (int j = 0; j < NW; j++) t[j] = src[32 * j];       // word (w0 + j) mod 12
      u32 acc[W][4];
#pragma unroll
      for (int k = 0; k < W; k++) {
        const u32 lo = (u32)rc[k], hi = (u32)(rc[k] >> 32);
        acc[k][0] = lo & 0xFFFFu; acc[k][1] = lo >> 16; acc[k][2] = hi & 0xFFFFu; acc[k][3] = hi >> 16;
      }
#pragma unroll
      for (int p = 0; p < NP; p++) {
        const u32 alo = (u32)t[2 * p], ahi = (u32)(t[2 * p] >> 32), blo = (u32)t[2 * p + 1], bhi = (u32)(t[2 * p + 1] >> 32);
        const u32 q0 = __byte_perm(alo, blo, 0x5410), q1 = __byte_perm(alo, blo, 0x7632);
        const u32 q2 = __byte_perm(ahi, bhi, 0x5410), q3 = __byte_perm(ahi, bhi, 0x7632);
#pragma unroll
        for (int k = 0; k < W; k++) {
          const int i0 = 2 * p - k, i1 = 2 * p + 1 - k;            // circulant index of the two words in row k
          const u32 c0 = (i0 >= 0 && i0 < 12) ? CIRC[i0] : 0u, c1 = (i1 >= 0 && i1 < 12) ? CIRC[i1] : 0u;
          const u32 b = c0 | (c1 << 8);
          if (b) {
            acc[k][0] = dp2a_lo(q0, b, acc[k][0]); acc[k][1] = dp2a_lo(q1, b, acc[k][1]);
            acc[k][2] = dp2a_lo(q2, b, acc[k][2]); acc[k][3] = dp2a_lo(q3, b, acc[k][3]);
          }
        }
        if (p == 0 && wid == 0) {                                   // DIAG[0] = 8 on row 0 / word 0
          acc[0][0] = dp2a_lo(q0, 8u, acc[0][0]); acc[0][1] = dp2a_lo(q1, 8u, acc[0][1]);
          acc[0][2] = dp2a_lo(q2, 8u, acc[0][2]); acc[0][3] = dp2a_lo(q3, 8u, acc[0][3]);
        }
      }
#pragma unroll
      for (int k = 0; k < W; k++) s[k] = limbs_recombine(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
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
  if (state_out) {
    if (live) {
#pragma unroll
      for (int k = 0; k < W; k++) state_out[(size_t)(w0 + k) * n_leaves + pos] = s[k];
    }
  } else if (live) {
    u64* d = digests + 4ull * leaf_index_of(pos, log_block);
#pragma unroll
    for (int k = 0; k < W; k++)
      if (w0 + k < 4) d[w0 + k] = gl_canon(s[k]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Throughput shapes, second form: dp2a full rounds + SPARSE partial rounds in the same three-words-per-thread layout.
// The dp kernel spends 22 of its 30 rounds on a dense MDS of which only word 0 went through the S-box: per leaf and
// round 336 IDP.2A + 112 PRMT + 12 recombinations, with three of the four warps waiting at the barrier for warp 0's
// S-box (ncu: 26 % of the warp samples).  In sparse form (tables as in the sp kernel) a partial round is
//     helpers:  x_i += v_{r-1}[i] y(r-1)            (the update of the PREVIOUS round, 3 multiply-adds per thread)
//               d_w  = sum_{i in mine} w^_r[i] x_i    (3 multiplies into one 192-bit accumulator, one reduction)
//     warp 0 :  the same for words 1, 2 and  y(r) = x0^7 + a_r          -- its S-box overlaps everybody's linear part
//     one barrier, then  x0 = 25 y(r) + d_0 + d_1 + d_2 + d_3  (warp 0)  and everybody picks up y(r)
// : ~635 instructions per leaf and round instead of ~940, one barrier per round as before.  Round 3's linear layer is
// the dense D3 / K3 form (FIRST and the 11x11 INIT matrix folded in), 12 full multiplies per row.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void acc192_add(Acc192& a, u64 x) {
  asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, 0;\n\taddc.u32 %2, %2, 0;" : "+l"(a.lo), "+l"(a.hi), "+r"(a.top) : "l"(x));
}

__global__ void __launch_bounds__(128) leaf_sponge_ds_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                             uint32_t n_leaves, unsigned log_block,
                                                             u64* __restrict__ digests) {
  constexpr int W = 3, NW = W + 11, NP = NW / 2;
  __shared__ __align__(16) u64 xch[2][24][32];
  __shared__ __align__(16) u64 dsl[2][5][32];
  // the sparse tables in shared memory: indexed per warp and per round, 8 KB in all -- more than the constant cache
  // holds, and an indexed LDC that misses costs hundreds of cycles on the round's critical path
  __shared__ u64 tb_what[22 * 11], tb_vs[22 * 11], tb_post[22], tb_d3[144], tb_k3[12];
  for (unsigned i = threadIdx.x; i < 22 * 11; i += blockDim.x) { tb_what[i] = c_fast_what[i]; tb_vs[i] = c_fast_vs[i]; }
  for (unsigned i = threadIdx.x; i < 144; i += blockDim.x) tb_d3[i] = c_fast_d3[i];
  if (threadIdx.x < 22) tb_post[threadIdx.x] = c_fast_post[threadIdx.x];
  if (threadIdx.x < 12) tb_k3[threadIdx.x] = c_fast_k3[threadIdx.x];
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned w0 = wid * W;
  const uint32_t pos_raw = blockIdx.x * 32 + lane;
  const bool live = pos_raw < n_leaves;
  const uint32_t pos = live ? pos_raw : n_leaves - 1;
  constexpr u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

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
  unsigned xb = 0, db = 0;
  for (uint32_t m = 0; m < n_chunks; m++) {
    const unsigned take = min(8u, leaf_len - 8 * m);
#pragma unroll
    for (int k = 0; k < W; k++)
      if (w0 + k < take) s[k] = nx[k];
    if (m + 1 < n_chunks) fetch(m + 1);
#pragma unroll
    for (int k = 0; k < W; k++) s[k] = gl_add_lazy_canon(s[k], c_poseidon_rc[w0 + k]);

    // dense MDS on dp2a (identical to leaf_sponge_dp_kernel): S-box the three words, exchange, three rows, next constants
    auto full_round = [&](int rd) {
      u64 v[W];
#pragma unroll
      for (int k = 0; k < W; k++) v[k] = poseidon_sbox(s[k]);
      u64* dst = &xch[xb][w0][lane];
#pragma unroll
      for (int k = 0; k < W; k++) { dst[32 * k] = v[k]; dst[32 * (k + 12)] = v[k]; }
      u64 rc[W];
#pragma unroll
      for (int k = 0; k < W; k++) rc[k] = c_poseidon_rc[12 * (rd + 1) + w0 + k];
      __syncthreads();
      const u64* src = &xch[xb][w0][lane];
      u64 t[NW];
#pragma unroll
      for (int j = 0; j < NW; j++) t[j] = src[32 * j];       // word (w0 + j) mod 12
      u32 acc[W][4];
#pragma unroll
      for (int k = 0; k < W; k++) {
        const u32 lo = (u32)rc[k], hi = (u32)(rc[k] >> 32);
        acc[k][0] = lo & 0xFFFFu; acc[k][1] = lo >> 16; acc[k][2] = hi & 0xFFFFu; acc[k][3] = hi >> 16;
      }
#pragma unroll
      for (int p = 0; p < NP; p++) {
        const u32 alo = (u32)t[2 * p], ahi = (u32)(t[2 * p] >> 32), blo = (u32)t[2 * p + 1], bhi = (u32)(t[2 * p + 1] >> 32);
        const u32 q0 = __byte_perm(alo, blo, 0x5410), q1 = __byte_perm(alo, blo, 0x7632);
        const u32 q2 = __byte_perm(ahi, bhi, 0x5410), q3 = __byte_perm(ahi, bhi, 0x7632);
#pragma unroll
        for (int k = 0; k < W; k++) {
          const int i0 = 2 * p - k, i1 = 2 * p + 1 - k;
          const u32 c0 = (i0 >= 0 && i0 < 12) ? CIRC[i0] : 0u, c1 = (i1 >= 0 && i1 < 12) ? CIRC[i1] : 0u;
          const u32 b = c0 | (c1 << 8);
          if (b) {
            acc[k][0] = dp2a_lo(q0, b, acc[k][0]); acc[k][1] = dp2a_lo(q1, b, acc[k][1]);
            acc[k][2] = dp2a_lo(q2, b, acc[k][2]); acc[k][3] = dp2a_lo(q3, b, acc[k][3]);
          }
        }
        if (p == 0 && wid == 0) {
          acc[0][0] = dp2a_lo(q0, 8u, acc[0][0]); acc[0][1] = dp2a_lo(q1, 8u, acc[0][1]);
          acc[0][2] = dp2a_lo(q2, 8u, acc[0][2]); acc[0][3] = dp2a_lo(q3, 8u, acc[0][3]);
        }
      }
#pragma unroll
      for (int k = 0; k < W; k++) s[k] = limbs_recombine(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
      xb ^= 1;
    };
#pragma unroll 1
    for (int rd = 0; rd < 3; rd++) full_round(rd);
    // ---- partial rounds in their own layout: warp 0 owns x0 and nothing else (its S-box + combine chain IS the critical
    // path of a round), warps 1..3 own words 1-4, 5-8, 9-11 ----
    const unsigned p0 = wid == 0 ? 0u : 1u + 4u * (wid - 1u), np = wid == 0 ? 1u : (wid == 3 ? 3u : 4u);
    u64 x[4] = {0, 0, 0, 0};
    {  // round 3: x = D3 v + K3  (row 0 of D3 is the MDS row, K3[0] = FIRST[0]); every thread computes the rows it owns
      u64* dst = &xch[xb][w0][lane];
#pragma unroll
      for (int k = 0; k < W; k++) dst[32 * k] = poseidon_sbox(s[k]);
      __syncthreads();
      u64 t[12];
#pragma unroll
      for (int c = 0; c < 12; c++) t[c] = xch[xb][c][lane];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if ((unsigned)q < np) {
          const u64* row = tb_d3 + 12 * (p0 + q);
          Acc192 a = {tb_k3[p0 + q], 0, 0};
#pragma unroll
          for (int c = 0; c < 12; c++) acc192_mul(a, row[c], t[c]);
          x[q] = acc192_reduce(a);
        }
      }
      xb ^= 1;
    }
    u64 y = 0;
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
      Acc192 b = {0, 0, 0};
      if (wid == 0) {
        y = gl_add_lazy_canon(poseidon_sbox(x[0]), tb_post[r]);
        dsl[db][3][lane] = y;
        acc192_mul(b, 25ull, y);
      } else {
        const u64* wh = tb_what + 11 * r + (p0 - 1);
        const u64* vp = tb_vs + 11 * (r - 1) + (p0 - 1);
        Acc192 a = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if ((unsigned)q < np) {
            if (r > 0) x[q] = gl_mad_lazy(vp[q], y, x[q]);
            acc192_mul(a, wh[q], x[q]);
          }
        }
        dsl[db][wid - 1][lane] = acc192_reduce(a);
      }
      __syncthreads();
      if (wid == 0) {
#pragma unroll
        for (int q = 0; q < 3; q++) acc192_add(b, dsl[db][q][lane]);
        x[0] = acc192_reduce(b);
      } else {
        y = dsl[db][3][lane];
      }
      db ^= 1;
    }
    {  // the update of the last partial round, the constants of round 26, and back to the three-words-per-thread layout
      const u64* vp = tb_vs + 11 * 21 + (p0 - 1);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if ((unsigned)q < np) {
          if (wid > 0) x[q] = gl_mad_lazy(vp[q], y, x[q]);
          xch[xb][p0 + q][lane] = gl_add_lazy_canon(x[q], c_poseidon_rc[12 * 26 + p0 + q]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < W; k++) s[k] = xch[xb][w0 + k][lane];
      xb ^= 1;
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
