// K2, third form: the dense MDS layer of the Poseidon leaf sponge on the integer tensor-core path (IMMA).
// Replaces PoseidonHash::hash_or_noop of every leaf in plonky2's MerkleTree::new (reached from
// starky::prover::prove, reference call sites /root/reference/src/aggregate_proof.rs:59,105,138,169,212).
//
// The dp kernel (leafhash.cuh) spends two thirds of its instructions moving the state between the four threads of a
// leaf: shared-memory exchange, one barrier per round, 16-bit pair packing, 340 IDP.2A per leaf and round.  Here the
// exchange IS the matrix instruction.  mma.sync.m16n8k32 (u8 x u8 -> s32) computes D[16 x 8] = A[16 x 32] B[32 x 8] + C
// with the operands spread over the lanes of a warp in a fixed pattern (lane = 4 g + t):
//     A: lane (g, t) supplies k = 4t..4t+3 and 16+4t..16+4t+3 of rows g and g+8
//     D: lane (g, t) receives columns 2t, 2t+1 of rows g and g+8
// Rows are leaves.  Lane (g, t) owns words 3t..3t+2 of the two leaves of rows g, g+8.  For a pair of byte limbs
// (2q, 2q+1) the K index is (word, limb class): k = 4t + i is limb 2q of word 3t+i (i = 3: padding, its B rows are
// zero), k = 16 + 4t + i is limb 2q+1.  Three B tiles j = 0..2 map it to N = (output row 3t'+j, limb class e) at column
// 2t'+e, so lane (g, t') receives the limb sums of ITS OWN rows 3t'..3t'+2: after the instruction every lane holds the
// next state of the words it owned before.  No shared-memory exchange, no barrier, one warp per 16 (or 32) leaves.
//     per 16 leaves and round: 24 PRMT (byte transposes of 3 words x 8 limbs), 12 IMMA, 6 limb recombinations
// Limb sums are < 264 * 255 < 2^17, exact in s32; the round constants of the next round join the two half sums of
// the recombination (as the C operand they cost four register moves per instruction).  The S-box of a partial round touches word 0 only, which lives on the t = 0
// lanes: the leaves of a quad are handed out to its lanes by shuffle so that the x^7 runs once per warp and round.
#pragma once
#include "leafhash.cuh"

template <int DBG = 0>
__device__ __forceinline__ void imma_16832_u8(u32 (&d)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1) {
  if (DBG & 2) { d[0] = a0 ^ b0; d[1] = a1 ^ b1; d[2] = a2 ^ b0; d[3] = a3 ^ b1; return; }
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}
// the same with the accumulator as input: d = a b + d
template <int DBG = 0>
__device__ __forceinline__ void imma_16832_u8_acc(u32 (&d)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1) {
  if (DBG & 2) { d[0] += a0 ^ b0; d[1] += a1 ^ b1; d[2] += a2 ^ b0; d[3] += a3 ^ b1; return; }
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
static __constant__ u64 c_poseidon_rc_eq[31 * 12] = POSEIDON_RC_EQ;
// the Goldilocks fold of gl_mul_lazy with x2 * eps + (x1:x0) as one multiply-add (two instructions move from the ALU to the FMA pipe)
__device__ __forceinline__ u64 gl_mul_lazy_fma(u64 a, u64 b) {
  const u64 lo = a * b, hi = __umul64hi(a, b);
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 c1, b1;\n\t"
      "mad.lo.cc.u32 %0, %4, 0xFFFFFFFF, %2;\n\t"
      "madc.hi.cc.u32 %1, %4, 0xFFFFFFFF, %3;\n\t"
      "addc.u32 c1, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, %5;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 b1, 0, 0;\n\t"
      "neg.s32 c1, c1;\n\t"
      "add.cc.u32 %0, %0, c1;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "sub.cc.u32 %0, %0, b1;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"((u32)hi), "r"((u32)(hi >> 32)));
  return ((u64)r1 << 32) | r0;
}
// x^7.  One warp per scheduler (NL == 1): the multiply-add fold is 3 % shorter in latency; with more warps the two
// forms measure the same (profiles/r2_poseidon_lab_mm.txt)
template <int NL, int DBG>
__device__ __forceinline__ u64 mm_sbox(u64 x) {
  if ((NL == 1) != ((DBG & 16) != 0)) {
    const u64 x2 = gl_mul_lazy_fma(x, x), x4 = gl_mul_lazy_fma(x2, x2), x3 = gl_mul_lazy_fma(x2, x);
    return gl_mul_lazy_fma(x3, x4);
  }
  return poseidon_sbox(x);
}
__device__ __forceinline__ u64 shfl64(u64 v, unsigned src) {
  const u32 lo = __shfl_sync(0xFFFFFFFFu, (u32)v, src), hi = __shfl_sync(0xFFFFFFFFu, (u32)(v >> 32), src);
  return ((u64)hi << 32) | lo;
}

// NL = leaves per lane.  4: two 16-row tiles, 32 leaves per warp; 2: one tile, 16 leaves per warp; 1: half a tile (rows
// 8..15 unused), 8 leaves per warp -- the latency-bound shapes (MillerLoop 2048, PairingPrecomp 4096 leaves, a FinalExp
// shard on 8 GPUs), where wall time = chain length x time of one permutation and the fewest instructions per warp win.
// DBG (lab only; tools/perf/poseidon_lab mmx).  Timing by elimination, wrong digests: 1 = no S-box in partial rounds,
// 2 = no matrix instruction, 4 = no recombination, 8 = no S-box in full rounds.  Correct variants: 16 = the other fold in
// the S-box, 32 = limb pairs as PRMT + IADD3 on the ALU pipe instead of IMAD.  256 (NL == 1, product): helper lanes.
// Round constants in the equivalent form with one non-zero constant per partial round (poseidon_fast.h, RC_EQ):
//   rcs: every round as (low half, high half), each a u64, for the layers that feed a full round; row 30 = zeros
//   rcw: the word-0 constants of rounds 5..25 as accumulator images (16-bit chunks on the even limbs), [1] for the lanes
//        that own word 0 and zeros [0] for the others; the image depends on the tile layout (NL == 1 or not)
struct __align__(16) MmTables {
  u64 rcs[31][12][2];
  u32 rcw[21][2][4][4];   // [round][owns word 0][instruction][accumulator register]
};
template <int NL>
__device__ __forceinline__ void mm_fill_tables(MmTables& T) {      // by the whole block; the caller synchronises
  for (unsigned i = threadIdx.x; i < 31 * 12; i += blockDim.x) {
    const u64 c = c_poseidon_rc_eq[i];
    T.rcs[i / 12][i % 12][0] = c & 0xFFFFFFFFull;
    T.rcs[i / 12][i % 12][1] = c >> 32;
  }
  for (unsigned i = threadIdx.x; i < 21 * 4; i += blockDim.x) {
    const u64 c = c_poseidon_rc_eq[12 * (5 + i / 4)];
    const unsigned q = i % 4;
    u32* o = T.rcw[i / 4][1][q];
    u32* z = T.rcw[i / 4][0][q];
    if (NL == 1) {      // instruction q < 2 holds limbs 4q..4q+3: chunks 2q, 2q+1 on its registers 0 and 2
      o[0] = q < 2 ? (u32)(c >> (32 * q)) & 0xFFFFu : 0u; o[1] = 0; o[2] = q < 2 ? (u32)(c >> (32 * q + 16)) & 0xFFFFu : 0u; o[3] = 0;
    } else {            // instruction q holds limbs 2q, 2q+1 of rows g and g+8: chunk q on registers 0 and 2
      const u32 ch = (u32)(c >> (16 * q)) & 0xFFFFu;
      o[0] = ch; o[1] = 0; o[2] = ch; o[3] = 0;
    }
    z[0] = 0; z[1] = 0; z[2] = 0; z[3] = 0;
  }
}

// the sponge of the 8 NL leaves base .. base + 8 NL - 1, by one warp
template <int NL, int DBG>
__device__ __forceinline__ void mm_sponge_warp(const MmTables& T, uint32_t base, const u64* __restrict__ cols, uint32_t leaf_len,
                                               uint32_t n_leaves, unsigned log_block, u64* __restrict__ digests,
                                               const u64* __restrict__ state_in, u64* __restrict__ state_out) {
  constexpr int MT = NL > 2 ? NL / 2 : 1, H = NL > 1 ? 2 : 1;       // row tiles, rows per lane and tile
  const auto& rcs = T.rcs;
  const auto& rcw = T.rcw;
  const unsigned lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (base >= n_leaves) return;

  uint32_t pos[NL];
  bool live[NL];
#pragma unroll
  for (int L = 0; L < NL; L++) {
    // (helper-lane form, DBG & 256: four leaves per warp on quads 0..3; the lanes of quads 4..7 own nothing)
    const uint32_t raw = NL == 1 ? (((DBG & 256) && g >= 4) ? n_leaves : base + g) : base + 16 * (L >> 1) + 2 * g + (L & 1);
    live[L] = raw < n_leaves;
    pos[L] = live[L] ? raw : n_leaves - 1;
  }
  // B tiles: B_j[k = (word 3t+i, class)][n = g -> (row 3 (g >> 1) + j, class g & 1)]
  u32 bf[3][2];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    u32 v = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const unsigned r = 3 * (g >> 1) + j, c = 3 * t + i;
      const u32 coef = c_poseidon_circ[(c + 12 - r) % 12] + ((r == 0 && c == 0) ? 8u : 0u);
      v |= coef << (8 * i);
    }
    bf[j][0] = (g & 1) ? 0u : v;
    bf[j][1] = (g & 1) ? v : 0u;
  }

  u64 s[NL][3], nx[NL][3];
#pragma unroll
  for (int L = 0; L < NL; L++)
#pragma unroll
    for (int k = 0; k < 3; k++) { s[L][k] = state_in ? state_in[(size_t)(3 * t + k) * n_leaves + pos[L]] : 0; nx[L][k] = 0; }
  const uint32_t n_chunks = (leaf_len + 7) / 8;
  auto fetch = [&](uint32_t chunk) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const uint32_t w = 3 * t + k, c = chunk * 8 + w;
      if (w < 8 && c < leaf_len) {
        const u64* p = cols + (size_t)c * n_leaves;
#pragma unroll
        for (int L = 0; L < NL; L++) nx[L][k] = p[pos[L]];
      }
    }
  };
  fetch(0);

  // A operands of row tile m: Q[i] = the four registers of instruction i.
  //   NL >= 2: i = limb pair q (4 instructions): (limb 2q of row g, of row g+8, limb 2q+1 of row g, of row g+8)
  //   NL == 1: the rows g+8 carry two more limbs of the SAME leaf, i = limb quad (2 instructions):
  //            (limb 4i of row g, limb 4i+2 as row g+8, limb 4i+1 of row g, limb 4i+3 as row g+8)  ->  D = limbs 4i..4i+3
  constexpr int NI = NL == 1 ? 2 : 4;
  auto pack = [&](int m, u32 (&Q)[NI][4]) {
#pragma unroll
    for (int h = 0; h < H; h++) {
      const u64* w = s[2 * m + h];
#pragma unroll
      for (int half = 0; half < 2; half++) {
        // word 0 (the only one behind the S-box of a partial round) enters last: one PRMT between x^7 and the instruction
        const u32 x0 = (u32)(w[0] >> (32 * half)), x1 = (u32)(w[1] >> (32 * half)), x2 = (u32)(w[2] >> (32 * half));
        const u32 u0 = __byte_perm(x1, x2, 0x5140), u1 = __byte_perm(x1, x2, 0x7362);
        const u32 l0 = __byte_perm(x0, u0, 0x5540), l1 = __byte_perm(x0, u0, 0x7761);
        const u32 l2 = __byte_perm(x0, u1, 0x5542), l3 = __byte_perm(x0, u1, 0x7763);
        if constexpr (NL == 1) { Q[half][0] = l0; Q[half][2] = l1; Q[half][1] = l2; Q[half][3] = l3; }
        else { Q[2 * half][h] = l0; Q[2 * half][2 + h] = l1; Q[2 * half + 1][h] = l2; Q[2 * half + 1][2 + h] = l3; }
      }
    }
  };
  // limb sums (j, i) -> the state word of row h:  16-bit pairs, then the two half sums (+ the constant), then mod p
  auto recombine = [&](const u32 (&D)[NI][4], int h, u64 cl, u64 ch) -> u64 {
    u32 p[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const u32 de = NL == 1 ? D[q >> 1][2 * (q & 1)] : D[q][2 * h], dod = NL == 1 ? D[q >> 1][2 * (q & 1) + 1] : D[q][2 * h + 1];
      p[q] = (DBG & 32) ? de + __byte_perm(dod, 0, 0x2104) : de + (dod << 8);
    }
    if (DBG & 4) return ((u64)(p[0] ^ p[2]) << 32 | (p[1] ^ p[3])) + cl + ch;
    const u64 al = (u64)p[1] * 65536ull + cl + p[0], ah = (u64)p[3] * 65536ull + ch + p[2];     // both < 2^41
    return mds_recombine((u32)al, (u32)(al >> 32), (u32)ah, (u32)(ah >> 32));
  };
  // linear layer that feeds a full round (or ends the permutation): every word gets its constant, in the recombination
  auto linear_layer = [&](int rd) {
    const ulonglong2* rcp = reinterpret_cast<const ulonglong2*>(&rcs[rd + 1][3 * t][0]);
    ulonglong2 rc[3];
#pragma unroll
    for (int j = 0; j < 3; j++) rc[j] = rcp[j];
#pragma unroll
    for (int m = 0; m < MT; m++) {
      u32 Q[NI][4];
      pack(m, Q);
      u32 D[3][NI][4];
#pragma unroll
      for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < NI; i++) imma_16832_u8<DBG>(D[j][i], Q[i][0], Q[i][1], Q[i][2], Q[i][3], bf[j][0], bf[j][1]);
#pragma unroll
      for (int h = 0; h < H; h++)
#pragma unroll
        for (int j = 0; j < 3; j++) s[2 * m + h][j] = recombine(D[j], h, rc[j].x, rc[j].y);
    }
  };
  // linear layer that feeds a partial round: only word 0 has a constant; it enters as the accumulator of row 0's
  // instructions, which are issued first -- the next S-box waits for them alone
  const unsigned own0 = t == 0 ? 1u : 0u;
  auto issue_p = [&](int m, int rd, u32 (&D)[3][NI][4]) {
    const uint4* cw = reinterpret_cast<const uint4*>(&rcw[rd - 4][own0][0][0]);
    u32 Q[NI][4];
    pack(m, Q);
#pragma unroll
    for (int i = 0; i < NI; i++) {
      const uint4 c = cw[i];
      D[0][i][0] = c.x; D[0][i][1] = c.y; D[0][i][2] = c.z; D[0][i][3] = c.w;
      imma_16832_u8_acc<DBG>(D[0][i], Q[i][0], Q[i][1], Q[i][2], Q[i][3], bf[0][0], bf[0][1]);
    }
#pragma unroll
    for (int j = 1; j < 3; j++)
#pragma unroll
      for (int i = 0; i < NI; i++) imma_16832_u8<DBG>(D[j][i], Q[i][0], Q[i][1], Q[i][2], Q[i][3], bf[j][0], bf[j][1]);
  };
  auto linear_layer_p = [&](int rd) {
#pragma unroll
    for (int m = 0; m < MT; m++) {
      u32 D[3][NI][4];
      issue_p(m, rd, D);
#pragma unroll
      for (int h = 0; h < H; h++)
#pragma unroll
        for (int j = 0; j < 3; j++) s[2 * m + h][j] = recombine(D[j], h, 0, 0);
    }
  };
  // the 3 NL S-boxes of a full round.  DBG 64 / 128 / 192 (lab): groups of 3 / 6 / 12 values advance in lock step in
  // source order (x^2 of every value, then x^4 and x^3, then x^7) instead of one x^7 after the other
  auto sbox_all = [&]() {
    constexpr int G = (DBG & 192) == 64 ? 3 : (DBG & 192) == 128 ? 6 : (DBG & 192) == 192 ? 12 : 1;
    if constexpr (NL == 1 && (DBG & 256) != 0) {
      // Helper lanes (the shapes with so few leaves that half of every warp can stay empty: MillerLoop 2048, FP12Mul 32):
      // a lone warp issues one instruction every ~2 cycles, so a full round costs its instruction count -- lane + 16
      // takes the third x^7 of lane's leaf and the warp executes two S-boxes instead of three.
      const bool helper = lane >= 16;
      const u64 w2 = shfl64(s[0][2], lane & 15);
      u64 a = helper ? w2 : s[0][0], b = s[0][1];
      a = mm_sbox<NL, DBG>(a); b = mm_sbox<NL, DBG>(b);
      const u64 back = shfl64(a, lane | 16);
      if (!helper) { s[0][0] = a; s[0][1] = b; s[0][2] = back; }
    } else if constexpr (G == 1 || (3 * NL) % G != 0) {
#pragma unroll
      for (int L = 0; L < NL; L++)
#pragma unroll
        for (int k = 0; k < 3; k++) s[L][k] = (DBG & 8) ? s[L][k] + 3 : mm_sbox<NL, DBG>(s[L][k]);
    } else {
#pragma unroll
      for (int g0 = 0; g0 < 3 * NL; g0 += G) {
        u64 x2[G], x3[G], x4[G];
#pragma unroll
        for (int i = 0; i < G; i++) { const u64 x = s[(g0 + i) / 3][(g0 + i) % 3]; x2[i] = gl_mul_lazy(x, x); }
#pragma unroll
        for (int i = 0; i < G; i++) { const u64 x = s[(g0 + i) / 3][(g0 + i) % 3]; x4[i] = gl_mul_lazy(x2[i], x2[i]); x3[i] = gl_mul_lazy(x2[i], x); }
#pragma unroll
        for (int i = 0; i < G; i++) s[(g0 + i) / 3][(g0 + i) % 3] = gl_mul_lazy(x3[i], x4[i]);
      }
    }
  };
  // word 0 of the quad's NL leaves sits on lane t = 0: lane L of the quad takes leaf L (lanes >= NL redo leaf 0)
  auto sbox_word0 = [&]() {
    const unsigned q0 = lane & ~3u;
    u64 v = s[0][0];
#pragma unroll
    for (int L = 1; L < NL; L++) { const u64 x = shfl64(s[L][0], q0); if (t == (unsigned)L) v = x; }
    const u64 y = (DBG & 1) ? v + 3 : mm_sbox<NL, DBG>(v);
#pragma unroll
    for (int L = 1; L < NL; L++) { const u64 x = shfl64(y, q0 | L); if (t == 0) s[L][0] = x; }
    if (t == 0) s[0][0] = y;
  };

  for (uint32_t m = 0; m < n_chunks; m++) {
    const unsigned take = min(8u, leaf_len - 8 * m);
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (3 * t + k < take) {
#pragma unroll
        for (int L = 0; L < NL; L++) s[L][k] = nx[L][k];
      }
    if (m + 1 < n_chunks) fetch(m + 1);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const u64 c = rcs[0][3 * t + k][0] | (rcs[0][3 * t + k][1] << 32);
#pragma unroll
      for (int L = 0; L < NL; L++) s[L][k] = gl_add_lazy_canon(s[L][k], c);
    }
#pragma unroll 1
    for (int rd = 0; rd < 4; rd++) { sbox_all(); linear_layer(rd); }
#pragma unroll 1
    for (int rd = 4; rd < 25; rd++) { sbox_word0(); linear_layer_p(rd); }
    sbox_word0();
    linear_layer(25);
#pragma unroll 1
    for (int rd = 26; rd < 30; rd++) { sbox_all(); linear_layer(rd); }
  }
#pragma unroll
  for (int L = 0; L < NL; L++) {
    if (!live[L]) continue;
    if (state_out) {
#pragma unroll
      for (int k = 0; k < 3; k++) state_out[(size_t)(3 * t + k) * n_leaves + pos[L]] = s[L][k];
    } else {
      u64* d = digests + 4ull * leaf_index_of(pos[L], log_block);
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (3 * t + k < 4) d[3 * t + k] = gl_canon(s[L][k]);
    }
  }
}

template <int NL, int DBG = 0>
__global__ void __launch_bounds__(NL == 1 ? 32 : 64) leaf_sponge_mm_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                            uint32_t n_leaves, unsigned log_block,
                                                            u64* __restrict__ digests,
                                                            const u64* __restrict__ state_in = nullptr,
                                                            u64* __restrict__ state_out = nullptr) {
  __shared__ MmTables T;
  mm_fill_tables<NL>(T);
  __syncthreads();
  constexpr uint32_t LPW = (NL == 1 && (DBG & 256)) ? 4 : 8 * NL;          // leaves per warp
  const uint32_t base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * LPW;
  mm_sponge_warp<NL, DBG>(T, base, cols, leaf_len, n_leaves, log_block, digests, state_in, state_out);
}

// Throughput-bound shapes (32 768 leaves on 148 SMs = 55.35 leaves per scheduler): with one kind of warp the schedulers end
// up with 48 or 64 leaves and the kernel lasts as long as the fuller ones (13.5 % of the machine idle).  Here one block per
// SM gives EVERY scheduler (warp index mod 4) the same mix: N4 32-leaf warps, N2 16-leaf warps and N1 8-leaf warps -- the
// product uses (0, 3, 1) = 56 leaves per scheduler.  Launch with 128 (N4 + N2 + N1) threads and enough dynamic shared
// memory that two blocks do not share an SM (merkle.cu).
template <int N4, int N2, int N1>
struct MmHet {
  static constexpr uint32_t PER_SCHED = 32 * N4 + 16 * N2 + 8 * N1, LEAVES = 4 * PER_SCHED, THREADS = 128 * (N4 + N2 + N1);
};
template <int N4, int N2, int N1, int DBG = 0>
__global__ void __launch_bounds__(128 * (N4 + N2 + N1)) leaf_sponge_mm_het_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                                  uint32_t n_leaves, unsigned log_block, u64* __restrict__ digests,
                                                                  const u64* __restrict__ state_in = nullptr,
                                                                  u64* __restrict__ state_out = nullptr) {
  using H = MmHet<N4, N2, N1>;
  __shared__ MmTables T2, T1;
  mm_fill_tables<2>(T2);                                       // (the 32- and 16-leaf warps share the accumulator images)
  mm_fill_tables<1>(T1);
  __syncthreads();
  const unsigned w = threadIdx.x >> 5, sched = w & 3;
  const int slot = (int)(w >> 2);
  // leaves of the block: first the 32-leaf warps' (slot-major, scheduler-minor), then the 16-leaf warps', then the 8-leaf
  // warps' -- every warp reads whole 128-byte lines of a column (64 bytes for the 8-leaf warps)
  const uint32_t b0 = blockIdx.x * H::LEAVES;
  if (slot < N4) mm_sponge_warp<4, DBG>(T2, b0 + 32 * (4 * slot + sched), cols, leaf_len, n_leaves, log_block, digests, state_in, state_out);
  else if (slot < N4 + N2)
    mm_sponge_warp<2, DBG>(T2, b0 + 128 * N4 + 16 * (4 * (slot - N4) + sched), cols, leaf_len, n_leaves, log_block, digests, state_in, state_out);
  else
    mm_sponge_warp<1, DBG>(T1, b0 + 128 * N4 + 64 * N2 + 8 * (4 * (slot - N4 - N2) + sched), cols, leaf_len, n_leaves, log_block, digests, state_in, state_out);
}
