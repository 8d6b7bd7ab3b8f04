// Poseidon-12 over Goldilocks (width 12, rate 8, x^7, 4+22+4 rounds) for sm_100a and for the host transcript.
// Replaces plonky2::hash::poseidon::{PoseidonHash, PoseidonPermutation} as used by MerkleTree::new,
// Challenger and fri_proof_of_work inside starky::prover::prove (reference call sites
// /root/reference/src/aggregate_proof.rs:59,...).  One thread owns one 12-word state in registers.
//
// The MDS layer uses the smallness of the circulant (entries <= 41): each state word is split into 32-bit
// halves, both halves are accumulated with 32x(6-bit)+64 IMADs (no per-term reduction) and the two sums are
// recombined and reduced once per output word.  The state is kept *lazy* (any u64) between layers and only
// canonicalised when it leaves the permutation.
#pragma once
#include "gl.cuh"
#include "poseidon_rc.h"

#if defined(__CUDACC__)
// 30 x 12 round constants followed by 12 zeros ("the constants of the round after the last one"): kernels that fold
// the next round's constant addition into the MDS accumulators read row rd + 1 unconditionally.
static __constant__ u64 c_poseidon_rc[POSEIDON_RC_COUNT + 12] = POSEIDON_RC_TABLE;

// (hi:lo) += x * c as ONE IMAD.WIDE.U32 on the FMA pipe: ptxas fuses the mad.lo.cc / madc.hi pair and, unlike
// with mad.wide.u32, keeps a chain of them as written instead of re-associating it into IMAD + IADD3 trees.
// No carry out: callers guarantee the running sum fits 64 bits.
__device__ __forceinline__ void mac32(u32& lo, u32& hi, u32 x, u32 c) {
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(c));
}
// MDS output word from its two half accumulators: al + ah * 2^32 with al, ah < 2^43  ->  lazy u64
__device__ __forceinline__ u64 mds_recombine(u32 al0, u32 al1, u32 ah0, u32 ah1) {
  // t = al + ah1 * (2^64 mod p) < 2^44;  hi(t) += ah0;  a carry out owes 2^64 = eps, and the wrapped value is < 2^44
  asm("{\n\t"
      ".reg .u32 cy;\n\t"
      "mad.lo.cc.u32 %0, %3, 0xFFFFFFFF, %0;\n\t"
      "madc.hi.u32 %1, %3, 0xFFFFFFFF, %1;\n\t"
      "add.cc.u32 %1, %1, %2;\n\t"
      "addc.u32 cy, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, cy, 0xFFFFFFFF, %0;\n\t"
      "madc.hi.u32 %1, cy, 0xFFFFFFFF, %1;\n\t"
      "}"
      : "+r"(al0), "+r"(al1)
      : "r"(ah0), "r"(ah1));
  return ((u64)al1 << 32) | al0;
}
#endif
static const u64 h_poseidon_rc[POSEIDON_RC_COUNT] = POSEIDON_RC_TABLE;

GL_HD u64 poseidon_rc(int i) {
#if defined(__CUDA_ARCH__)
  return c_poseidon_rc[i];
#else
  return h_poseidon_rc[i];
#endif
}

// lazy + canonical -> lazy
GL_HD u64 gl_add_lazy_canon(u64 a, u64 b_canon) {
  u64 s = a + b_canon;
  return (s < a) ? s + GL_EPS : s;
}

GL_HD u64 poseidon_sbox(u64 x) {
  u64 x2 = gl_mul_lazy(x, x);
  u64 x4 = gl_mul_lazy(x2, x2);
  u64 x3 = gl_mul_lazy(x2, x);
  return gl_mul_lazy(x3, x4);
}

// out[r] = sum_i CIRC[i] * s[(i+r)%12] + DIAG[r]*s[r],  CIRC = {17,15,41,16,2,28,13,13,39,18,34,20}, DIAG = {8,0,...}
GL_HD void poseidon_mds(u64 s[12]) {
  const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u32 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    lo[i] = (u32)s[i];
    hi[i] = (u32)(s[i] >> 32);
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
#if defined(__CUDA_ARCH__)
    u32 al0 = 0, al1 = 0, ah0 = 0, ah1 = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      mac32(al0, al1, lo[(i + r) % 12], C[i]);
      mac32(ah0, ah1, hi[(i + r) % 12], C[i]);
    }
    if (r == 0) { mac32(al0, al1, lo[0], 8u); mac32(ah0, ah1, hi[0], 8u); }
    s[r] = mds_recombine(al0, al1, ah0, ah1);
#else
    u64 al = 0, ah = 0;
    for (int i = 0; i < 12; i++) {
      al += (u64)C[i] * lo[(i + r) % 12];
      ah += (u64)C[i] * hi[(i + r) % 12];
    }
    if (r == 0) {
      al += (u64)8 * lo[0];
      ah += (u64)8 * hi[0];
    }
    // value = al + ah * 2^32,  al, ah < 2^41.  ah*2^32 = (ah_lo << 32) + (ah >> 32) * 2^64, 2^64 = EPS (mod p)
    u64 c = (ah >> 32) * GL_EPS;          // < 2^41
    u64 b = (ah & GL_EPS) << 32;          // < 2^64
    u64 t = al + c;                       // < 2^42
    u64 v = b + t;
    if (v < t) v += GL_EPS;
    s[r] = v;                             // lazy
#endif
  }
}

GL_HD void poseidon_permute(u64 s[12]) {
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add_lazy_canon(s[i], poseidon_rc(12 * r + i)));
    poseidon_mds(s);
  }
#pragma unroll 1
  for (int r = 4; r < 26; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_lazy_canon(s[i], poseidon_rc(12 * r + i));
    s[0] = poseidon_sbox(s[0]);
    poseidon_mds(s);
  }
#pragma unroll 1
  for (int r = 26; r < 30; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add_lazy_canon(s[i], poseidon_rc(12 * r + i)));
    poseidon_mds(s);
  }
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}

// two_to_one(l, r): permute [l, r, 0,0,0,0], take the first four words (A.3)
GL_HD void poseidon_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
  u64 s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
  poseidon_permute(s);
  out[0] = s[0]; out[1] = s[1]; out[2] = s[2]; out[3] = s[3];
}
