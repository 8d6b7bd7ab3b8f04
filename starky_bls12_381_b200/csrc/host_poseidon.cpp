// Host-side Poseidon-12 permutation for the Fiat-Shamir transcript (plonky2 Challenger).
// observe_openings is an inherently sequential sponge over 2*(2C+nq) field elements (up to ~49k chained permutations
// for MillerLoop), so the permutation's *latency* on one host core is on the proof's critical path.  Branch-free
// Goldilocks reduction (data-dependent carries mispredict ~50% of the time) and an AVX2 MDS layer on 32-bit halves;
// a portable scalar path is kept for CPUs without AVX2.  Same function as poseidon.cuh's poseidon_permute.
#include <stdint.h>
#include <string.h>

#include "poseidon_rc.h"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

static const u64 RC[POSEIDON_RC_COUNT] = POSEIDON_RC_TABLE;
static const u64 P = 0xFFFFFFFF00000001ULL, EPS = 0xFFFFFFFFULL;
static const u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

static inline u64 mask_of(bool c) { return 0 - (u64)c; }
static inline u64 red128(u128 x) {   // -> any u64 congruent to x
  u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & EPS;
  u64 t0 = lo - hh;
  t0 -= mask_of(lo < hh) & EPS;
  u64 t1 = hl * EPS, r = t0 + t1;
  r += mask_of(r < t1) & EPS;
  return r;
}
static inline u64 mul(u64 a, u64 b) { return red128((u128)a * b); }
static inline u64 sbox(u64 x) {
  u64 x2 = mul(x, x), x4 = mul(x2, x2), x3 = mul(x2, x);
  return mul(x3, x4);
}
static inline u64 recombine(u64 al, u64 ah) {   // al + ah * 2^32, al, ah < 2^42
  u64 c = (ah >> 32) * EPS, b = (ah & EPS) << 32, t = al + c, v = b + t;
  v += mask_of(v < t) & EPS;
  return v;
}

static void mds_scalar(u64 s[12]) {
  u32 lo[24], hi[24];
  u64 al[12], ah[12];
  for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = (u32)s[i]; hi[i] = hi[i + 12] = (u32)(s[i] >> 32); }
  for (int r = 0; r < 12; r++) { al[r] = 0; ah[r] = 0; }
  for (int i = 0; i < 12; i++)
    for (int r = 0; r < 12; r++) { al[r] += (u64)CIRC[i] * lo[i + r]; ah[r] += (u64)CIRC[i] * hi[i + r]; }
  al[0] += 8 * (u64)lo[0]; ah[0] += 8 * (u64)hi[0];
  for (int r = 0; r < 12; r++) s[r] = recombine(al[r], ah[r]);
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void mds_avx2(u64 s[12]) {
  alignas(32) u32 lo[32], hi[32];
  alignas(32) u64 al[12], ah[12];
  for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = (u32)s[i]; hi[i] = hi[i + 12] = (u32)(s[i] >> 32); }
  __m256i a0 = _mm256_setzero_si256(), a1 = a0, a2 = a0, h0 = a0, h1 = a0, h2 = a0;
  for (int i = 0; i < 12; i++) {
    const __m256i c = _mm256_set1_epi64x(CIRC[i]);
#define LD(p) _mm256_cvtepu32_epi64(_mm_loadu_si128((const __m128i*)(p)))
    a0 = _mm256_add_epi64(a0, _mm256_mul_epu32(LD(lo + i), c));
    a1 = _mm256_add_epi64(a1, _mm256_mul_epu32(LD(lo + i + 4), c));
    a2 = _mm256_add_epi64(a2, _mm256_mul_epu32(LD(lo + i + 8), c));
    h0 = _mm256_add_epi64(h0, _mm256_mul_epu32(LD(hi + i), c));
    h1 = _mm256_add_epi64(h1, _mm256_mul_epu32(LD(hi + i + 4), c));
    h2 = _mm256_add_epi64(h2, _mm256_mul_epu32(LD(hi + i + 8), c));
#undef LD
  }
  _mm256_store_si256((__m256i*)al, a0); _mm256_store_si256((__m256i*)(al + 4), a1); _mm256_store_si256((__m256i*)(al + 8), a2);
  _mm256_store_si256((__m256i*)ah, h0); _mm256_store_si256((__m256i*)(ah + 4), h1); _mm256_store_si256((__m256i*)(ah + 8), h2);
  al[0] += 8 * (u64)lo[0]; ah[0] += 8 * (u64)hi[0];
  for (int r = 0; r < 12; r++) s[r] = recombine(al[r], ah[r]);
}
#endif

template <class Mds>
static inline void permute_with(u64 s[12], Mds mds) {
  for (int r = 0; r < 30; r++) {
    for (int i = 0; i < 12; i++) { u64 v = s[i] + RC[12 * r + i]; v += mask_of(v < s[i]) & EPS; s[i] = v; }
    if (r < 4 || r >= 26) { for (int i = 0; i < 12; i++) s[i] = sbox(s[i]); }
    else s[0] = sbox(s[0]);
    mds(s);
  }
  for (int i = 0; i < 12; i++) s[i] -= mask_of(s[i] >= P) & P;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void permute_avx2(u64 s[12]) { permute_with(s, mds_avx2); }
#endif
static void permute_scalar(u64 s[12]) { permute_with(s, mds_scalar); }

void sb_host_poseidon_permute(u64 s[12]) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) { permute_avx2(s); return; }
#endif
  permute_scalar(s);
}
