// Host-side Poseidon-12 permutation for the Fiat-Shamir transcript (plonky2 Challenger).
// observe_openings is an inherently sequential sponge over 2*(2C+nq) field elements (up to ~49k chained permutations
// for MillerLoop), so the permutation's *latency* on one host core is on the proof's critical path.  Branch-free
// Goldilocks reduction (data-dependent carries mispredict ~50% of the time) and an AVX2 MDS layer on 32-bit halves;
// a portable scalar path is kept for CPUs without AVX2.  Same function as poseidon.cuh's poseidon_permute.
#include <stdint.h>
#include <string.h>

#include "poseidon_fast.h"
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

static void permute_scalar(u64 s[12]) {
  for (int r = 0; r < 30; r++) {
    for (int i = 0; i < 12; i++) { u64 v = s[i] + RC[12 * r + i]; v += mask_of(v < s[i]) & EPS; s[i] = v; }
    if (r < 4 || r >= 26) { for (int i = 0; i < 12; i++) s[i] = sbox(s[i]); }
    else s[0] = sbox(s[0]);
    mds_scalar(s);
  }
  for (int i = 0; i < 12; i++) s[i] -= mask_of(s[i] >= P) & P;
}

#if defined(__x86_64__)
// ---- AVX2: the 12-word state lives in three 4-lane vectors; full-round S-boxes, round constants and the MDS layer
// are vectorised, the single partial-round S-box stays scalar (it is a pure latency chain). ----
#define TGT __attribute__((target("avx2"))) static inline
typedef __m256i V;
TGT V vset(u64 x) { return _mm256_set1_epi64x((long long)x); }
TGT V ult(V a, V b) {   // unsigned a < b per lane -> all-ones
  const V S = vset(1ull << 63);
  return _mm256_cmpgt_epi64(_mm256_xor_si256(b, S), _mm256_xor_si256(a, S));
}
TGT V vadd_lazy(V a, V c) {   // lazy + canonical -> lazy
  V s = _mm256_add_epi64(a, c);
  return _mm256_add_epi64(s, _mm256_srli_epi64(ult(s, a), 32));
}
TGT V vmul(V a, V b) {        // lazy x lazy -> lazy
  const V M32 = vset(EPS);
  V ah = _mm256_srli_epi64(a, 32), bh = _mm256_srli_epi64(b, 32);
  V ll = _mm256_mul_epu32(a, b), lh = _mm256_mul_epu32(a, bh), hl = _mm256_mul_epu32(ah, b), hh = _mm256_mul_epu32(ah, bh);
  V mid = _mm256_add_epi64(lh, _mm256_srli_epi64(ll, 32));
  V mid2 = _mm256_add_epi64(hl, _mm256_and_si256(mid, M32));
  V lo = _mm256_or_si256(_mm256_and_si256(ll, M32), _mm256_slli_epi64(mid2, 32));
  V hi = _mm256_add_epi64(_mm256_add_epi64(hh, _mm256_srli_epi64(mid, 32)), _mm256_srli_epi64(mid2, 32));
  V hi_hi = _mm256_srli_epi64(hi, 32), hi_lo = _mm256_and_si256(hi, M32);
  V t0 = _mm256_sub_epi64(_mm256_sub_epi64(lo, hi_hi), _mm256_srli_epi64(ult(lo, hi_hi), 32));
  V t1 = _mm256_sub_epi64(_mm256_slli_epi64(hi_lo, 32), hi_lo);
  V r = _mm256_add_epi64(t0, t1);
  return _mm256_add_epi64(r, _mm256_srli_epi64(ult(r, t1), 32));
}
TGT V vsbox(V x) { V x2 = vmul(x, x), x4 = vmul(x2, x2), x3 = vmul(x2, x); return vmul(x3, x4); }

// CV[w][j][l] = coefficient of input word w in output row 4j + l
alignas(32) static u64 CV[12][3][4];
static const bool cv_ready = [] {
  for (int w = 0; w < 12; w++)
    for (int j = 0; j < 3; j++)
      for (int l = 0; l < 4; l++) {
        int row = 4 * j + l;
        CV[w][j][l] = CIRC[((w - row) % 12 + 12) % 12] + ((w == 0 && row == 0) ? 8 : 0);
      }
  return true;
}();

TGT void vmds(V& s0, V& s1, V& s2) {
  alignas(32) u64 st[12];
  _mm256_store_si256((__m256i*)st, s0); _mm256_store_si256((__m256i*)(st + 4), s1); _mm256_store_si256((__m256i*)(st + 8), s2);
  V al0 = _mm256_setzero_si256(), al1 = al0, al2 = al0, ah0 = al0, ah1 = al0, ah2 = al0;
  for (int w = 0; w < 12; w++) {
    V b = _mm256_set1_epi64x((long long)st[w]), bh = _mm256_srli_epi64(b, 32);
    V c0 = _mm256_load_si256((const __m256i*)CV[w][0]), c1 = _mm256_load_si256((const __m256i*)CV[w][1]),
      c2 = _mm256_load_si256((const __m256i*)CV[w][2]);
    al0 = _mm256_add_epi64(al0, _mm256_mul_epu32(b, c0)); al1 = _mm256_add_epi64(al1, _mm256_mul_epu32(b, c1));
    al2 = _mm256_add_epi64(al2, _mm256_mul_epu32(b, c2));
    ah0 = _mm256_add_epi64(ah0, _mm256_mul_epu32(bh, c0)); ah1 = _mm256_add_epi64(ah1, _mm256_mul_epu32(bh, c1));
    ah2 = _mm256_add_epi64(ah2, _mm256_mul_epu32(bh, c2));
  }
  V al[3] = {al0, al1, al2}, ah[3] = {ah0, ah1, ah2}, out[3];
  for (int j = 0; j < 3; j++) {   // al + ah * 2^32  (mod p), lazy
    V a_hi = _mm256_srli_epi64(ah[j], 32);
    V c = _mm256_sub_epi64(_mm256_slli_epi64(a_hi, 32), a_hi), b = _mm256_slli_epi64(ah[j], 32);
    V t = _mm256_add_epi64(al[j], c), v = _mm256_add_epi64(b, t);
    out[j] = _mm256_add_epi64(v, _mm256_srli_epi64(ult(v, t), 32));
  }
  s0 = out[0]; s1 = out[1]; s2 = out[2];
}

alignas(32) static const u64 RCA[POSEIDON_RC_COUNT] = POSEIDON_RC_TABLE;
__attribute__((target("avx2"))) static void permute_avx2(u64 s[12]) {
  V s0 = _mm256_loadu_si256((const __m256i*)s), s1 = _mm256_loadu_si256((const __m256i*)(s + 4)),
    s2 = _mm256_loadu_si256((const __m256i*)(s + 8));
  for (int r = 0; r < 30; r++) {
    s0 = vadd_lazy(s0, _mm256_load_si256((const __m256i*)(RCA + 12 * r)));
    s1 = vadd_lazy(s1, _mm256_load_si256((const __m256i*)(RCA + 12 * r + 4)));
    s2 = vadd_lazy(s2, _mm256_load_si256((const __m256i*)(RCA + 12 * r + 8)));
    if (r < 4 || r >= 26) { s0 = vsbox(s0); s1 = vsbox(s1); s2 = vsbox(s2); }
    else {
      u64 x = sbox((u64)_mm256_extract_epi64(s0, 0));
      s0 = _mm256_blend_epi32(s0, _mm256_set1_epi64x((long long)x), 0x03);
    }
    vmds(s0, s1, s2);
  }
  _mm256_storeu_si256((__m256i*)s, s0); _mm256_storeu_si256((__m256i*)(s + 4), s1); _mm256_storeu_si256((__m256i*)(s + 8), s2);
  for (int i = 0; i < 12; i++) s[i] -= mask_of(s[i] >= P) & P;
}
#endif

#if defined(__x86_64__)
// ---- AVX-512: the state lives in two 8-lane vectors (words 0..7, words 8..11 + four idle lanes); unsigned compares
// are native mask operations, so every carry / borrow repair is one compare and one masked add. ----
#define T512 __attribute__((target("avx512f,avx512dq,avx512bw,avx512vl"))) static inline
typedef __m512i W;
T512 W wset(u64 x) { return _mm512_set1_epi64((long long)x); }
T512 W wadd_lazy(W a, W c) {   // lazy + canonical -> lazy
  W s = _mm512_add_epi64(a, c);
  return _mm512_mask_add_epi64(s, _mm512_cmplt_epu64_mask(s, a), s, wset(EPS));
}
T512 W wmul(W a, W b) {        // lazy x lazy -> lazy
  const W M32 = wset(EPS);
  W ah = _mm512_srli_epi64(a, 32), bh = _mm512_srli_epi64(b, 32);
  W ll = _mm512_mul_epu32(a, b), lh = _mm512_mul_epu32(a, bh), hl = _mm512_mul_epu32(ah, b), hh = _mm512_mul_epu32(ah, bh);
  W mid = _mm512_add_epi64(lh, _mm512_srli_epi64(ll, 32));
  W mid2 = _mm512_add_epi64(hl, _mm512_and_si512(mid, M32));
  W lo = _mm512_or_si512(_mm512_and_si512(ll, M32), _mm512_slli_epi64(mid2, 32));
  W hi = _mm512_add_epi64(_mm512_add_epi64(hh, _mm512_srli_epi64(mid, 32)), _mm512_srli_epi64(mid2, 32));
  W hi_hi = _mm512_srli_epi64(hi, 32), hi_lo = _mm512_and_si512(hi, M32);
  W t0 = _mm512_sub_epi64(lo, hi_hi);
  t0 = _mm512_mask_sub_epi64(t0, _mm512_cmplt_epu64_mask(lo, hi_hi), t0, M32);
  W t1 = _mm512_sub_epi64(_mm512_slli_epi64(hi_lo, 32), hi_lo);
  W r = _mm512_add_epi64(t0, t1);
  return _mm512_mask_add_epi64(r, _mm512_cmplt_epu64_mask(r, t1), r, M32);
}
T512 W wsbox(W x) { W x2 = wmul(x, x), x4 = wmul(x2, x2), x3 = wmul(x2, x); return wmul(x3, x4); }

// CW[w][j][l] = coefficient of input word w in output row 8j + l  (rows 12..15 do not exist: zero)
alignas(64) static u64 CW[12][2][8];
static const bool cw_ready = [] {
  for (int w = 0; w < 12; w++)
    for (int j = 0; j < 2; j++)
      for (int l = 0; l < 8; l++) {
        int row = 8 * j + l;
        CW[w][j][l] = row < 12 ? CIRC[((w - row) % 12 + 12) % 12] + ((w == 0 && row == 0) ? 8 : 0) : 0;
      }
  return true;
}();

T512 void wmds(W& s0, W& s1) {
  alignas(64) u64 st[16];
  _mm512_store_si512((void*)st, s0); _mm512_store_si512((void*)(st + 8), s1);
  W al0 = _mm512_setzero_si512(), al1 = al0, ah0 = al0, ah1 = al0;
  for (int w = 0; w < 12; w++) {
    W b = _mm512_set1_epi64((long long)st[w]), bh = _mm512_srli_epi64(b, 32);
    W c0 = _mm512_load_si512((const void*)CW[w][0]), c1 = _mm512_load_si512((const void*)CW[w][1]);
    al0 = _mm512_add_epi64(al0, _mm512_mul_epu32(b, c0)); al1 = _mm512_add_epi64(al1, _mm512_mul_epu32(b, c1));
    ah0 = _mm512_add_epi64(ah0, _mm512_mul_epu32(bh, c0)); ah1 = _mm512_add_epi64(ah1, _mm512_mul_epu32(bh, c1));
  }
  W al[2] = {al0, al1}, ah[2] = {ah0, ah1}, out[2];
  for (int j = 0; j < 2; j++) {   // al + ah * 2^32  (mod p), lazy
    W a_hi = _mm512_srli_epi64(ah[j], 32);
    W c = _mm512_sub_epi64(_mm512_slli_epi64(a_hi, 32), a_hi), b = _mm512_slli_epi64(ah[j], 32);
    W t = _mm512_add_epi64(al[j], c), v = _mm512_add_epi64(b, t);
    out[j] = _mm512_mask_add_epi64(v, _mm512_cmplt_epu64_mask(v, t), v, wset(EPS));
  }
  s0 = out[0]; s1 = out[1];
}

alignas(64) static u64 RCW[30][16];
static const bool rcw_ready = [] {
  for (int r = 0; r < 30; r++)
    for (int i = 0; i < 16; i++) RCW[r][i] = i < 12 ? RC[12 * r + i] : 0;
  return true;
}();

__attribute__((target("avx512f,avx512dq,avx512bw,avx512vl"))) static void permute_avx512(u64 s[12]) {
  W s0 = _mm512_loadu_si512((const void*)s);
  W s1 = _mm512_maskz_loadu_epi64(0x0F, (const void*)(s + 8));
  for (int r = 0; r < 30; r++) {
    s0 = wadd_lazy(s0, _mm512_load_si512((const void*)RCW[r]));
    s1 = wadd_lazy(s1, _mm512_load_si512((const void*)(RCW[r] + 8)));
    if (r < 4 || r >= 26) { s0 = wsbox(s0); s1 = wsbox(s1); }
    else {
      u64 x = sbox((u64)_mm_cvtsi128_si64(_mm512_castsi512_si128(s0)));
      s0 = _mm512_mask_set1_epi64(s0, 0x01, (long long)x);
    }
    wmds(s0, s1);
  }
  _mm512_storeu_si512((void*)s, s0);
  _mm512_mask_storeu_epi64((void*)(s + 8), 0x0F, s1);
  for (int i = 0; i < 12; i++) s[i] -= mask_of(s[i] >= P) & P;
}
// ---- AVX-512 full rounds + SPARSE partial rounds (poseidon_fast.h, the factorisation plonky2 ships as FAST_PARTIAL_*).
// The 22 partial rounds of the dense form above cost a full vector MDS each although only word 0 went through the
// S-box; in sparse form a partial round is
//     y = x0^7 + POST[r];   x0' = 25 y + sum_i WHAT[r][i] x_i;   x_i' = x_i + VS[r][i] y
// and with  sum_i WHAT[r+1][i] x_i' = sum_i WHAT[r+1][i] x_i + U[r+1] y  the dot product of the NEXT round is taken from
// the values before this round's update, so the only thing on the round's dependency chain is the S-box and one
// multiply-add (~40 cycles); the eleven multiply-adds of the update and the eleven of the look-ahead dot product run
// beside it.  128-bit sums of products are kept as (sum of low words, sum of high words) and folded once.
static const u64 F_FIRST[12] = POSEIDON_FAST_FIRST, F_POST[22] = POSEIDON_FAST_POST, F_INIT[121] = POSEIDON_FAST_INIT,
                 F_WHAT[22 * 11] = POSEIDON_FAST_WHAT, F_VS[22 * 11] = POSEIDON_FAST_VS, F_U[22] = POSEIDON_FAST_U;

alignas(64) static u64 VSW[22][16];      // VS[r][0..10] padded to two vectors
static const bool vsw_ready = [] {
  for (int r = 0; r < 22; r++)
    for (int i = 0; i < 16; i++) VSW[r][i] = i < 11 ? F_VS[11 * r + i] : 0;
  return true;
}();
// lazy -> canonical
T512 W wcanon(W a) { return _mm512_mask_sub_epi64(a, _mm512_cmpge_epu64_mask(a, wset(P)), a, wset(P)); }
static inline u64 add_lazy(u64 a, u64 c) { u64 v = a + c; v += mask_of(v < a) & EPS; return v; }   // lazy + any -> lazy (c <= 2^64 - 2^32)
// sum_i c[i] * x[i] (n <= 16 terms) -> lazy u64
static inline u64 dot_lazy(const u64* c, const u64* x, int n) {
  u128 lo = 0, hi = 0;
  for (int i = 0; i < n; i++) { u128 m = (u128)c[i] * x[i]; lo += (u64)m; hi += (u64)(m >> 64); }
  // lo + hi * 2^64 with hi < 2^68:  2^64 = EPS (mod p), hi * EPS < 2^100
  return red128(lo + hi * (u128)EPS);
}

__attribute__((target("avx512f,avx512dq,avx512bw,avx512vl"))) static void permute_avx512_sparse(u64 s[12]) {
  W s0 = _mm512_loadu_si512((const void*)s);
  W s1 = _mm512_maskz_loadu_epi64(0x0F, (const void*)(s + 8));
  for (int r = 0; r < 4; r++) {
    s0 = wadd_lazy(s0, _mm512_load_si512((const void*)RCW[r]));
    s1 = wadd_lazy(s1, _mm512_load_si512((const void*)(RCW[r] + 8)));
    s0 = wsbox(s0); s1 = wsbox(s1);
    wmds(s0, s1);
  }
  alignas(64) u64 x[16];
  _mm512_store_si512((void*)x, s0); _mm512_store_si512((void*)(x + 8), s1);
  // x += FIRST;  x[1..] = INIT . x[1..]
  u64 t[12];
  for (int i = 0; i < 12; i++) t[i] = add_lazy(x[i], F_FIRST[i]);
  x[0] = t[0];
  for (int j = 0; j < 11; j++) x[j + 1] = dot_lazy(F_INIT + 11 * j, t + 1, 11);
  // partial rounds: x0 and the dot products in scalar registers, the eleven updates x_i += VS[r][i] y as two vectors
  // (words 1..8 and 9..11; stored after every round so that the scalar dot product of the next round can read them)
  u64 x0 = x[0], y_prev = 0;
  u64 A = dot_lazy(F_WHAT, x + 1, 11);                      // sum_i WHAT[0][i] x_i
  W xa = _mm512_loadu_si512((const void*)(x + 1));
  W xb = _mm512_maskz_loadu_epi64(0x07, (const void*)(x + 9));
  for (int r = 0; r < 22; r++) {
    const u64 D = r ? red128((u128)F_U[r] * y_prev + A) : A;   // sum_i WHAT[r][i] x_i(r)
    const u64 y = add_lazy(sbox(x0), F_POST[r]);
    const u64 A_next = r + 1 < 22 ? dot_lazy(F_WHAT + 11 * (r + 1), x + 1, 11) : 0;   // from the values BEFORE this update
    x0 = red128((u128)25 * y + D);
    const W yv = wset(y);
    xa = wadd_lazy(xa, wcanon(wmul(_mm512_load_si512((const void*)VSW[r]), yv)));
    xb = wadd_lazy(xb, wcanon(wmul(_mm512_load_si512((const void*)(VSW[r] + 8)), yv)));
    _mm512_storeu_si512((void*)(x + 1), xa);
    _mm512_mask_storeu_epi64((void*)(x + 9), 0x07, xb);
    A = A_next; y_prev = y;
  }
  x[0] = x0;
  s0 = _mm512_load_si512((const void*)x);
  s1 = _mm512_maskz_load_epi64(0x0F, (const void*)(x + 8));
  for (int r = 26; r < 30; r++) {
    s0 = wadd_lazy(s0, _mm512_load_si512((const void*)RCW[r]));
    s1 = wadd_lazy(s1, _mm512_load_si512((const void*)(RCW[r] + 8)));
    s0 = wsbox(s0); s1 = wsbox(s1);
    wmds(s0, s1);
  }
  _mm512_storeu_si512((void*)s, s0);
  _mm512_mask_storeu_epi64((void*)(s + 8), 0x0F, s1);
  for (int i = 0; i < 12; i++) s[i] -= mask_of(s[i] >= P) & P;
}
// ---- Hybrid form (variant 4): AVX-512 IFMA + BMI2.  What the variants above leave on the table (rdtsc per phase, one
// core): a full round is throughput-bound on the two 512-bit ports (two wsbox = 176 uops, half of the second vector
// idle) and a partial round executes ~360 instructions around a 50-cycle dependency chain.  Here
//   * full rounds: words 0..7 in one vector, words 8..11 in scalar registers (mulx) -- the scalar ports are otherwise
//     idle, so the four scalar S-boxes run beside the vector one;
//   * the MDS layer is 12 broadcast loads per half and one fused multiply-add per (word, half, row block):
//     vpmadd52luq (32-bit half x coefficient <= 49 fits the 52-bit product), accumulators seeded with the next round's
//     constants, four chains per accumulator;
//   * partial rounds: mulx with three-word carry-chain sums for the look-ahead dot product, the eleven updates as two
//     vector multiplies stored to a 64-byte aligned array with full-width stores (a masked store does not forward to
//     the scalar loads of the next dot product).
#define THYB __attribute__((target("avx512f,avx512dq,avx512bw,avx512vl,avx512ifma,bmi2")))
typedef __m256i Y;
// GCC materialises every carry of the _addcarry_u64 / _subborrow_u64 forms with setb + movzbl (a partial round compiled to
// ~300 instructions); the carry sequences are written out instead: `sbb r32, r32` turns the flag into the 0 / 2^32-1 mask.
THYB static inline u64 hred(u64 lo, u64 hi) {   // lo + hi * 2^64 -> lazy u64
  const u64 hh = hi >> 32, hl = (u32)hi, t1 = (hl << 32) - hl;
  u64 m;
  asm("sub %[hh], %[lo]\n\t"
      "sbb %k[m], %k[m]\n\t"     // borrow: wrapped by 2^64 = EPS (mod p)
      "sub %[m], %[lo]\n\t"
      "add %[t1], %[lo]\n\t"
      "sbb %k[m], %k[m]\n\t"
      "add %[m], %[lo]"
      : [lo] "+r"(lo), [m] "=&r"(m) : [hh] "r"(hh), [t1] "r"(t1) : "cc");
  return lo;
}
THYB static inline u64 hmul(u64 a, u64 b) { unsigned long long hi; const u64 lo = _mulx_u64(a, b, &hi); return hred(lo, hi); }
THYB static inline u64 hsbox(u64 x) { const u64 x2 = hmul(x, x), x4 = hmul(x2, x2), x3 = hmul(x2, x); return hmul(x3, x4); }
// sum_{i<11} c[i] * x[i] -> lazy u64.  Two carry chains (even / odd terms), each lo + hi 2^64 + top 2^128 with
// 2^128 = -2^32 (mod p)
#define HDOT_TERM(i, LO, HI, TOP)                                                              \
  asm("mulx %[xm], %[l], %[h]\n\tadd %[l], %[lo]\n\tadc %[h], %[hi]\n\tadc $0, %[top]"         \
      : [lo] "+r"(LO), [hi] "+r"(HI), [top] "+r"(TOP), [l] "=&r"(l), [h] "=&r"(h) : "d"(c[i]), [xm] "m"(x[i]) : "cc")
THYB static inline u64 hdot11(const u64* c, const u64* x) {
  u64 lo0 = 0, hi0 = 0, top0 = 0, lo1 = 0, hi1 = 0, top1 = 0, l, h;
  HDOT_TERM(0, lo0, hi0, top0); HDOT_TERM(1, lo1, hi1, top1); HDOT_TERM(2, lo0, hi0, top0); HDOT_TERM(3, lo1, hi1, top1);
  HDOT_TERM(4, lo0, hi0, top0); HDOT_TERM(5, lo1, hi1, top1); HDOT_TERM(6, lo0, hi0, top0); HDOT_TERM(7, lo1, hi1, top1);
  HDOT_TERM(8, lo0, hi0, top0); HDOT_TERM(9, lo1, hi1, top1); HDOT_TERM(10, lo0, hi0, top0);
  asm("add %[a], %[lo]\n\tadc %[b], %[hi]\n\tadc %[c], %[top]" : [lo] "+r"(lo0), [hi] "+r"(hi0), [top] "+r"(top0) : [a] "r"(lo1), [b] "r"(hi1), [c] "r"(top1) : "cc");
  u64 r = hred(lo0, hi0), m;
  top0 <<= 32;
  asm("sub %[t], %[r]\n\tsbb %k[m], %k[m]\n\tsub %[m], %[r]" : [r] "+r"(r), [m] "=&r"(m) : [t] "r"(top0) : "cc");
  return r;
}
// a * b + c with a small enough that the high word is < 2^32 (a <= 2^32): lo + h * EPS
THYB static inline u64 hmad_small(u64 a, u64 b, u64 c) {
  unsigned long long h;
  u64 l = _mulx_u64(a, b, &h), m;
  asm("add %[c], %[l]\n\tadc $0, %[h]" : [l] "+r"(l), [h] "+r"(h) : [c] "r"(c) : "cc");
  const u64 t1 = (h << 32) - h;
  asm("add %[t1], %[l]\n\tsbb %k[m], %k[m]\n\tadd %[m], %[l]" : [l] "+r"(l), [m] "=&r"(m) : [t1] "r"(t1) : "cc");
  return l;
}
THYB static inline u64 hmad(u64 a, u64 b, u64 c) {   // a * b + c -> lazy
  unsigned long long h;
  u64 l = _mulx_u64(a, b, &h);
  asm("add %[c], %[l]\n\tadc $0, %[h]" : [l] "+r"(l), [h] "+r"(h) : [c] "r"(c) : "cc");   // h <= 2^64 - 2, no overflow
  return hred(l, h);
}

// MDS tables: column w of the matrix for rows 0..7 (one vector) and rows 8..11 (half a vector)
alignas(64) static u64 HC8[12][8], HC4[12][4];
// accumulator seeds: seed s = 0..7 -> the constants added before the S-box of round s+1 (s < 3), FIRST (s = 3),
// rounds 27..29 (s = 4..6), nothing (s = 7); split into 32-bit halves
alignas(64) static u64 HSL8[8][8], HSH8[8][8], HSL4[8][4], HSH4[8][4];
static u64 HU2[22];                       // U2[r] = WHAT[r] . VS[r-2]
static const bool hyb_ready = [] {
  for (int r = 2; r < 22; r++) {
    u64 acc = 0;
    for (int i = 0; i < 11; i++) {
      u64 m = mul(F_WHAT[11 * r + i], F_VS[11 * (r - 2) + i]);
      m -= mask_of(m >= P) & P;
      acc = add_lazy(acc, m);
    }
    HU2[r] = acc - (mask_of(acc >= P) & P);
  }
  for (int w = 0; w < 12; w++)
    for (int row = 0; row < 12; row++) {
      const u64 c = CIRC[((w - row) % 12 + 12) % 12] + ((w == 0 && row == 0) ? 8 : 0);
      if (row < 8) HC8[w][row] = c; else HC4[w][row - 8] = c;
    }
  for (int sd = 0; sd < 8; sd++)
    for (int i = 0; i < 12; i++) {
      u64 c = 0;
      if (sd < 3) c = RC[12 * (sd + 1) + i];
      else if (sd == 3) c = F_FIRST[i];
      else if (sd < 7) c = RC[12 * (27 + sd - 4) + i];
      if (i < 8) { HSL8[sd][i] = c & EPS; HSH8[sd][i] = c >> 32; } else { HSL4[sd][i - 8] = c & EPS; HSH4[sd][i - 8] = c >> 32; }
    }
  return true;
}();

THYB static inline W hrecombine8(W al, W ah) {      // al + ah * 2^32 (mod p), al, ah < 2^43 -> lazy
  const W a_hi = _mm512_srli_epi64(ah, 32);
  const W c = _mm512_sub_epi64(_mm512_slli_epi64(a_hi, 32), a_hi), b = _mm512_slli_epi64(ah, 32);
  const W t = _mm512_add_epi64(al, c), v = _mm512_add_epi64(b, t);
  return _mm512_mask_add_epi64(v, _mm512_cmplt_epu64_mask(v, t), v, wset(EPS));
}
THYB static inline Y hrecombine4(Y al, Y ah) {
  const Y a_hi = _mm256_srli_epi64(ah, 32);
  const Y c = _mm256_sub_epi64(_mm256_slli_epi64(a_hi, 32), a_hi), b = _mm256_slli_epi64(ah, 32);
  const Y t = _mm256_add_epi64(al, c), v = _mm256_add_epi64(b, t);
  return _mm256_mask_add_epi64(v, _mm256_cmplt_epu64_mask(v, t), v, _mm256_set1_epi64x((long long)EPS));
}

// one full round on (v = words 0..7, q[0..3] = words 8..11): S-box everywhere, MDS, + the constants of seed `sd`
THYB static inline void hfull(W& v, u64 (&q)[4], int sd) {
  alignas(64) u64 lo[16], hi[16];
  v = wsbox(v);
  _mm512_store_si512((void*)lo, _mm512_and_si512(v, wset(EPS)));
  _mm512_store_si512((void*)hi, _mm512_srli_epi64(v, 32));
#pragma GCC unroll 4
  for (int k = 0; k < 4; k++) { const u64 t = hsbox(q[k]); lo[8 + k] = t & EPS; hi[8 + k] = t >> 32; }
  // four accumulation chains per half (a chain of twelve dependent vpmadd52luq would be 48 cycles)
  W al[4], ah[4];
  Y bl[4], bh[4];
  al[0] = _mm512_load_si512((const void*)HSL8[sd]); ah[0] = _mm512_load_si512((const void*)HSH8[sd]);
  bl[0] = _mm256_load_si256((const __m256i*)HSL4[sd]); bh[0] = _mm256_load_si256((const __m256i*)HSH4[sd]);
  for (int k = 1; k < 4; k++) { al[k] = ah[k] = _mm512_setzero_si512(); bl[k] = bh[k] = _mm256_setzero_si256(); }
#pragma GCC unroll 12
  for (int w = 0; w < 12; w++) {
    const W c = _mm512_load_si512((const void*)HC8[w]);
    const Y d = _mm256_load_si256((const __m256i*)HC4[w]);
    const W l = _mm512_set1_epi64((long long)lo[w]), h = _mm512_set1_epi64((long long)hi[w]);
    al[w & 3] = _mm512_madd52lo_epu64(al[w & 3], l, c); ah[w & 3] = _mm512_madd52lo_epu64(ah[w & 3], h, c);
    bl[w & 3] = _mm256_madd52lo_epu64(bl[w & 3], _mm512_castsi512_si256(l), d);
    bh[w & 3] = _mm256_madd52lo_epu64(bh[w & 3], _mm512_castsi512_si256(h), d);
  }
  v = hrecombine8(_mm512_add_epi64(_mm512_add_epi64(al[0], al[1]), _mm512_add_epi64(al[2], al[3])),
                  _mm512_add_epi64(_mm512_add_epi64(ah[0], ah[1]), _mm512_add_epi64(ah[2], ah[3])));
  const Y bls = _mm256_add_epi64(_mm256_add_epi64(bl[0], bl[1]), _mm256_add_epi64(bl[2], bl[3]));
  const Y bhs = _mm256_add_epi64(_mm256_add_epi64(bh[0], bh[1]), _mm256_add_epi64(bh[2], bh[3]));
  alignas(32) u64 out[4];
  _mm256_store_si256((__m256i*)out, hrecombine4(bls, bhs));
  q[0] = out[0]; q[1] = out[1]; q[2] = out[2]; q[3] = out[3];
}

// LA = how many rounds ahead the dot product of the partial rounds is taken (1 or 2)
template <int LA> THYB static void permute_hybrid(u64 s[12]) {
  W v = wadd_lazy(_mm512_loadu_si512((const void*)s), _mm512_load_si512((const void*)RCW[0]));
  u64 q[4];
  for (int k = 0; k < 4; k++) q[k] = add_lazy(s[8 + k], RC[8 + k]);
  for (int r = 0; r < 4; r++) hfull(v, q, r);                  // the fourth one adds FIRST
  // x[1..] = INIT . x[1..]
  alignas(64) u64 t[16], xs[16];
  _mm512_store_si512((void*)t, v);
  t[8] = q[0]; t[9] = q[1]; t[10] = q[2]; t[11] = q[3];
  u64 x0 = t[0];
  for (int j = 0; j < 11; j++) xs[j] = hdot11(F_INIT + 11 * j, t + 1);
  for (int j = 11; j < 16; j++) xs[j] = 0;
  // look-ahead: with x(r) = the words before round r's update and A(r) = WHAT[r] . x(r),
  //   A(r+1) = WHAT[r+1] . x(r) + U[r+1] y(r)          A(r+2) = WHAT[r+2] . x(r) + U2[r+2] y(r) + U[r+2] y(r+1)
  // so the dot product needed by round r+LA reads the vector stores of round r-1: the store -> scalar-load forwarding,
  // the multiply of the update and the carry chains of the dot product have LA rounds of the S-box chain to finish.
  u64 B_cur = hdot11(F_WHAT, xs);                           // round 0: A(0), no correction
  u64 B_next = LA == 2 ? hdot11(F_WHAT + 11, xs) : 0;       // round 1: WHAT[1] . x(0); + U[1] y(0) when used
  u64 y_prev = 0;
  W xa = _mm512_load_si512((const void*)xs), xb = _mm512_load_si512((const void*)(xs + 8));
  for (int r = 0; r < 22; r++) {
    const u64 D = r ? hmad(F_U[r], y_prev, B_cur) : B_cur;  // A(r)
    const u64 y = add_lazy(hsbox(x0), F_POST[r]);
    x0 = hmad_small(25, y, D);
    u64 B2 = 0;                                             // xs = x(r): before this round's update
    if (LA == 2) { if (r + 2 < 22) B2 = hmad(HU2[r + 2], y, hdot11(F_WHAT + 11 * (r + 2), xs)); }
    else if (r + 1 < 22) B2 = hdot11(F_WHAT + 11 * (r + 1), xs);
    const W yv = wset(y);
    xa = wadd_lazy(xa, wcanon(wmul(_mm512_load_si512((const void*)VSW[r]), yv)));
    xb = wadd_lazy(xb, wcanon(wmul(_mm512_load_si512((const void*)(VSW[r] + 8)), yv)));
    _mm512_store_si512((void*)xs, xa);
    _mm512_store_si512((void*)(xs + 8), xb);
    y_prev = y;
    if (LA == 2) { B_cur = B_next; B_next = B2; } else B_cur = B2;
  }
  // words 0..7 = x0, xs[0..6]; words 8..11 = xs[7..10]; + the constants of round 26
  v = wadd_lazy(_mm512_alignr_epi64(xa, wset(x0), 7), _mm512_load_si512((const void*)RCW[26]));
  for (int k = 0; k < 4; k++) q[k] = add_lazy(xs[7 + k], RC[12 * 26 + 8 + k]);
  for (int r = 0; r < 4; r++) hfull(v, q, 4 + r);
  v = wcanon(v);
  _mm512_storeu_si512((void*)s, v);
  for (int k = 0; k < 4; k++) s[8 + k] = q[k] - (mask_of(q[k] >= P) & P);
}
#endif

// variant: 0 = scalar, 1 = AVX2, 2 = AVX-512 dense, 3 = AVX-512 full rounds + sparse partial rounds, 4, 5 = hybrid (AVX-512 IFMA +
// BMI2; dot products one / two rounds ahead) (tests compare them; returns 0 if the CPU lacks the extension)
#if defined(__x86_64__)
static bool have_hybrid() {
  static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512bw") &&
                         __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512ifma") && __builtin_cpu_supports("bmi2") &&
                         cw_ready && rcw_ready && vsw_ready && hyb_ready;
  return ok;
}
#endif

extern "C" int sb_host_poseidon_permute_variant(u64 s[12], int variant) {
#if defined(__x86_64__)
  if (variant == 4 || variant == 5) {
    if (!have_hybrid()) return 0;
    if (variant == 4) permute_hybrid<1>(s); else permute_hybrid<2>(s);
    return 1;
  }
  if (variant == 2 || variant == 3) {
    if (!(__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512bw") &&
          __builtin_cpu_supports("avx512vl"))) return 0;
    if (variant == 3) permute_avx512_sparse(s); else permute_avx512(s);
    return 1;
  }
  if (variant == 1) { if (!__builtin_cpu_supports("avx2")) return 0; permute_avx2(s); return 1; }
#endif
  if (variant != 0) return 0;
  permute_scalar(s); return 1;
}

void sb_host_poseidon_permute(u64 s[12]) {
#if defined(__x86_64__)
  static const bool have_avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") &&
                                  __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") && cw_ready && rcw_ready && vsw_ready;
  if (have_hybrid()) { permute_hybrid<1>(s); return; }
  if (have_avx512) { permute_avx512_sparse(s); return; }
  static const bool have_avx2 = __builtin_cpu_supports("avx2") && cv_ready;
  if (have_avx2) { permute_avx2(s); return; }
#endif
  permute_scalar(s);
}
