// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under starky_bls12_381_b200/ may include this.
//
// CPU restatement of plonky2_field 0.1.1 (git Electron-Labs/plonky2 @ 666f315, un-vendored
// dependency of /root/reference, see Cargo.lock:1502-1515): GoldilocksField and its quadratic
// extension, as used by every reference call site of starky::prover::prove
// (/root/reference/src/aggregate_proof.rs:59,105,138,169,212).  Spec: SURVEY.md Appendix A.1.
// Everything here is kept canonical in [0,p) -- simple and obviously right, not fast.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

namespace orc {

typedef uint64_t u64;
typedef unsigned __int128 u128;

static const u64 GL_P = 0xFFFFFFFF00000001ULL;
static const u64 GL_GEN = 7;                                  // MULTIPLICATIVE_GROUP_GENERATOR = coset shift
static const u64 GL_POW2_GEN = 1753635133440165772ULL;        // 7^((p-1)/2^32), order 2^32

static inline u64 gl_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }
static inline u64 gl_add(u64 a, u64 b) { u128 s = (u128)a + b; return s >= GL_P ? (u64)(s - GL_P) : (u64)s; }
static inline u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a + (GL_P - b); }
static inline u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
// 128-bit product reduced with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p); gl_mul_slow is the definition it must equal
// (checked on edge values + random pairs by tests/test_oracle_kat.py).
static inline u64 gl_mul_slow(u64 a, u64 b) { return (u64)(((u128)a * b) % GL_P); }
static inline u64 gl_reduce128(u128 x) {
  u64 lo = (u64)x, hi = (u64)(x >> 64);
  u64 hi_hi = hi >> 32, hi_lo = hi & 0xFFFFFFFFULL;
  u64 t0 = lo - hi_hi; if (lo < hi_hi) t0 -= 0xFFFFFFFFULL;      // borrow: -2^64 = -(2^32-1)
  u64 t1 = hi_lo * 0xFFFFFFFFULL;
  u64 r = t0 + t1; if (r < t1) r += 0xFFFFFFFFULL;               // carry: +2^64 = +(2^32-1)
  return r >= GL_P ? r - GL_P : r;
}
static inline u64 gl_mul(u64 a, u64 b) { return gl_reduce128((u128)a * b); }
static inline u64 gl_pow(u64 a, u64 e) {
  u64 r = 1;
  while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; }
  return r;
}
static inline u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }
// primitive_root_of_unity(k): POWER_OF_TWO_GENERATOR^(2^(32-k))
static inline u64 gl_root(unsigned log_n) {
  u64 r = GL_POW2_GEN;
  for (unsigned i = log_n; i < 32; i++) r = gl_mul(r, r);
  return r;
}

// F_p[X]/(X^2 - 7)
struct E2 { u64 a, b; };
static inline E2 e2(u64 a, u64 b = 0) { E2 r = {a, b}; return r; }
static inline bool e2_eq(E2 x, E2 y) { return x.a == y.a && x.b == y.b; }
static inline E2 e2_add(E2 x, E2 y) { return e2(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
static inline E2 e2_sub(E2 x, E2 y) { return e2(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
static inline E2 e2_mul(E2 x, E2 y) {
  return e2(gl_add(gl_mul(x.a, y.a), gl_mul(7, gl_mul(x.b, y.b))),
            gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
static inline E2 e2_scale(E2 x, u64 s) { return e2(gl_mul(x.a, s), gl_mul(x.b, s)); }
static inline E2 e2_inv(E2 x) {
  // 1/(a+bX) = (a-bX)/(a^2-7b^2)
  u64 d = gl_inv(gl_sub(gl_mul(x.a, x.a), gl_mul(7, gl_mul(x.b, x.b))));
  return e2(gl_mul(x.a, d), gl_mul(gl_neg(x.b), d));
}
static inline E2 e2_pow(E2 x, u64 e) {
  E2 r = e2(1);
  while (e) { if (e & 1) r = e2_mul(r, x); x = e2_mul(x, x); e >>= 1; }
  return r;
}

static inline unsigned bitrev(unsigned x, unsigned bits) {
  unsigned r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

}  // namespace orc
