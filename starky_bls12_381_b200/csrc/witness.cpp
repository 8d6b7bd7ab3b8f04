// Witness generation in C++: the reference's generate_trace / fill_trace_* functions restated for the
// host side of the drop-in, so that a caller can hand the prover a few field elements instead of a multi-GB trace.
//   fp.rs:185-428   addition / subtraction / multiply_single / reduce_single / range check / 12x12-limb multiplication / reduction
//   fp2.rs:187-456  Fp2 addition, subtraction, multiplication, +- with reduction, non-residue multiplication
//   fp6.rs:124-303  Fp6 addition, subtraction, +- with reduction, non-residue multiplication, multiplication
//   fp12.rs:132-426 multiply_by_014, Fp12 multiplication, cyclotomic square / exponentiation, Frobenius map, conjugate
//   fp12_mul.rs:44-48, ecc_aggregate.rs:37-82, calc_pairing_precomp.rs:150-366, miller_loop.rs:87-160,
//   final_exponentiate.rs:137-281: generate_trace of the five starks
// over the BLS12-381 tower of native.rs (quirks kept because they show in the trace: add_fp subtracts p at most once,
// -x is p - x).  Column offsets come from the reference's constants (witness_offsets.h, generated from
// witness/offsets.json).  Cells are written as row-major uint32_t -- every cell the reference writes is a u32 limb, a carry
// or a bit (utils.rs:7-19) -- which is SB_TRACE_ROWMAJOR_U32, half the PCIe bytes of the u64 layouts.
// The Python restatement (starky_bls12_381_b200/witness) is the checker: tests/test_witness_cpp.py compares cell for cell.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <string>
#include <vector>

#include "../../include/starky_b200.h"
#include "witness_offsets.h"

namespace {

typedef uint32_t u32;
typedef uint64_t u64;

// ---------------------------------------------------------------------------------------------------------
// fixed-width unsigned integers: 26 little-endian u32 limbs (832 bits: products of two 384-bit values plus slack)
// ---------------------------------------------------------------------------------------------------------
struct Big {
  static const int N = 26;
  u32 w[N];
  Big() { memset(w, 0, sizeof(w)); }
  explicit Big(u64 v) { memset(w, 0, sizeof(w)); w[0] = (u32)v; w[1] = (u32)(v >> 32); }
  static Big from_limbs(const u32* l, int n) { Big b; memcpy(b.w, l, 4 * n); return b; }
  bool is_zero() const { for (int i = 0; i < N; i++) if (w[i]) return false; return true; }
  int top() const { for (int i = N - 1; i >= 0; i--) if (w[i]) return i; return -1; }
};
int cmp(const Big& a, const Big& b) {
  for (int i = Big::N - 1; i >= 0; i--) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
  return 0;
}
Big add(const Big& a, const Big& b) {
  Big r; u64 c = 0;
  for (int i = 0; i < Big::N; i++) { c += (u64)a.w[i] + b.w[i]; r.w[i] = (u32)c; c >>= 32; }
  if (c) throw std::overflow_error("witness: Big addition overflow");
  return r;
}
Big sub(const Big& a, const Big& b) {          // a >= b
  Big r; int64_t c = 0;
  for (int i = 0; i < Big::N; i++) { c += (int64_t)a.w[i] - b.w[i]; r.w[i] = (u32)c; c >>= 32; }
  if (c) throw std::underflow_error("witness: Big subtraction underflow");
  return r;
}
Big mul(const Big& a, const Big& b) {
  Big r;
  const int ta = a.top(), tb = b.top();
  if (ta < 0 || tb < 0) return r;
  if (ta + tb + 2 > Big::N) throw std::overflow_error("witness: Big product overflow");
  for (int i = 0; i <= ta; i++) {
    u64 c = 0;
    for (int j = 0; j <= tb; j++) { c += (u64)a.w[i] * b.w[j] + r.w[i + j]; r.w[i + j] = (u32)c; c >>= 32; }
    r.w[i + tb + 1] = (u32)c;
  }
  return r;
}
Big mul_small(const Big& a, u32 k) { return mul(a, Big((u64)k)); }
Big shl_limbs(const Big& a, int k) {
  Big r;
  for (int i = Big::N - 1; i >= k; i--) r.w[i] = a.w[i - k];
  for (int i = Big::N - k; i < Big::N; i++) if (a.w[i]) throw std::overflow_error("witness: Big shift overflow");
  return r;
}
// x = q * m + r, 0 <= r < m: Knuth's algorithm D on 32-bit limbs (the generators divide ~10^5 times per large trace)
void divmod(const Big& x, const Big& m, Big& q, Big& r) {
  q = Big(); r = Big();
  const int n = m.top() + 1, tx = x.top() + 1;
  if (n == 0) throw std::domain_error("witness: division by zero");
  if (tx < n) { r = x; return; }
  if (n == 1) {
    u64 rem = 0;
    for (int i = tx - 1; i >= 0; i--) { const u64 cur = (rem << 32) | x.w[i]; q.w[i] = (u32)(cur / m.w[0]); rem = cur % m.w[0]; }
    r.w[0] = (u32)rem;
    return;
  }
  const int sh = __builtin_clz(m.w[n - 1]);
  u32 v[Big::N], u[Big::N + 1];
  for (int i = n - 1; i > 0; i--) v[i] = sh ? (m.w[i] << sh) | (m.w[i - 1] >> (32 - sh)) : m.w[i];
  v[0] = m.w[0] << sh;
  u[tx] = sh ? x.w[tx - 1] >> (32 - sh) : 0;
  for (int i = tx - 1; i > 0; i--) u[i] = sh ? (x.w[i] << sh) | (x.w[i - 1] >> (32 - sh)) : x.w[i];
  u[0] = x.w[0] << sh;
  for (int j = tx - n; j >= 0; j--) {
    const u64 num = ((u64)u[j + n] << 32) | u[j + n - 1];
    u64 qh = num / v[n - 1], rh = num % v[n - 1];
    while (qh >> 32 || qh * v[n - 2] > ((rh << 32) | u[j + n - 2])) {
      qh--; rh += v[n - 1];
      if (rh >> 32) break;
    }
    int64_t borrow = 0;
    u64 carry = 0;
    for (int i = 0; i < n; i++) {
      const u64 pr = qh * v[i] + carry;
      carry = pr >> 32;
      const int64_t t = (int64_t)u[i + j] - (int64_t)(u32)pr + borrow;
      u[i + j] = (u32)t;
      borrow = t >> 32;
    }
    const int64_t t = (int64_t)u[j + n] - (int64_t)carry + borrow;
    u[j + n] = (u32)t;
    if (t < 0) {                                   // qh was one too large: add the divisor back
      qh--;
      u64 c = 0;
      for (int i = 0; i < n; i++) { c += (u64)u[i + j] + v[i]; u[i + j] = (u32)c; c >>= 32; }
      u[j + n] += (u32)c;
    }
    q.w[j] = (u32)qh;
  }
  for (int i = 0; i < n; i++) r.w[i] = sh ? (u[i] >> sh) | ((u64)u[i + 1] << (32 - sh)) : u[i];
}

const u32 P_LIMBS[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                         0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};   // native.rs:12-14
const Big& MODP() { static const Big p = Big::from_limbs(P_LIMBS, 12); return p; }
const Big& MODP2() { static const Big p2 = mul(MODP(), MODP()); return p2; }
const Big& RC_ADD() {            // fp.rs:1343: 2^382 - p
  static const Big v = [] { Big t; t.w[11] = 1u << 30; return sub(t, MODP()); }();
  return v;
}

// ---- Fp / Fp2 / Fp6 / Fp12 of native.rs (values, not traces) ----
typedef Big Fp;
struct Fp2 { Fp c[2]; };
struct Fp6 { Fp c[6]; };
struct Fp12 { Fp c[12]; };
Fp fp_add(const Fp& x, const Fp& y) { Fp s = add(x, y); return cmp(s, MODP()) >= 0 ? sub(s, MODP()) : s; }   // native.rs:452-468
Fp fp_mod(const Big& x) { Big q, r; divmod(x, MODP(), q, r); return r; }
Fp fp_sub(const Fp& x, const Fp& y) { return fp_mod(sub(add(MODP(), x), y)); }
Fp fp_mul(const Fp& x, const Fp& y) { return fp_mod(mul(x, y)); }
Fp fp_inv(const Fp& x) {                                  // x^(p-2)
  const Big e = sub(MODP(), Big(2));
  Fp r = Big(1), b = x;
  for (int bit = 0; bit < 384; bit++) {
    if ((e.w[bit >> 5] >> (bit & 31)) & 1u) r = fp_mul(r, b);
    b = fp_mul(b, b);
  }
  return r;
}
Fp2 fp2_add(const Fp2& x, const Fp2& y) { return {{fp_add(x.c[0], y.c[0]), fp_add(x.c[1], y.c[1])}}; }
Fp2 fp2_sub(const Fp2& x, const Fp2& y) { return {{fp_sub(x.c[0], y.c[0]), fp_sub(x.c[1], y.c[1])}}; }
Fp2 fp2_mul(const Fp2& x, const Fp2& y) {
  return {{fp_sub(fp_mul(x.c[0], y.c[0]), fp_mul(x.c[1], y.c[1])), fp_add(fp_mul(x.c[0], y.c[1]), fp_mul(x.c[1], y.c[0]))}};
}
Fp2 fp2_mul_by_nonresidue(const Fp2& x) { return {{fp_sub(x.c[0], x.c[1]), fp_add(x.c[0], x.c[1])}}; }
Fp2 part(const Fp6& x, int i) { return {{x.c[2 * i], x.c[2 * i + 1]}}; }
Fp6 join(const Fp2& a, const Fp2& b, const Fp2& c) { return {{a.c[0], a.c[1], b.c[0], b.c[1], c.c[0], c.c[1]}}; }
Fp6 fp6_add(const Fp6& x, const Fp6& y) { Fp6 r; for (int i = 0; i < 6; i++) r.c[i] = fp_add(x.c[i], y.c[i]); return r; }
Fp6 fp6_sub(const Fp6& x, const Fp6& y) { Fp6 r; for (int i = 0; i < 6; i++) r.c[i] = fp_sub(x.c[i], y.c[i]); return r; }
Fp6 fp6_mul(const Fp6& x, const Fp6& y) {               // native.rs:836-861
  const Fp2 c0 = part(x, 0), c1 = part(x, 1), c2 = part(x, 2), r0 = part(y, 0), r1 = part(y, 1), r2 = part(y, 2);
  const Fp2 t0 = fp2_mul(c0, r0), t1 = fp2_mul(c1, r1), t2 = fp2_mul(c2, r2);
  const Fp2 t5 = fp2_mul(fp2_add(c1, c2), fp2_add(r1, r2));
  const Fp2 xx = fp2_add(fp2_mul_by_nonresidue(fp2_sub(fp2_sub(t5, t1), t2)), t0);
  const Fp2 t11 = fp2_mul(fp2_add(c0, c1), fp2_add(r0, r1));
  const Fp2 yy = fp2_add(fp2_sub(fp2_sub(t11, t0), t1), fp2_mul_by_nonresidue(t2));
  const Fp2 t17 = fp2_mul(fp2_add(c0, c2), fp2_add(r0, r2));
  const Fp2 zz = fp2_add(fp2_sub(fp2_sub(t17, t0), t2), t1);
  return join(xx, yy, zz);
}
Fp6 fp6_mul_by_nonresidue(const Fp6& x) {               // native.rs:863-873
  const Fp2 c0 = fp2_mul_by_nonresidue(part(x, 2));
  return {{c0.c[0], c0.c[1], x.c[0], x.c[1], x.c[2], x.c[3]}};
}
Fp6 half(const Fp12& x, int i) { Fp6 r; for (int k = 0; k < 6; k++) r.c[k] = x.c[6 * i + k]; return r; }
Fp12 fp12_mul_native(const Fp12& x, const Fp12& y) {    // native.rs:1009-1027
  const Fp6 c0 = half(x, 0), c1 = half(x, 1), r0 = half(y, 0), r1 = half(y, 1);
  const Fp6 t0 = fp6_mul(c0, r0), t1 = fp6_mul(c1, r1);
  const Fp6 xx = fp6_add(t0, fp6_mul_by_nonresidue(t1));
  const Fp6 t5 = fp6_mul(fp6_add(c0, c1), fp6_add(r0, r1));
  const Fp6 yy = fp6_sub(fp6_sub(t5, t0), t1);
  Fp12 r;
  for (int k = 0; k < 6; k++) { r.c[k] = xx.c[k]; r.c[6 + k] = yy.c[k]; }
  return r;
}


// ---- the rest of native.rs that the three larger starks need ----
const u64 BLS_X = 15132376222941642752ull;                                                    // native.rs:20-22
Fp fp_neg(const Fp& x) { return sub(MODP(), x); }                                             // native.rs:417-424: -0 = p
const Fp& HALF() { static const Fp h = fp_inv(Big(2)); return h; }
Fp2 fp2_neg(const Fp2& x) { return {{fp_neg(x.c[0]), fp_neg(x.c[1])}}; }
Fp2 fp2_mul_fp(const Fp2& x, const Fp& k) { return {{fp_mul(x.c[0], k), fp_mul(x.c[1], k)}}; }
Fp2 fp2_multiply_by_b(const Fp2& x) {
  const Fp t0 = fp_mul(x.c[0], Big(4)), t1 = fp_mul(x.c[1], Big(4));
  return {{fp_sub(t0, t1), fp_add(t0, t1)}};
}
Fp2 fp2_inv(const Fp2& x) {
  const Fp f = fp_inv(fp_add(fp_mul(x.c[0], x.c[0]), fp_mul(x.c[1], x.c[1])));
  return {{fp_mul(f, x.c[0]), fp_mul(f, fp_neg(x.c[1]))}};
}
Fp frob_fp(const u32 (*tab)[12], int i) { return Big::from_limbs(tab[i], 12); }
Fp2 frob_fp2(const u32 (*tab)[12], int i) { return {{frob_fp(tab, 2 * i), frob_fp(tab, 2 * i + 1)}}; }
Fp2 fp2_frobenius(const Fp2& x, unsigned pw) { return {{x.c[0], fp_mul(x.c[1], frob_fp(woff::FP2_FROB, pw % 2))}}; }   // native.rs:1058-1064
Fp6 fp6_neg(const Fp6& x) { Fp6 r; for (int i = 0; i < 6; i++) r.c[i] = fp_neg(x.c[i]); return r; }
Fp6 fp6_multiply_by_01(const Fp6& x, const Fp2& b0, const Fp2& b1) {
  const Fp2 c0 = part(x, 0), c1 = part(x, 1), c2 = part(x, 2);
  const Fp2 t0 = fp2_mul(c0, b0), t1 = fp2_mul(c1, b1);
  const Fp2 xx = fp2_add(fp2_mul_by_nonresidue(fp2_mul(c2, b1)), t0);
  const Fp2 t6 = fp2_mul(fp2_add(b0, b1), fp2_add(c0, c1));
  return join(xx, fp2_sub(fp2_sub(t6, t0), t1), fp2_add(fp2_mul(c2, b0), t1));
}
Fp6 fp6_multiply_by_1(const Fp6& x, const Fp2& b1) {
  return join(fp2_mul_by_nonresidue(fp2_mul(part(x, 2), b1)), fp2_mul(part(x, 0), b1), fp2_mul(part(x, 1), b1));
}
Fp6 fp6_inv(const Fp6& x) {                                                                   // native.rs:720-734
  const Fp2 c0 = part(x, 0), c1 = part(x, 1), c2 = part(x, 2);
  const Fp2 t0 = fp2_sub(fp2_mul(c0, c0), fp2_mul_by_nonresidue(fp2_mul(c2, c1)));
  const Fp2 t1 = fp2_sub(fp2_mul_by_nonresidue(fp2_mul(c2, c2)), fp2_mul(c0, c1));
  const Fp2 t2 = fp2_sub(fp2_mul(c1, c1), fp2_mul(c0, c2));
  const Fp2 t4 = fp2_inv(fp2_add(fp2_mul_by_nonresidue(fp2_add(fp2_mul(c2, t1), fp2_mul(c1, t2))), fp2_mul(c0, t0)));
  return join(fp2_mul(t4, t0), fp2_mul(t4, t1), fp2_mul(t4, t2));
}
Fp6 fp6_frobenius(const Fp6& x, unsigned pw) {                                                // native.rs:1126-1145
  return join(fp2_frobenius(part(x, 0), pw), fp2_mul(fp2_frobenius(part(x, 1), pw), frob_fp2(woff::FP6_FROB_1, pw % 6)),
              fp2_mul(fp2_frobenius(part(x, 2), pw), frob_fp2(woff::FP6_FROB_2, pw % 6)));
}
Fp12 join12(const Fp6& a, const Fp6& b) { Fp12 r; for (int k = 0; k < 6; k++) { r.c[k] = a.c[k]; r.c[6 + k] = b.c[k]; } return r; }
Fp12 fp12_one() { Fp12 r; r.c[0] = Big(1); return r; }
Fp12 fp12_inv(const Fp12& x) {                                                                // native.rs:932-940
  const Fp6 c0 = half(x, 0), c1 = half(x, 1);
  const Fp6 t = fp6_inv(fp6_sub(fp6_mul(c0, c0), fp6_mul_by_nonresidue(fp6_mul(c1, c1))));
  return join12(fp6_mul(c0, t), fp6_neg(fp6_mul(c1, t)));
}
Fp12 fp12_multiply_by_014(const Fp12& x, const Fp2& o0, const Fp2& o1, const Fp2& o4) {       // native.rs:1228-1244
  const Fp6 c0 = half(x, 0), c1 = half(x, 1);
  const Fp6 t0 = fp6_multiply_by_01(c0, o0, o1), t1 = fp6_multiply_by_1(c1, o4);
  const Fp6 xx = fp6_add(fp6_mul_by_nonresidue(t1), t0);
  const Fp6 t5 = fp6_multiply_by_01(fp6_add(c1, c0), o0, fp2_add(o1, o4));
  return join12(xx, fp6_sub(fp6_sub(t5, t0), t1));
}
Fp12 fp12_conjugate(const Fp12& x) { return join12(half(x, 0), fp6_neg(half(x, 1))); }        // native.rs:1246-1252
Fp12 fp12_frobenius(const Fp12& x, unsigned pw) {                                             // native.rs:1202-1224
  const Fp6 r0 = fp6_frobenius(half(x, 0), pw), c = fp6_frobenius(half(x, 1), pw);
  const Fp2 k = frob_fp2(woff::FP12_FROB, pw % 12);
  return join12(r0, join(fp2_mul(part(c, 0), k), fp2_mul(part(c, 1), k), fp2_mul(part(c, 2), k)));
}
Fp2 part12(const Fp12& x, int i) { return {{x.c[2 * i], x.c[2 * i + 1]}}; }
void fp4_square(const Fp2& a, const Fp2& b, Fp2& o0, Fp2& o1) {                               // native.rs:224-231
  const Fp2 a2 = fp2_mul(a, a), b2 = fp2_mul(b, b), sm = fp2_add(a, b);
  o0 = fp2_add(fp2_mul_by_nonresidue(b2), a2);
  o1 = fp2_sub(fp2_sub(fp2_mul(sm, sm), a2), b2);
}
Fp12 fp12_cyclotomic_square(const Fp12& x) {                                                  // native.rs:1254-1298
  const Fp2 c0c0 = part12(x, 0), c0c1 = part12(x, 1), c0c2 = part12(x, 2), c1c0 = part12(x, 3), c1c1 = part12(x, 4), c1c2 = part12(x, 5);
  Fp2 t00, t01, t10, t11, t20, t21;
  fp4_square(c0c0, c1c1, t00, t01); fp4_square(c1c0, c0c2, t10, t11); fp4_square(c0c1, c1c2, t20, t21);
  const Fp2 t3 = fp2_mul_by_nonresidue(t21);
  const Fp two = Big(2);
  const Fp2 c[6] = {fp2_add(fp2_mul_fp(fp2_sub(t00, c0c0), two), t00), fp2_add(fp2_mul_fp(fp2_sub(t10, c0c1), two), t10),
                    fp2_add(fp2_mul_fp(fp2_sub(t20, c0c2), two), t20), fp2_add(fp2_mul_fp(fp2_add(t3, c1c0), two), t3),
                    fp2_add(fp2_mul_fp(fp2_add(t01, c1c1), two), t01), fp2_add(fp2_mul_fp(fp2_add(t11, c1c2), two), t11)};
  Fp12 r;
  for (int i = 0; i < 6; i++) { r.c[2 * i] = c[i].c[0]; r.c[2 * i + 1] = c[i].c[1]; }
  return r;
}
// calc_precomp_stuff_loop0 / loop1 (native.rs:291-372): the named intermediates of one doubling / one addition step
struct Loop0 { Fp2 new_rx, new_ry, new_rz, t0, t1, x0, t2, t3, x1, t4, x3, x2, x4, x5, x6, x7, x8, x9, x10, x11, x12, x13; };
Loop0 calc_precomp_stuff_loop0(const Fp2& rx, const Fp2& ry, const Fp2& rz) {
  Loop0 v;
  const Fp three = Big(3), two = Big(2);
  v.t0 = fp2_mul(ry, ry); v.t1 = fp2_mul(rz, rz); v.x0 = fp2_mul_fp(v.t1, three); v.t2 = fp2_multiply_by_b(v.x0);
  v.t3 = fp2_mul_fp(v.t2, three); v.x1 = fp2_mul(ry, rz); v.t4 = fp2_mul_fp(v.x1, two); v.x2 = fp2_sub(v.t2, v.t0);
  v.x3 = fp2_mul(rx, rx); v.x4 = fp2_mul_fp(v.x3, three); v.x5 = fp2_neg(v.t4); v.x6 = fp2_sub(v.t0, v.t3);
  v.x7 = fp2_mul(rx, ry); v.x8 = fp2_mul(v.x6, v.x7); v.x9 = fp2_add(v.t0, v.t3); v.x10 = fp2_mul_fp(v.x9, HALF());
  v.x11 = fp2_mul(v.x10, v.x10); v.x12 = fp2_mul(v.t2, v.t2); v.x13 = fp2_mul_fp(v.x12, three);
  v.new_rx = fp2_mul_fp(v.x8, HALF()); v.new_ry = fp2_sub(v.x11, v.x13); v.new_rz = fp2_mul(v.t0, v.t4);
  return v;
}
struct Loop1 { Fp2 new_rx, new_ry, new_rz, t[19]; };
Loop1 calc_precomp_stuff_loop1(const Fp2& rx, const Fp2& ry, const Fp2& rz, const Fp2& qx, const Fp2& qy) {
  Loop1 w;
  Fp2* t = w.t;
  t[0] = fp2_mul(qy, rz); t[1] = fp2_sub(ry, t[0]); t[2] = fp2_mul(qx, rz); t[3] = fp2_sub(rx, t[2]); t[4] = fp2_mul(t[1], qx);
  t[5] = fp2_mul(t[3], qy); t[6] = fp2_sub(t[4], t[5]); t[7] = fp2_neg(t[1]); t[8] = fp2_mul(t[3], t[3]); t[9] = fp2_mul(t[8], t[3]);
  t[10] = fp2_mul(t[8], rx); t[11] = fp2_mul(t[1], t[1]); t[12] = fp2_mul(t[11], rz); t[13] = fp2_mul_fp(t[10], Big(2));
  t[14] = fp2_sub(t[9], t[13]); t[15] = fp2_add(t[14], t[12]); t[16] = fp2_sub(t[10], t[15]); t[17] = fp2_mul(t[16], t[1]);
  t[18] = fp2_mul(t[9], ry);
  w.new_rx = fp2_mul(t[3], t[15]); w.new_ry = fp2_sub(t[17], t[18]); w.new_rz = fp2_mul(rz, t[9]);
  return w;
}
struct Ell { Fp2 c[3]; };
std::vector<Ell> pairing_precomp_native(const Fp2& x, const Fp2& y, const Fp2& z) {             // native.rs:1358-1437
  const Fp2 zi = fp2_inv(z), qx = fp2_mul(x, zi), qy = fp2_mul(y, zi);
  Fp2 rx = qx, ry = qy, rz = {{Big(1), Big()}};
  std::vector<Ell> ell;
  for (int i = 62; i >= 0; i--) {
    const Loop0 v = calc_precomp_stuff_loop0(rx, ry, rz);
    ell.push_back({{v.x2, v.x4, v.x5}});
    rx = v.new_rx; ry = v.new_ry; rz = v.new_rz;
    if ((BLS_X >> i) & 1) {
      const Loop1 w = calc_precomp_stuff_loop1(rx, ry, rz, qx, qy);
      ell.push_back({{w.t[6], w.t[7], w.t[3]}});
      rx = w.new_rx; ry = w.new_ry; rz = w.new_rz;
    }
  }
  return ell;
}
Fp12 miller_loop_native(const Fp& px, const Fp& py, const std::vector<Ell>& pre) {            // native.rs:1440-1466
  Fp12 f = fp12_one();
  size_t j = 0;
  for (int i = 62; i >= 0; i--) {
    f = fp12_multiply_by_014(f, pre[j].c[0], fp2_mul_fp(pre[j].c[1], px), fp2_mul_fp(pre[j].c[2], py));
    if ((BLS_X >> i) & 1) {
      j++;
      f = fp12_multiply_by_014(f, pre[j].c[0], fp2_mul_fp(pre[j].c[1], px), fp2_mul_fp(pre[j].c[2], py));
    }
    if (i != 0) f = fp12_mul_native(f, f);
    j++;
  }
  return fp12_conjugate(f);
}

// ---------------------------------------------------------------------------------------------------------
// the trace: row-major [rows][cols] uint32_t
// ---------------------------------------------------------------------------------------------------------
struct Trace {
  u32* cells; size_t rows, cols;
  u32& at(size_t row, size_t col) {
    if (row >= rows || col >= cols) throw std::out_of_range("witness: cell outside the trace");
    return cells[row * cols + col];
  }
  void put(size_t row, size_t col, const u32* v, size_t n) { for (size_t i = 0; i < n; i++) at(row, col + i) = v[i]; }
  void put_rows(size_t r0, size_t r1, size_t col, const u32* v, size_t n) { for (size_t r = r0; r <= r1; r++) put(r, col, v, n); }
  void set_rows(size_t r0, size_t r1, size_t col, u32 v) { for (size_t r = r0; r <= r1; r++) at(r, col) = v; }
  // `for row in start_row..end_row+1 { fill(row) }` with identical values on every row: fill once, replicate the block
  void rep(size_t r0, size_t r1, size_t col, size_t width) {
    for (size_t r = r0 + 1; r <= r1; r++) memcpy(&at(r, col), &at(r0, col), 4 * width);
    if (width) (void)at(r1, col + width - 1);
  }
};
// get_u32_vec_from_literal / _24 (native.rs:233-240, 261-267)
void limbs(const Big& x, int n, u32* out) {
  for (int i = n; i < Big::N; i++) if (x.w[i]) throw std::overflow_error("witness: value does not fit its limbs");
  memcpy(out, x.w, 4 * n);
}
void put_big(Trace& tr, size_t row, size_t col, const Big& x, int n) { u32 l[24]; limbs(x, n, l); tr.put(row, col, l, n); }
void put_big_rows(Trace& tr, size_t r0, size_t r1, size_t col, const Big& x, int n) { u32 l[24]; limbs(x, n, l); tr.put_rows(r0, r1, col, l, n); }

// add_u32_slices / _12 (native.rs:69-100): limbs of x + y mod 2^(32 n) and the carry out of each limb
void add_carries(const Big& x, const Big& y, int n, u32* sum, u32* car) {
  u64 c = 0;
  for (int i = 0; i < n; i++) { c += (u64)x.w[i] + y.w[i]; sum[i] = (u32)c; c >>= 32; car[i] = (u32)c; }
}
// sub_u32_slices / _12 (native.rs:102-141), x >= y
void sub_borrows(const Big& x, const Big& y, int n, u32* diff, u32* bor) {
  u32 b = 0;
  for (int i = 0; i < n; i++) {
    const u64 xi = x.w[i], yi = (u64)y.w[i] + b;
    diff[i] = (u32)(xi - yi);
    b = xi >= yi ? 0 : 1;
    bor[i] = b;
  }
}

using namespace woff;

// ------------------------------------------------------------------ fp.rs
void fill_addition_trace(Trace& tr, const Big& x, const Big& y, size_t row, size_t col) {                  // fp.rs:185-201
  tr.at(row, col + fp::ADDITION_CHECK_OFFSET) = 1;
  u32 s[24], c[24];
  add_carries(x, y, 24, s, c);
  put_big(tr, row, col + fp::ADDITION_X_OFFSET, x, 24);
  put_big(tr, row, col + fp::ADDITION_Y_OFFSET, y, 24);
  tr.put(row, col + fp::ADDITION_SUM_OFFSET, s, 24);
  tr.put(row, col + fp::ADDITION_CARRY_OFFSET, c, 24);
}
void fill_trace_addition_fp(Trace& tr, const Big& x, const Big& y, size_t row, size_t col) {               // fp.rs:204-220
  tr.at(row, col + fp::FP_ADDITION_CHECK_OFFSET) = 1;
  u32 s[12], c[12];
  add_carries(x, y, 12, s, c);
  put_big(tr, row, col + fp::FP_ADDITION_X_OFFSET, x, 12);
  put_big(tr, row, col + fp::FP_ADDITION_Y_OFFSET, y, 12);
  tr.put(row, col + fp::FP_ADDITION_SUM_OFFSET, s, 12);
  tr.put(row, col + fp::FP_ADDITION_CARRY_OFFSET, c, 12);
}
void fill_subtraction_trace(Trace& tr, const Big& x, const Big& y, size_t row, size_t col) {               // fp.rs:237-253
  if (cmp(x, y) < 0) throw std::underflow_error("witness: subtraction of a larger value");
  tr.at(row, col + fp::SUBTRACTION_CHECK_OFFSET) = 1;
  u32 d[24], b[24];
  sub_borrows(x, y, 24, d, b);
  put_big(tr, row, col + fp::SUBTRACTION_X_OFFSET, x, 24);
  put_big(tr, row, col + fp::SUBTRACTION_Y_OFFSET, y, 24);
  tr.put(row, col + fp::SUBTRACTION_DIFF_OFFSET, d, 24);
  tr.put(row, col + fp::SUBTRACTION_BORROW_OFFSET, b, 24);
}
void fill_trace_subtraction_fp(Trace& tr, const Big& x, const Big& y, size_t row, size_t col) {            // fp.rs:256-272
  if (cmp(x, y) < 0) throw std::underflow_error("witness: subtraction of a larger value");
  tr.at(row, col + fp::FP_SUBTRACTION_CHECK_OFFSET) = 1;
  u32 d[12], b[12];
  sub_borrows(x, y, 12, d, b);
  put_big(tr, row, col + fp::FP_SUBTRACTION_X_OFFSET, x, 12);
  put_big(tr, row, col + fp::FP_SUBTRACTION_Y_OFFSET, y, 12);
  tr.put(row, col + fp::FP_SUBTRACTION_DIFF_OFFSET, d, 12);
  tr.put(row, col + fp::FP_SUBTRACTION_BORROW_OFFSET, b, 12);
}
void fill_trace_multiply_single_fp(Trace& tr, const Big& x, u32 y, size_t row, size_t col) {               // fp.rs:275-291
  tr.at(row, col + fp::FP_MULTIPLY_SINGLE_CHECK_OFFSET) = 1;
  u32 xl[12], res[12], car[12];
  limbs(x, 12, xl);
  u64 c = 0;
  for (int i = 0; i < 12; i++) { const u64 t = (u64)xl[i] * y + c; res[i] = (u32)t; c = t >> 32; car[i] = (u32)c; }
  if (c) throw std::overflow_error("witness: multiply_single overflows twelve limbs");
  tr.put(row, col + fp::FP_MULTIPLY_SINGLE_X_OFFSET, xl, 12);
  tr.at(row, col + fp::FP_MULTIPLY_SINGLE_Y_OFFSET) = y;
  tr.put(row, col + fp::FP_MULTIPLY_SINGLE_SUM_OFFSET, res, 12);
  tr.put(row, col + fp::FP_MULTIPLY_SINGLE_CARRY_OFFSET, car, 12);
}
Big fill_trace_reduce_single(Trace& tr, const Big& x, size_t row, size_t col) {                            // fp.rs:294-312
  Big div, rem;
  divmod(x, MODP(), div, rem);
  if (div.top() > 0) throw std::overflow_error("witness: reduce_single quotient exceeds one limb");
  fill_trace_multiply_single_fp(tr, MODP(), div.w[0], row, col + fp::FP_SINGLE_REDUCE_MULTIPLICATION_OFFSET);
  put_big(tr, row, col + fp::FP_SINGLE_REDUCE_X_OFFSET, x, 12);
  put_big(tr, row, col + fp::FP_SINGLE_REDUCED_OFFSET, rem, 12);
  fill_trace_addition_fp(tr, mul_small(MODP(), div.w[0]), rem, row, col + fp::FP_SINGLE_REDUCTION_ADDITION_OFFSET);
  return rem;
}
void fill_range_check_trace(Trace& tr, const Big& x, size_t row, size_t col) {                             // fp.rs:315-331
  u32 s[12], c[12];
  add_carries(x, RC_ADD(), 12, s, c);
  tr.at(row, col + fp::RANGE_CHECK_SELECTOR_OFFSET) = 1;
  tr.put(row, col + fp::RANGE_CHECK_SUM_OFFSET, s, 12);
  tr.put(row, col + fp::RANGE_CHECK_SUM_CARRY_OFFSET, c, 12);
  for (int i = 0; i < 32; i++) tr.at(row, col + fp::RANGE_CHECK_BIT_DECOMP_OFFSET + i) = (s[11] >> i) & 1u;
}
void fill_multiplication_trace_no_mod_reduction(Trace& tr, const Big& x, const Big& y, size_t s_row, size_t e_row, size_t col) {   // fp.rs:334-383
  tr.at(s_row, col + fp::MULTIPLICATION_FIRST_ROW_OFFSET) = 1;
  tr.set_rows(s_row, s_row + 10, col + fp::MULTIPLICATION_SELECTOR_OFFSET, 1);
  u32 xl[12], yl[12];
  limbs(x, 12, xl); limbs(y, 12, yl);
  tr.put_rows(s_row, e_row, col + fp::X_INPUT_OFFSET, xl, 12);
  tr.put_rows(s_row, e_row, col + fp::Y_INPUT_OFFSET, yl, 12);
  for (size_t r = 0; r + s_row <= e_row; r++)            // get_selector_bits_from_u32 keeps the low 12 bits (native.rs:250-259)
    for (int i = 0; i < 12; i++) tr.at(s_row + r, col + fp::SELECTOR_OFFSET + i) = (r < 32 && i == (int)r) ? 1u : 0u;
  Big prev;
  for (int i = 0; i < 12; i++) {
    u32 xy[13], car[12];                                  // multiply_by_slice (native.rs:50-66)
    u64 c = 0;
    for (int j = 0; j < 12; j++) { const u64 t = (u64)xl[j] * yl[i] + c; xy[j] = (u32)t; c = t >> 32; car[j] = (u32)c; }
    xy[12] = (u32)c;
    const size_t r = s_row + i;
    tr.put(r, col + fp::XY_OFFSET, xy, 13);
    tr.put(r, col + fp::XY_CARRIES_OFFSET, car, 12);
    const Big shifted = shl_limbs(mul_small(x, yl[i]), i);
    put_big(tr, r, col + fp::SHIFTED_XY_OFFSET, shifted, 24);
    u32 s[24], cs[24];
    add_carries(shifted, prev, 24, s, cs);
    tr.put(r, col + fp::SUM_OFFSET, s, 24);
    tr.put(r, col + fp::SUM_CARRIES_OFFSET, cs, 24);
    prev = add(shifted, prev);
    if (prev.top() >= 24) throw std::overflow_error("witness: product exceeds 768 bits");
  }
}
Big fill_reduction_trace(Trace& tr, const Big& x, size_t s_row, size_t e_row, size_t col) {                // fp.rs:386-424
  Big div, rem;
  divmod(x, MODP(), div, rem);
  fill_multiplication_trace_no_mod_reduction(tr, div, MODP(), s_row, e_row, col + fp::REDUCE_MULTIPLICATION_OFFSET);
  put_big_rows(tr, s_row, e_row, col + fp::REDUCE_X_OFFSET, x, 24);
  put_big_rows(tr, s_row, e_row, col + fp::REDUCED_OFFSET, rem, 12);
  fill_addition_trace(tr, mul(div, MODP()), rem, s_row + 11, col + fp::REDUCTION_ADDITION_OFFSET);
  return rem;
}
const size_t RED = fp::FP_SINGLE_REDUCE_TOTAL + fp::RANGE_CHECK_TOTAL;

// ------------------------------------------------------------------ fp2.rs
void put_fp2_rows(Trace& tr, size_t r0, size_t r1, size_t col, const Fp2& x) {
  put_big_rows(tr, r0, r1, col, x.c[0], 12);
  put_big_rows(tr, r0, r1, col + 12, x.c[1], 12);
}
void fill_trace_addition_fp2(Trace& tr, const Fp2& x, const Fp2& y, size_t row, size_t col) {              // fp2.rs:187-199
  fill_trace_addition_fp(tr, x.c[0], y.c[0], row, col + fp2::FP2_ADDITION_0_OFFSET);
  fill_trace_addition_fp(tr, x.c[1], y.c[1], row, col + fp2::FP2_ADDITION_1_OFFSET);
}
void fill_trace_subtraction_fp2(Trace& tr, const Fp2& x, const Fp2& y, size_t row, size_t col) {           // fp2.rs:202-214
  fill_trace_subtraction_fp(tr, x.c[0], y.c[0], row, col + fp2::FP2_SUBTRACTION_0_OFFSET);
  fill_trace_subtraction_fp(tr, x.c[1], y.c[1], row, col + fp2::FP2_SUBTRACTION_1_OFFSET);
}
void generate_trace_fp2_mul(Trace& tr, const Fp2& x, const Fp2& y, size_t s, size_t e, size_t col) {       // fp2.rs:246-321
  tr.set_rows(s, e, col + fp2::FP2_FP2_SELECTOR_OFFSET, 1);
  put_fp2_rows(tr, s, e, col + fp2::FP2_FP2_X_INPUT_OFFSET, x);
  put_fp2_rows(tr, s, e, col + fp2::FP2_FP2_Y_INPUT_OFFSET, y);
  tr.at(e, col + fp2::FP2_FP2_SELECTOR_OFFSET) = 0;
  fill_multiplication_trace_no_mod_reduction(tr, x.c[0], y.c[0], s, e, col + fp2::X_0_Y_0_MULTIPLICATION_OFFSET);
  fill_multiplication_trace_no_mod_reduction(tr, x.c[1], y.c[1], s, e, col + fp2::X_1_Y_1_MULTIPLICATION_OFFSET);
  const Big x0y0 = mul(x.c[0], y.c[0]), x1y1 = mul(x.c[1], y.c[1]);
  fill_addition_trace(tr, x0y0, MODP2(), s + 11, col + fp2::Z1_ADD_MODULUS_OFFSET);
  fill_subtraction_trace(tr, add(x0y0, MODP2()), x1y1, s + 11, col + fp2::Z1_SUBTRACTION_OFFSET);
  Big rem = fill_reduction_trace(tr, sub(add(x0y0, MODP2()), x1y1), s, e, col + fp2::Z1_REDUCE_OFFSET);
  fill_range_check_trace(tr, rem, s, col + fp2::Z1_RANGECHECK_OFFSET);
  fill_multiplication_trace_no_mod_reduction(tr, x.c[0], y.c[1], s, e, col + fp2::X_0_Y_1_MULTIPLICATION_OFFSET);
  fill_multiplication_trace_no_mod_reduction(tr, x.c[1], y.c[0], s, e, col + fp2::X_1_Y_0_MULTIPLICATION_OFFSET);
  const Big x0y1 = mul(x.c[0], y.c[1]), x1y0 = mul(x.c[1], y.c[0]);
  fill_addition_trace(tr, x0y1, x1y0, s + 11, col + fp2::Z2_ADDITION_OFFSET);
  rem = fill_reduction_trace(tr, add(x0y1, x1y0), s, e, col + fp2::Z2_REDUCE_OFFSET);
  fill_range_check_trace(tr, rem, s, col + fp2::Z2_RANGECHECK_OFFSET);
}
void fill_trace_subtraction_with_reduction(Trace& tr, const Fp2& x, const Fp2& y, size_t row, size_t col) {   // fp2.rs:346-371
  const Fp2 pp = {{MODP(), MODP()}};
  fill_trace_addition_fp2(tr, x, pp, row, col);
  const Fp2 xm = {{add(x.c[0], MODP()), add(x.c[1], MODP())}};
  fill_trace_subtraction_fp2(tr, xm, y, row, col + fp2::FP2_ADDITION_TOTAL);
  const size_t base = col + fp2::FP2_ADDITION_TOTAL + fp2::FP2_SUBTRACTION_TOTAL;
  Big rem = fill_trace_reduce_single(tr, sub(xm.c[0], y.c[0]), row, base);
  fill_range_check_trace(tr, rem, row, base + fp::FP_SINGLE_REDUCE_TOTAL);
  rem = fill_trace_reduce_single(tr, sub(xm.c[1], y.c[1]), row, base + RED);
  fill_range_check_trace(tr, rem, row, base + fp::FP_SINGLE_REDUCE_TOTAL * 2 + fp::RANGE_CHECK_TOTAL);
}
void fill_trace_addition_with_reduction(Trace& tr, const Fp2& x, const Fp2& y, size_t row, size_t col) {   // fp2.rs:413-429
  fill_trace_addition_fp2(tr, x, y, row, col);
  const size_t base = col + fp2::FP2_ADDITION_TOTAL;
  Big rem = fill_trace_reduce_single(tr, add(x.c[0], y.c[0]), row, base);
  fill_range_check_trace(tr, rem, row, base + fp::FP_SINGLE_REDUCE_TOTAL);
  rem = fill_trace_reduce_single(tr, add(x.c[1], y.c[1]), row, base + RED);
  fill_range_check_trace(tr, rem, row, base + fp::FP_SINGLE_REDUCE_TOTAL * 2 + fp::RANGE_CHECK_TOTAL);
}
void fill_trace_non_residue_multiplication(Trace& tr, const Fp2& x, size_t row, size_t col) {              // fp2.rs:432-456
  tr.at(row, col + fp2::FP2_NON_RESIDUE_MUL_CHECK_OFFSET) = 1;
  put_big(tr, row, col + fp2::FP2_NON_RESIDUE_MUL_INPUT_OFFSET, x.c[0], 12);
  put_big(tr, row, col + fp2::FP2_NON_RESIDUE_MUL_INPUT_OFFSET + 12, x.c[1], 12);
  fill_trace_addition_fp(tr, x.c[0], MODP(), row, col + fp2::FP2_NON_RESIDUE_MUL_C0_C1_SUB_OFFSET);
  fill_trace_subtraction_fp(tr, add(x.c[0], MODP()), x.c[1], row, col + fp2::FP2_NON_RESIDUE_MUL_C0_C1_SUB_OFFSET + fp::FP_ADDITION_TOTAL);
  Big rem = fill_trace_reduce_single(tr, sub(add(x.c[0], MODP()), x.c[1]), row, col + fp2::FP2_NON_RESIDUE_MUL_Z0_REDUCE_OFFSET);
  fill_range_check_trace(tr, rem, row, col + fp2::FP2_NON_RESIDUE_MUL_Z0_RANGECHECK_OFFSET);
  fill_trace_addition_fp(tr, x.c[0], x.c[1], row, col + fp2::FP2_NON_RESIDUE_MUL_C0_C1_ADD_OFFSET);
  rem = fill_trace_reduce_single(tr, add(x.c[0], x.c[1]), row, col + fp2::FP2_NON_RESIDUE_MUL_Z1_REDUCE_OFFSET);
  fill_range_check_trace(tr, rem, row, col + fp2::FP2_NON_RESIDUE_MUL_Z1_RANGECHECK_OFFSET);
}
const size_t SUB_RED_FP2 = fp2::FP2_ADDITION_TOTAL + fp2::FP2_SUBTRACTION_TOTAL + 2 * RED;
const size_t ADD_RED_FP2 = fp2::FP2_ADDITION_TOTAL + 2 * RED;
void add_red_rows(Trace& tr, const Fp2& x, const Fp2& y, size_t s, size_t e, size_t col) {
  fill_trace_addition_with_reduction(tr, x, y, s, col); tr.rep(s, e, col, ADD_RED_FP2);
}
void sub_red_rows(Trace& tr, const Fp2& x, const Fp2& y, size_t s, size_t e, size_t col) {
  fill_trace_subtraction_with_reduction(tr, x, y, s, col); tr.rep(s, e, col, SUB_RED_FP2);
}
void nonres_rows(Trace& tr, const Fp2& x, size_t s, size_t e, size_t col) {
  fill_trace_non_residue_multiplication(tr, x, s, col); tr.rep(s, e, col, fp2::FP2_NON_RESIDUE_MUL_TOTAL);
}

// ------------------------------------------------------------------ fp6.rs
void put_fp6_rows(Trace& tr, size_t r0, size_t r1, size_t col, const Fp6& x) {
  for (int i = 0; i < 6; i++) put_big_rows(tr, r0, r1, col + 12 * i, x.c[i], 12);
}
void fill_trace_addition_fp6(Trace& tr, const Fp6& x, const Fp6& y, size_t row, size_t col) {              // fp6.rs:124-132
  const u32 off[3] = {fp6::FP6_ADDITION_0_OFFSET, fp6::FP6_ADDITION_1_OFFSET, fp6::FP6_ADDITION_2_OFFSET};
  for (int i = 0; i < 3; i++) fill_trace_addition_fp2(tr, part(x, i), part(y, i), row, col + off[i]);
}
void fill_trace_subtraction_fp6(Trace& tr, const Fp6& x, const Fp6& y, size_t row, size_t col) {           // fp6.rs:175-183
  const u32 off[3] = {fp6::FP6_SUBTRACTION_0_OFFSET, fp6::FP6_SUBTRACTION_1_OFFSET, fp6::FP6_SUBTRACTION_2_OFFSET};
  for (int i = 0; i < 3; i++) fill_trace_subtraction_fp2(tr, part(x, i), part(y, i), row, col + off[i]);
}
void fill_trace_addition_with_reduction_fp6(Trace& tr, const Fp6& x, const Fp6& y, size_t row, size_t col) {   // fp6.rs:135-148
  fill_trace_addition_fp6(tr, x, y, row, col);
  for (int i = 0; i < 6; i++) {
    const size_t base = col + fp6::FP6_ADDITION_TOTAL + RED * i;
    const Big rem = fill_trace_reduce_single(tr, add(x.c[i], y.c[i]), row, base);
    fill_range_check_trace(tr, rem, row, base + fp::FP_SINGLE_REDUCE_TOTAL);
  }
}
void fill_trace_subtraction_with_reduction_fp6(Trace& tr, const Fp6& x, const Fp6& y, size_t row, size_t col) {   // fp6.rs:151-172
  Fp6 pp, xm;
  for (int i = 0; i < 6; i++) { pp.c[i] = MODP(); xm.c[i] = add(x.c[i], MODP()); }
  fill_trace_addition_fp6(tr, x, pp, row, col);
  fill_trace_subtraction_fp6(tr, xm, y, row, col + fp6::FP6_ADDITION_TOTAL);
  for (int i = 0; i < 6; i++) {
    const size_t base = col + fp6::FP6_ADDITION_TOTAL + fp6::FP6_SUBTRACTION_TOTAL + RED * i;
    const Big rem = fill_trace_reduce_single(tr, sub(xm.c[i], y.c[i]), row, base);
    fill_range_check_trace(tr, rem, row, base + fp::FP_SINGLE_REDUCE_TOTAL);
  }
}
void fill_trace_non_residue_multiplication_fp6(Trace& tr, const Fp6& x, size_t row, size_t col) {          // fp6.rs:199-210
  tr.at(row, col + fp6::FP6_NON_RESIDUE_MUL_CHECK_OFFSET) = 1;
  for (int i = 0; i < 6; i++) put_big(tr, row, col + fp6::FP6_NON_RESIDUE_MUL_INPUT_OFFSET + i * 12, x.c[i], 12);
  fill_trace_non_residue_multiplication(tr, part(x, 2), row, col + fp6::FP6_NON_RESIDUE_MUL_C2);
}
const size_t ADD_RED_FP6 = fp6::FP6_ADDITION_TOTAL + 6 * RED;
const size_t SUB_RED_FP6 = fp6::FP6_ADDITION_TOTAL + fp6::FP6_SUBTRACTION_TOTAL + 6 * RED;
void add_red6_rows(Trace& tr, const Fp6& x, const Fp6& y, size_t s, size_t e, size_t col) {
  fill_trace_addition_with_reduction_fp6(tr, x, y, s, col); tr.rep(s, e, col, ADD_RED_FP6);
}
void sub_red6_rows(Trace& tr, const Fp6& x, const Fp6& y, size_t s, size_t e, size_t col) {
  fill_trace_subtraction_with_reduction_fp6(tr, x, y, s, col); tr.rep(s, e, col, SUB_RED_FP6);
}
void nonres6_rows(Trace& tr, const Fp6& x, size_t s, size_t e, size_t col) {
  fill_trace_non_residue_multiplication_fp6(tr, x, s, col); tr.rep(s, e, col, fp6::FP6_NON_RESIDUE_MUL_TOTAL);
}
void fill_trace_fp6_multiplication(Trace& tr, const Fp6& x, const Fp6& y, size_t s, size_t e, size_t col) {   // fp6.rs:213-303
  put_fp6_rows(tr, s, e, col + fp6::FP6_MUL_X_INPUT_OFFSET, x);
  put_fp6_rows(tr, s, e, col + fp6::FP6_MUL_Y_INPUT_OFFSET, y);
  tr.set_rows(s, e, col + fp6::FP6_MUL_SELECTOR_OFFSET, 1);
  tr.at(e, col + fp6::FP6_MUL_SELECTOR_OFFSET) = 0;
  const Fp2 c0 = part(x, 0), c1 = part(x, 1), c2 = part(x, 2), r0 = part(y, 0), r1 = part(y, 1), r2 = part(y, 2);
  const Fp2 t0 = fp2_mul(c0, r0); generate_trace_fp2_mul(tr, c0, r0, s, e, col + fp6::FP6_MUL_T0_CALC_OFFSET);
  const Fp2 t1 = fp2_mul(c1, r1); generate_trace_fp2_mul(tr, c1, r1, s, e, col + fp6::FP6_MUL_T1_CALC_OFFSET);
  const Fp2 t2 = fp2_mul(c2, r2); generate_trace_fp2_mul(tr, c2, r2, s, e, col + fp6::FP6_MUL_T2_CALC_OFFSET);
  const Fp2 t3 = fp2_add(c1, c2); add_red_rows(tr, c1, c2, s, e, col + fp6::FP6_MUL_T3_CALC_OFFSET);
  const Fp2 t4 = fp2_add(r1, r2); add_red_rows(tr, r1, r2, s, e, col + fp6::FP6_MUL_T4_CALC_OFFSET);
  const Fp2 t5 = fp2_mul(t3, t4); generate_trace_fp2_mul(tr, t3, t4, s, e, col + fp6::FP6_MUL_T5_CALC_OFFSET);
  const Fp2 t6 = fp2_sub(t5, t1); sub_red_rows(tr, t5, t1, s, e, col + fp6::FP6_MUL_T6_CALC_OFFSET);
  const Fp2 t7 = fp2_sub(t6, t2); sub_red_rows(tr, t6, t2, s, e, col + fp6::FP6_MUL_T7_CALC_OFFSET);
  const Fp2 t8 = fp2_mul_by_nonresidue(t7); nonres_rows(tr, t7, s, e, col + fp6::FP6_MUL_T8_CALC_OFFSET);
  add_red_rows(tr, t8, t0, s, e, col + fp6::FP6_MUL_X_CALC_OFFSET);
  const Fp2 t9 = fp2_add(c0, c1); add_red_rows(tr, c0, c1, s, e, col + fp6::FP6_MUL_T9_CALC_OFFSET);
  const Fp2 t10 = fp2_add(r0, r1); add_red_rows(tr, r0, r1, s, e, col + fp6::FP6_MUL_T10_CALC_OFFSET);
  const Fp2 t11 = fp2_mul(t9, t10); generate_trace_fp2_mul(tr, t9, t10, s, e, col + fp6::FP6_MUL_T11_CALC_OFFSET);
  const Fp2 t12 = fp2_sub(t11, t0); sub_red_rows(tr, t11, t0, s, e, col + fp6::FP6_MUL_T12_CALC_OFFSET);
  const Fp2 t13 = fp2_sub(t12, t1); sub_red_rows(tr, t12, t1, s, e, col + fp6::FP6_MUL_T13_CALC_OFFSET);
  const Fp2 t14 = fp2_mul_by_nonresidue(t2); nonres_rows(tr, t2, s, e, col + fp6::FP6_MUL_T14_CALC_OFFSET);
  add_red_rows(tr, t13, t14, s, e, col + fp6::FP6_MUL_Y_CALC_OFFSET);
  const Fp2 t15 = fp2_add(c0, c2); add_red_rows(tr, c0, c2, s, e, col + fp6::FP6_MUL_T15_CALC_OFFSET);
  const Fp2 t16 = fp2_add(r0, r2); add_red_rows(tr, r0, r2, s, e, col + fp6::FP6_MUL_T16_CALC_OFFSET);
  const Fp2 t17 = fp2_mul(t15, t16); generate_trace_fp2_mul(tr, t15, t16, s, e, col + fp6::FP6_MUL_T17_CALC_OFFSET);
  const Fp2 t18 = fp2_sub(t17, t0); sub_red_rows(tr, t17, t0, s, e, col + fp6::FP6_MUL_T18_CALC_OFFSET);
  const Fp2 t19 = fp2_sub(t18, t2); sub_red_rows(tr, t18, t2, s, e, col + fp6::FP6_MUL_T19_CALC_OFFSET);
  add_red_rows(tr, t19, t1, s, e, col + fp6::FP6_MUL_Z_CALC_OFFSET);
}

// ------------------------------------------------------------------ fp12.rs
void fill_trace_fp12_multiplication(Trace& tr, const Fp12& x, const Fp12& y, size_t s, size_t e, size_t col) {   // fp12.rs:186-232
  for (int i = 0; i < 12; i++) {
    put_big_rows(tr, s, e, col + fp12::FP12_MUL_X_INPUT_OFFSET + 12 * i, x.c[i], 12);
    put_big_rows(tr, s, e, col + fp12::FP12_MUL_Y_INPUT_OFFSET + 12 * i, y.c[i], 12);
  }
  tr.set_rows(s, e, col + fp12::FP12_MUL_SELECTOR_OFFSET, 1);
  tr.at(e, col + fp12::FP12_MUL_SELECTOR_OFFSET) = 0;
  const Fp6 c0 = half(x, 0), c1 = half(x, 1), r0 = half(y, 0), r1 = half(y, 1);
  const Fp6 t0 = fp6_mul(c0, r0); fill_trace_fp6_multiplication(tr, c0, r0, s, e, col + fp12::FP12_MUL_T0_CALC_OFFSET);
  const Fp6 t1 = fp6_mul(c1, r1); fill_trace_fp6_multiplication(tr, c1, r1, s, e, col + fp12::FP12_MUL_T1_CALC_OFFSET);
  const Fp6 t2 = fp6_mul_by_nonresidue(t1); nonres6_rows(tr, t1, s, e, col + fp12::FP12_MUL_T2_CALC_OFFSET);
  add_red6_rows(tr, t0, t2, s, e, col + fp12::FP12_MUL_X_CALC_OFFSET);
  const Fp6 t3 = fp6_add(c0, c1); add_red6_rows(tr, c0, c1, s, e, col + fp12::FP12_MUL_T3_CALC_OFFSET);
  const Fp6 t4 = fp6_add(r0, r1); add_red6_rows(tr, r0, r1, s, e, col + fp12::FP12_MUL_T4_CALC_OFFSET);
  const Fp6 t5 = fp6_mul(t3, t4); fill_trace_fp6_multiplication(tr, t3, t4, s, e, col + fp12::FP12_MUL_T5_CALC_OFFSET);
  const Fp6 t6 = fp6_sub(t5, t0); sub_red6_rows(tr, t5, t0, s, e, col + fp12::FP12_MUL_T6_CALC_OFFSET);
  sub_red6_rows(tr, t6, t1, s, e, col + fp12::FP12_MUL_Y_CALC_OFFSET);
}


// ------------------------------------------------------------------ the gadgets of the three larger starks
void fill_trace_negate_fp2(Trace& tr, const Fp2& x, size_t row, size_t col) { fill_trace_addition_fp2(tr, x, fp2_neg(x), row, col); }   // fp2.rs:232-243
void negate_rows(Trace& tr, const Fp2& x, size_t s, size_t e, size_t col) {
  fill_trace_negate_fp2(tr, x, s, col); tr.rep(s, e, col, fp2::FP2_ADDITION_TOTAL);
}
void fill_trace_fp2_fp_mul(Trace& tr, const Fp2& x, const Fp& y, size_t s, size_t e, size_t col) {         // fp2.rs:324-343
  tr.set_rows(s, e, col + fp2::FP2_FP_MUL_SELECTOR_OFFSET, 1);
  put_fp2_rows(tr, s, e, col + fp2::FP2_FP_X_INPUT_OFFSET, x);
  put_big_rows(tr, s, e, col + fp2::FP2_FP_Y_INPUT_OFFSET, y, 12);
  tr.at(e, col + fp2::FP2_FP_MUL_SELECTOR_OFFSET) = 0;
  fill_multiplication_trace_no_mod_reduction(tr, x.c[0], y, s, e, col + fp2::X0_Y_MULTIPLICATION_OFFSET);
  Big rem = fill_reduction_trace(tr, mul(x.c[0], y), s, e, col + fp2::X0_Y_REDUCE_OFFSET);
  fill_range_check_trace(tr, rem, s, col + fp2::X0_Y_RANGECHECK_OFFSET);
  fill_multiplication_trace_no_mod_reduction(tr, x.c[1], y, s, e, col + fp2::X1_Y_MULTIPLICATION_OFFSET);
  rem = fill_reduction_trace(tr, mul(x.c[1], y), s, e, col + fp2::X1_Y_REDUCE_OFFSET);
  fill_range_check_trace(tr, rem, s, col + fp2::X1_Y_RANGECHECK_OFFSET);
}
void fill_multiply_by_b_trace(Trace& tr, const Fp2& x, size_t s, size_t e, size_t col) {                   // fp2.rs:374-410
  tr.set_rows(s, e, col + fp2::MULTIPLY_B_SELECTOR_OFFSET, 1);
  put_fp2_rows(tr, s, e, col + fp2::MULTIPLY_B_X_OFFSET, x);
  tr.at(e, col + fp2::MULTIPLY_B_SELECTOR_OFFSET) = 0;
  const Big four(4);
  fill_multiplication_trace_no_mod_reduction(tr, x.c[0], four, s, e, col + fp2::MULTIPLY_B_X0_B_MUL_OFFSET);
  fill_multiplication_trace_no_mod_reduction(tr, x.c[1], four, s, e, col + fp2::MULTIPLY_B_X1_B_MUL_OFFSET);
  const Big x0y = mul(x.c[0], four), x1y = mul(x.c[1], four);
  fill_addition_trace(tr, x0y, MODP2(), s + 11, col + fp2::MULTIPLY_B_ADD_MODSQ_OFFSET);
  fill_subtraction_trace(tr, add(x0y, MODP2()), x1y, s + 11, col + fp2::MULTIPLY_B_SUB_OFFSET);
  Big rem = fill_reduction_trace(tr, sub(add(x0y, MODP2()), x1y), s, e, col + fp2::MULTIPLY_B_Z0_REDUCE_OFFSET);
  fill_range_check_trace(tr, rem, s, col + fp2::MULTIPLY_B_Z0_RANGECHECK_OFFSET);
  fill_addition_trace(tr, x0y, x1y, s + 11, col + fp2::MULTIPLY_B_ADD_OFFSET);
  rem = fill_reduction_trace(tr, add(x0y, x1y), s, e, col + fp2::MULTIPLY_B_Z1_REDUCE_OFFSET);
  fill_range_check_trace(tr, rem, s, col + fp2::MULTIPLY_B_Z1_RANGECHECK_OFFSET);
}
void fill_trace_fp4_sq(Trace& tr, const Fp2& x, const Fp2& y, size_t s, size_t e, size_t col) {            // fp2.rs:459-502
  put_fp2_rows(tr, s, e, col + fp2::FP4_SQ_INPUT_X_OFFSET, x);
  put_fp2_rows(tr, s, e, col + fp2::FP4_SQ_INPUT_Y_OFFSET, y);
  tr.set_rows(s, e, col + fp2::FP4_SQ_SELECTOR_OFFSET, 1);
  tr.at(e, col + fp2::FP4_SQ_SELECTOR_OFFSET) = 0;
  const Fp2 t0 = fp2_mul(x, x); generate_trace_fp2_mul(tr, x, x, s, e, col + fp2::FP4_SQ_T0_CALC_OFFSET);
  const Fp2 t1 = fp2_mul(y, y); generate_trace_fp2_mul(tr, y, y, s, e, col + fp2::FP4_SQ_T1_CALC_OFFSET);
  const Fp2 t2 = fp2_mul_by_nonresidue(t1); nonres_rows(tr, t1, s, e, col + fp2::FP4_SQ_T2_CALC_OFFSET);
  add_red_rows(tr, t2, t0, s, e, col + fp2::FP4_SQ_X_CALC_OFFSET);
  const Fp2 t3 = fp2_add(x, y); add_red_rows(tr, x, y, s, e, col + fp2::FP4_SQ_T3_CALC_OFFSET);
  const Fp2 t4 = fp2_mul(t3, t3); generate_trace_fp2_mul(tr, t3, t3, s, e, col + fp2::FP4_SQ_T4_CALC_OFFSET);
  const Fp2 t5 = fp2_sub(t4, t0); sub_red_rows(tr, t4, t0, s, e, col + fp2::FP4_SQ_T5_CALC_OFFSET);
  sub_red_rows(tr, t5, t1, s, e, col + fp2::FP4_SQ_Y_CALC_OFFSET);
}
void fill_trace_fp2_forbenius_map(Trace& tr, const Fp2& x, unsigned pw, size_t s, size_t e, size_t col) {  // fp2.rs:505-531
  const unsigned div = pw / 2, rem = pw % 2;
  put_fp2_rows(tr, s, e, col + fp2::FP2_FORBENIUS_MAP_INPUT_OFFSET, x);
  tr.set_rows(s, e, col + fp2::FP2_FORBENIUS_MAP_SELECTOR_OFFSET, 1);
  tr.set_rows(s, e, col + fp2::FP2_FORBENIUS_MAP_POW_OFFSET, pw);
  tr.set_rows(s, e, col + fp2::FP2_FORBENIUS_MAP_DIV_OFFSET, div);
  tr.set_rows(s, e, col + fp2::FP2_FORBENIUS_MAP_REM_OFFSET, rem);
  tr.at(e, col + fp2::FP2_FORBENIUS_MAP_SELECTOR_OFFSET) = 0;
  const Fp k = frob_fp(FP2_FROB, rem);
  fill_multiplication_trace_no_mod_reduction(tr, x.c[1], k, s, e, col + fp2::FP2_FORBENIUS_MAP_T0_CALC_OFFSET);
  tr.at(s + 11, col + fp2::FP2_FORBENIUS_MAP_MUL_RES_ROW) = 1;
  const size_t base = col + fp2::FP2_FORBENIUS_MAP_T0_CALC_OFFSET + fp::FP_MULTIPLICATION_TOTAL_COLUMNS;
  const Big res = fill_reduction_trace(tr, mul(x.c[1], k), s, e, base);
  fill_range_check_trace(tr, res, s, base + fp::REDUCTION_TOTAL);
  tr.rep(s, e, base + fp::REDUCTION_TOTAL, fp::RANGE_CHECK_TOTAL);
}
void fill_trace_negate_fp6(Trace& tr, const Fp6& x, size_t row, size_t col) { fill_trace_addition_fp6(tr, x, fp6_neg(x), row, col); }   // fp6.rs:186-196
void negate6_rows(Trace& tr, const Fp6& x, size_t s, size_t e, size_t col) {
  fill_trace_negate_fp6(tr, x, s, col); tr.rep(s, e, col, fp6::FP6_ADDITION_TOTAL);
}
void fill_trace_multiply_by_1(Trace& tr, const Fp6& x, const Fp2& b1, size_t s, size_t e, size_t col) {    // fp6.rs:306-333
  put_fp6_rows(tr, s, e, col + fp6::MULTIPLY_BY_1_INPUT_OFFSET, x);
  put_fp2_rows(tr, s, e, col + fp6::MULTIPLY_BY_1_B1_OFFSET, b1);
  tr.set_rows(s, e, col + fp6::MULTIPLY_BY_1_SELECTOR_OFFSET, 1);
  tr.at(e, col + fp6::MULTIPLY_BY_1_SELECTOR_OFFSET) = 0;
  const Fp2 c0 = part(x, 0), c1 = part(x, 1), c2 = part(x, 2);
  const Fp2 t0 = fp2_mul(c2, b1); generate_trace_fp2_mul(tr, c2, b1, s, e, col + fp6::MULTIPLY_BY_1_T0_CALC_OFFSET);
  nonres_rows(tr, t0, s, e, col + fp6::MULTIPLY_BY_1_X_CALC_OFFSET);
  generate_trace_fp2_mul(tr, c0, b1, s, e, col + fp6::MULTIPLY_BY_1_Y_CALC_OFFSET);
  generate_trace_fp2_mul(tr, c1, b1, s, e, col + fp6::MULTIPLY_BY_1_Z_CALC_OFFSET);
}
void fill_trace_multiply_by_01(Trace& tr, const Fp6& x, const Fp2& b0, const Fp2& b1, size_t s, size_t e, size_t col) {   // fp6.rs:336-394
  put_fp6_rows(tr, s, e, col + fp6::MULTIPLY_BY_01_INPUT_OFFSET, x);
  put_fp2_rows(tr, s, e, col + fp6::MULTIPLY_BY_01_B0_OFFSET, b0);
  put_fp2_rows(tr, s, e, col + fp6::MULTIPLY_BY_01_B1_OFFSET, b1);
  tr.set_rows(s, e, col + fp6::MULTIPLY_BY_01_SELECTOR_OFFSET, 1);
  tr.at(e, col + fp6::MULTIPLY_BY_01_SELECTOR_OFFSET) = 0;
  const Fp2 c0 = part(x, 0), c1 = part(x, 1), c2 = part(x, 2);
  const Fp2 t0 = fp2_mul(c0, b0); generate_trace_fp2_mul(tr, c0, b0, s, e, col + fp6::MULTIPLY_BY_01_T0_CALC_OFFSET);
  const Fp2 t1 = fp2_mul(c1, b1); generate_trace_fp2_mul(tr, c1, b1, s, e, col + fp6::MULTIPLY_BY_01_T1_CALC_OFFSET);
  const Fp2 t2 = fp2_mul(c2, b1); generate_trace_fp2_mul(tr, c2, b1, s, e, col + fp6::MULTIPLY_BY_01_T2_CALC_OFFSET);
  const Fp2 t3 = fp2_mul_by_nonresidue(t2); nonres_rows(tr, t2, s, e, col + fp6::MULTIPLY_BY_01_T3_CALC_OFFSET);
  add_red_rows(tr, t3, t0, s, e, col + fp6::MULTIPLY_BY_01_X_CALC_OFFSET);
  const Fp2 t4 = fp2_add(b0, b1); add_red_rows(tr, b0, b1, s, e, col + fp6::MULTIPLY_BY_01_T4_CALC_OFFSET);
  const Fp2 t5 = fp2_add(c0, c1); add_red_rows(tr, c0, c1, s, e, col + fp6::MULTIPLY_BY_01_T5_CALC_OFFSET);
  const Fp2 t6 = fp2_mul(t4, t5); generate_trace_fp2_mul(tr, t4, t5, s, e, col + fp6::MULTIPLY_BY_01_T6_CALC_OFFSET);
  const Fp2 t7 = fp2_sub(t6, t0); sub_red_rows(tr, t6, t0, s, e, col + fp6::MULTIPLY_BY_01_T7_CALC_OFFSET);
  sub_red_rows(tr, t7, t1, s, e, col + fp6::MULTIPLY_BY_01_Y_CALC_OFFSET);
  const Fp2 t8 = fp2_mul(c2, b0); generate_trace_fp2_mul(tr, c2, b0, s, e, col + fp6::MULTIPLY_BY_01_T8_CALC_OFFSET);
  add_red_rows(tr, t8, t1, s, e, col + fp6::MULTIPLY_BY_01_Z_CALC_OFFSET);
}
void fill_trace_fp6_forbenius_map(Trace& tr, const Fp6& x, unsigned pw, size_t s, size_t e, size_t col) {  // fp6.rs:397-431
  const unsigned div = pw / 6, rem = pw % 6;
  put_fp6_rows(tr, s, e, col + fp6::FP6_FORBENIUS_MAP_INPUT_OFFSET, x);
  tr.set_rows(s, e, col + fp6::FP6_FORBENIUS_MAP_SELECTOR_OFFSET, 1);
  tr.set_rows(s, e, col + fp6::FP6_FORBENIUS_MAP_POW_OFFSET, pw);
  tr.set_rows(s, e, col + fp6::FP6_FORBENIUS_MAP_DIV_OFFSET, div);
  tr.set_rows(s, e, col + fp6::FP6_FORBENIUS_MAP_REM_OFFSET, rem);
  tr.set_rows(s, e, col + fp6::FP6_FORBENIUS_MAP_BIT0_OFFSET, rem & 1);
  tr.set_rows(s, e, col + fp6::FP6_FORBENIUS_MAP_BIT1_OFFSET, (rem >> 1) & 1);
  tr.set_rows(s, e, col + fp6::FP6_FORBENIUS_MAP_BIT2_OFFSET, rem >> 2);
  tr.at(e, col + fp6::FP6_FORBENIUS_MAP_SELECTOR_OFFSET) = 0;
  const Fp2 c0 = part(x, 0), c1 = part(x, 1), c2 = part(x, 2);
  fill_trace_fp2_forbenius_map(tr, c0, pw, s, e, col + fp6::FP6_FORBENIUS_MAP_X_CALC_OFFSET);
  const Fp2 t0 = fp2_frobenius(c1, pw);
  fill_trace_fp2_forbenius_map(tr, c1, pw, s, e, col + fp6::FP6_FORBENIUS_MAP_T0_CALC_OFFSET);
  generate_trace_fp2_mul(tr, t0, frob_fp2(FP6_FROB_1, rem), s, e, col + fp6::FP6_FORBENIUS_MAP_Y_CALC_OFFSET);
  const Fp2 t1 = fp2_frobenius(c2, pw);
  fill_trace_fp2_forbenius_map(tr, c2, pw, s, e, col + fp6::FP6_FORBENIUS_MAP_T1_CALC_OFFSET);
  generate_trace_fp2_mul(tr, t1, frob_fp2(FP6_FROB_2, rem), s, e, col + fp6::FP6_FORBENIUS_MAP_Z_CALC_OFFSET);
}
void put_fp12_rows(Trace& tr, size_t r0, size_t r1, size_t col, const Fp12& x) {
  for (int i = 0; i < 12; i++) put_big_rows(tr, r0, r1, col + 12 * i, x.c[i], 12);
}
void fill_trace_multiply_by_014(Trace& tr, const Fp12& x, const Fp2& o0, const Fp2& o1, const Fp2& o4, size_t s, size_t e, size_t col) {   // fp12.rs:132-183
  put_fp12_rows(tr, s, e, col + fp12::MULTIPLY_BY_014_INPUT_OFFSET, x);
  put_fp2_rows(tr, s, e, col + fp12::MULTIPLY_BY_014_O0_OFFSET, o0);
  put_fp2_rows(tr, s, e, col + fp12::MULTIPLY_BY_014_O1_OFFSET, o1);
  put_fp2_rows(tr, s, e, col + fp12::MULTIPLY_BY_014_O4_OFFSET, o4);
  tr.set_rows(s, e, col + fp12::MULTIPLY_BY_014_SELECTOR_OFFSET, 1);
  tr.at(e, col + fp12::MULTIPLY_BY_014_SELECTOR_OFFSET) = 0;
  const Fp6 c0 = half(x, 0), c1 = half(x, 1);
  const Fp6 t0 = fp6_multiply_by_01(c0, o0, o1); fill_trace_multiply_by_01(tr, c0, o0, o1, s, e, col + fp12::MULTIPLY_BY_014_T0_CALC_OFFSET);
  const Fp6 t1 = fp6_multiply_by_1(c1, o4); fill_trace_multiply_by_1(tr, c1, o4, s, e, col + fp12::MULTIPLY_BY_014_T1_CALC_OFFSET);
  const Fp6 t2 = fp6_mul_by_nonresidue(t1); nonres6_rows(tr, t1, s, e, col + fp12::MULTIPLY_BY_014_T2_CALC_OFFSET);
  add_red6_rows(tr, t2, t0, s, e, col + fp12::MULTIPLY_BY_014_X_CALC_OFFSET);
  const Fp6 t3 = fp6_add(c0, c1); add_red6_rows(tr, c0, c1, s, e, col + fp12::MULTIPLY_BY_014_T3_CALC_OFFSET);
  const Fp2 t4 = fp2_add(o1, o4); add_red_rows(tr, o1, o4, s, e, col + fp12::MULTIPLY_BY_014_T4_CALC_OFFSET);
  const Fp6 t5 = fp6_multiply_by_01(t3, o0, t4); fill_trace_multiply_by_01(tr, t3, o0, t4, s, e, col + fp12::MULTIPLY_BY_014_T5_CALC_OFFSET);
  const Fp6 t6 = fp6_sub(t5, t0); sub_red6_rows(tr, t5, t0, s, e, col + fp12::MULTIPLY_BY_014_T6_CALC_OFFSET);
  sub_red6_rows(tr, t6, t1, s, e, col + fp12::MULTIPLY_BY_014_Y_CALC_OFFSET);
}
void fill_trace_cyclotomic_sq(Trace& tr, const Fp12& x, size_t s, size_t e, size_t col) {                  // fp12.rs:234-332
  put_fp12_rows(tr, s, e, col + fp12::CYCLOTOMIC_SQ_INPUT_OFFSET, x);
  tr.set_rows(s, e, col + fp12::CYCLOTOMIC_SQ_SELECTOR_OFFSET, 1);
  tr.at(e, col + fp12::CYCLOTOMIC_SQ_SELECTOR_OFFSET) = 0;
  const Fp2 c0c0 = part12(x, 0), c0c1 = part12(x, 1), c0c2 = part12(x, 2), c1c0 = part12(x, 3), c1c1 = part12(x, 4), c1c2 = part12(x, 5);
  Fp2 t00, t01, t10, t11, t20, t21;
  fp4_square(c0c0, c1c1, t00, t01); fill_trace_fp4_sq(tr, c0c0, c1c1, s, e, col + fp12::CYCLOTOMIC_SQ_T0_CALC_OFFSET);
  fp4_square(c1c0, c0c2, t10, t11); fill_trace_fp4_sq(tr, c1c0, c0c2, s, e, col + fp12::CYCLOTOMIC_SQ_T1_CALC_OFFSET);
  fp4_square(c0c1, c1c2, t20, t21); fill_trace_fp4_sq(tr, c0c1, c1c2, s, e, col + fp12::CYCLOTOMIC_SQ_T2_CALC_OFFSET);
  const Fp2 t3 = fp2_mul_by_nonresidue(t21); nonres_rows(tr, t21, s, e, col + fp12::CYCLOTOMIC_SQ_T3_CALC_OFFSET);
  const Fp two = Big(2);
  auto branch = [&](bool subtract, const Fp2& a, const Fp2& b, u32 o_t, u32 o_2, u32 o_c) {      // t = a -+ b ; t' = 2 t ; c = t' + a
    Fp2 t;
    if (subtract) { t = fp2_sub(a, b); sub_red_rows(tr, a, b, s, e, col + o_t); }
    else { t = fp2_add(a, b); add_red_rows(tr, a, b, s, e, col + o_t); }
    const Fp2 t_2 = fp2_mul_fp(t, two);
    fill_trace_fp2_fp_mul(tr, t, two, s, e, col + o_2);
    add_red_rows(tr, t_2, a, s, e, col + o_c);
  };
  branch(true, t00, c0c0, fp12::CYCLOTOMIC_SQ_T4_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_T5_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_C0_CALC_OFFSET);
  branch(true, t10, c0c1, fp12::CYCLOTOMIC_SQ_T6_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_T7_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_C1_CALC_OFFSET);
  branch(true, t20, c0c2, fp12::CYCLOTOMIC_SQ_T8_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_T9_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_C2_CALC_OFFSET);
  branch(false, t3, c1c0, fp12::CYCLOTOMIC_SQ_T10_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_T11_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_C3_CALC_OFFSET);
  branch(false, t01, c1c1, fp12::CYCLOTOMIC_SQ_T12_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_T13_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_C4_CALC_OFFSET);
  branch(false, t11, c1c2, fp12::CYCLOTOMIC_SQ_T14_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_T15_CALC_OFFSET, fp12::CYCLOTOMIC_SQ_C5_CALC_OFFSET);
}
Fp12 fill_trace_cyclotomic_exp(Trace& tr, const Fp12& x, size_t s, size_t e, size_t col) {                 // fp12.rs:335-375
  if (e + 1 - s != 70 * 12 + 1) throw std::logic_error("witness: cyclotomic_exp spans 841 rows");
  put_fp12_rows(tr, s, e, col + fp12::INPUT_OFFSET, x);
  tr.set_rows(s, e, col + fp12::CYCLOTOMIC_EXP_SELECTOR_OFFSET, 1);
  tr.at(e, col + fp12::CYCLOTOMIC_EXP_SELECTOR_OFFSET) = 0;
  tr.at(s, col + fp12::CYCLOTOMIC_EXP_START_ROW) = 1;
  Fp12 z = fp12_one();
  int i = 63;
  bool bitone = false;
  for (int j = 0; j < 70; j++) {
    const size_t s_row = s + 12 * j, e_row = s_row + 11;
    if (bitone) tr.set_rows(s_row, e_row, col + fp12::BIT1_SELECTOR_OFFSET, 1);
    put_fp12_rows(tr, s_row, e_row, col + fp12::Z_OFFSET, z);
    tr.at(s_row, col + fp12::FIRST_ROW_SELECTOR_OFFSET) = 1;
    if (bitone) {
      fill_trace_fp12_multiplication(tr, z, x, s_row, e_row, col + fp12::Z_MUL_INPUT_OFFSET);
      z = fp12_mul_native(z, x);
    } else {
      fill_trace_cyclotomic_sq(tr, z, s_row, e_row, col + fp12::Z_CYCLOTOMIC_SQ_OFFSET);
      z = fp12_cyclotomic_square(z);
    }
    if (((BLS_X >> i) & 1) && !bitone) bitone = true;
    else if (j < 69) { i--; bitone = false; }
  }
  tr.at(s + 70 * 12, col + fp12::RES_ROW_SELECTOR_OFFSET) = 1;
  put_fp12_rows(tr, s + 70 * 12, s + 70 * 12, col + fp12::Z_OFFSET, z);
  return z;
}
void fill_trace_fp12_forbenius_map(Trace& tr, const Fp12& x, unsigned pw, size_t s, size_t e, size_t col) {   // fp12.rs:378-411
  const unsigned div = pw / 12, rem = pw % 12;
  put_fp12_rows(tr, s, e, col + fp12::FP12_FORBENIUS_MAP_INPUT_OFFSET, x);
  tr.set_rows(s, e, col + fp12::FP12_FORBENIUS_MAP_SELECTOR_OFFSET, 1);
  tr.set_rows(s, e, col + fp12::FP12_FORBENIUS_MAP_POW_OFFSET, pw);
  tr.set_rows(s, e, col + fp12::FP12_FORBENIUS_MAP_DIV_OFFSET, div);
  tr.set_rows(s, e, col + fp12::FP12_FORBENIUS_MAP_REM_OFFSET, rem);
  tr.set_rows(s, e, col + fp12::FP12_FORBENIUS_MAP_BIT0_OFFSET, rem & 1);
  tr.set_rows(s, e, col + fp12::FP12_FORBENIUS_MAP_BIT1_OFFSET, (rem >> 1) & 1);
  tr.set_rows(s, e, col + fp12::FP12_FORBENIUS_MAP_BIT2_OFFSET, (rem >> 2) & 1);
  tr.set_rows(s, e, col + fp12::FP12_FORBENIUS_MAP_BIT3_OFFSET, rem >> 3);
  tr.at(e, col + fp12::FP12_FORBENIUS_MAP_SELECTOR_OFFSET) = 0;
  const Fp6 r0 = half(x, 0), r1 = half(x, 1);
  fill_trace_fp6_forbenius_map(tr, r0, pw, s, e, col + fp12::FP12_FORBENIUS_MAP_R0_CALC_OFFSET);
  const Fp6 c = fp6_frobenius(r1, pw);
  fill_trace_fp6_forbenius_map(tr, r1, pw, s, e, col + fp12::FP12_FORBENIUS_MAP_C0C1C2_CALC_OFFSET);
  const Fp2 k = frob_fp2(FP12_FROB, rem);
  generate_trace_fp2_mul(tr, part(c, 0), k, s, e, col + fp12::FP12_FORBENIUS_MAP_C0_CALC_OFFSET);
  generate_trace_fp2_mul(tr, part(c, 1), k, s, e, col + fp12::FP12_FORBENIUS_MAP_C1_CALC_OFFSET);
  generate_trace_fp2_mul(tr, part(c, 2), k, s, e, col + fp12::FP12_FORBENIUS_MAP_C2_CALC_OFFSET);
}
Fp12 fill_trace_fp12_conjugate(Trace& tr, const Fp12& x, size_t row, size_t col) {                         // fp12.rs:414-426
  put_fp12_rows(tr, row, row, col + fp12::FP12_CONJUGATE_INPUT_OFFSET, x);
  const Fp12 conj = fp12_conjugate(x);
  put_fp12_rows(tr, row, row, col + fp12::FP12_CONJUGATE_OUTPUT_OFFSET, conj);
  fill_trace_addition_fp6(tr, half(x, 1), half(conj, 1), row, col + fp12::FP12_CONJUGATE_ADDITIION_OFFSET);
  return conj;
}
// miller_loop.rs:87-146
// run `work(i)` for i in [0, n) on up to 8 host threads (the blocks of a trace are disjoint row ranges); the first
// exception is rethrown on the calling thread
template <class F> void parallel_blocks(size_t n, F&& work) {
  const unsigned hw = std::thread::hardware_concurrency();
  const size_t workers = std::min<size_t>(n, std::max(1u, std::min(8u, hw ? hw : 1u)));
  if (workers <= 1) { for (size_t i = 0; i < n; i++) work(i); return; }
  std::atomic<size_t> next{0};
  std::mutex err_mu;
  std::string err;
  std::vector<std::thread> th;
  for (size_t w = 0; w < workers; w++)
    th.emplace_back([&] {
      for (size_t i; (i = next.fetch_add(1)) < n;) {
        try { work(i); }
        catch (const std::exception& ex) { std::lock_guard<std::mutex> g(err_mu); if (err.empty()) err = ex.what(); }
      }
    });
  for (auto& x : th) x.join();
  if (!err.empty()) throw std::runtime_error(err);
}
void fill_trace_miller_loop(Trace& tr, const Fp& x, const Fp& y, const std::vector<Ell>& ell, size_t s, size_t e, size_t col) {
  namespace M = woff::miller_loop;
  put_big_rows(tr, s, e, col + M::PX_OFFSET, x, 12);
  put_big_rows(tr, s, e, col + M::PY_OFFSET, y, 12);
  // the accumulator chain first (values only: ~10 ms), then the 12-row blocks on several threads
  struct Block { Fp12 f12; bool first_bit, last_bit, bitone; };
  const size_t n_ops = std::min((e + 1 - s) / 12, ell.size());
  std::vector<Block> blocks(n_ops);
  Fp12 f12 = fp12_one();
  int i = 62;
  bool bitone = false;
  for (size_t j = 0; j < n_ops; j++) {
    blocks[j] = {f12, j == 0, i == 0, bitone};
    const Ell& c = ell[j];
    f12 = fp12_multiply_by_014(f12, c.c[0], fp2_mul_fp(c.c[1], x), fp2_mul_fp(c.c[2], y));
    if (((BLS_X >> i) & 1) && !bitone) bitone = true;
    else if (j + 1 < ell.size()) { f12 = fp12_mul_native(f12, f12); i--; bitone = false; }
  }
  parallel_blocks(n_ops, [&](size_t j) {
    const Block& b = blocks[j];
    const size_t s_row = s + 12 * j, e_row = s_row + 11;
    if (b.first_bit) tr.set_rows(s_row, e_row, col + M::FIRST_BIT_SELECTOR_OFFSET, 1);
    if (b.last_bit) tr.set_rows(s_row, e_row, col + M::LAST_BIT_SELECTOR_OFFSET, 1);
    if (b.bitone) tr.set_rows(s_row, e_row, col + M::BIT1_SELECTOR_OFFSET, 1);
    tr.set_rows(s_row, e_row, col + M::ELL_COEFFS_INDEX_OFFEST + j, 1);
    const Ell& c = ell[j];
    for (int k = 0; k < 3; k++) put_fp2_rows(tr, s_row, e_row, col + M::ELL_COEFFS_OFFSET + 24 * k, c.c[k]);
    put_fp12_rows(tr, s_row, e_row, col + M::F12_OFFSET, b.f12);
    if (j != 0) tr.at(s_row, col + M::FIRST_ROW_SELECTOR_OFFSET) = 1;
    fill_trace_fp2_fp_mul(tr, c.c[1], x, s_row, e_row, col + M::O1_CALC_OFFSET);
    const Fp2 o1 = fp2_mul_fp(c.c[1], x);
    fill_trace_fp2_fp_mul(tr, c.c[2], y, s_row, e_row, col + M::O4_CALC_OFFSET);
    const Fp2 o4 = fp2_mul_fp(c.c[2], y);
    fill_trace_multiply_by_014(tr, b.f12, c.c[0], o1, o4, s_row, e_row, col + M::F12_MUL_BY_014_OFFSET);
    const Fp12 g = fp12_multiply_by_014(b.f12, c.c[0], o1, o4);
    fill_trace_fp12_multiplication(tr, g, g, s_row, e_row, col + M::F12_SQ_CALC_OFFSET);
  });
  f12 = fp12_conjugate(f12);
  put_fp12_rows(tr, s, e, col + M::MILLER_LOOP_RES_OFFSET, f12);
  negate6_rows(tr, half(f12, 1), s, e, col + M::RES_CONJUGATE_OFFSET);
}
Fp12 fp12_cyclotomic_exponent(const Fp12& x) {                                               // native.rs:1300-1309
  Fp12 z = fp12_one();
  for (int i = 63; i >= 0; i--) {
    z = fp12_cyclotomic_square(z);
    if ((BLS_X >> i) & 1) z = fp12_mul_native(z, x);
  }
  return z;
}
Fp2 fp2_from_limbs(const u32* l) {
  Fp2 r = {{Big::from_limbs(l, 12), Big::from_limbs(l + 12, 12)}};
  for (int i = 0; i < 2; i++) if (cmp(r.c[i], MODP()) >= 0) throw std::invalid_argument("witness: Fp coefficient is not reduced modulo p");
  return r;
}
void pis_fp2(uint64_t* out, const Fp2& v) { for (int k = 0; k < 12; k++) { out[k] = v.c[0].w[k]; out[12 + k] = v.c[1].w[k]; } }

Fp12 fp12_from_limbs(const u32* l) {
  Fp12 r;
  for (int i = 0; i < 12; i++) {
    r.c[i] = Big::from_limbs(l + 12 * i, 12);
    if (cmp(r.c[i], MODP()) >= 0) throw std::invalid_argument("witness: Fp coefficient is not reduced modulo p");
  }
  return r;
}

// a fresh trace buffer is gigabytes of first-touch page faults: zero it from several threads (FinalExp 2.4 GB: 1.9 -> 0.1 s)
void zero_trace(void* p, size_t bytes) {
  if (bytes < (64u << 20)) { memset(p, 0, bytes); return; }
  const size_t parts = 8, step = (bytes / parts + 4095) & ~size_t(4095);
  std::vector<std::thread> th;
  for (size_t i = 0; i < parts; i++)
    th.emplace_back([=] { const size_t a = std::min(bytes, i * step), b = std::min(bytes, (i + 1) * step); memset((char*)p + a, 0, b - a); });
  for (auto& x : th) x.join();
}

thread_local std::string g_witness_error;

}  // namespace

// ------------------------------------------------------------------ g1.rs, ecc_aggregate.rs
struct G1 { Fp x, y; };
// g1.rs:26-255: chord addition of two affine points, x3 = l^2 - x2 - x1, y3 = l (x1 - x3) - y1, on rows s .. s + 11
G1 fill_trace_g1_addition(Trace& tr, const G1& p1, const G1& p2, size_t s, size_t col) {
  const Fp &x1 = p1.x, &y1 = p1.y, &x2 = p2.x, &y2 = p2.y;
  const Fp lam = fp_mul(fp_sub(y2, y1), fp_inv(fp_sub(x2, x1)));
  const Fp x3 = fp_sub(fp_sub(fp_mul(lam, lam), x2), x1);
  const Fp y3 = fp_sub(fp_mul(lam, fp_sub(x1, x3)), y1);
  const size_t e = s + 11;
  put_big_rows(tr, s, e, col + g1::G1_POINT_ADDITION_X1, x1, 12); put_big_rows(tr, s, e, col + g1::G1_POINT_ADDITION_Y1, y1, 12);
  put_big_rows(tr, s, e, col + g1::G1_POINT_ADDITION_X2, x2, 12); put_big_rows(tr, s, e, col + g1::G1_POINT_ADDITION_Y2, y2, 12);
  put_big_rows(tr, s, e, col + g1::G1_POINT_ADDITION_X3, x3, 12); put_big_rows(tr, s, e, col + g1::G1_POINT_ADDITION_Y3, y3, 12);
  auto add_rows = [&](const Big& a, const Big& b, size_t c) { fill_trace_addition_fp(tr, a, b, s, c); tr.rep(s, e, c, fp::FP_ADDITION_TOTAL); };
  auto sub_rows = [&](const Big& a, const Big& b, size_t c) { fill_trace_subtraction_fp(tr, a, b, s, c); tr.rep(s, e, c, fp::FP_SUBTRACTION_TOTAL); };
  auto mul_red = [&](const Big& a, const Big& b, size_t c) {
    fill_multiplication_trace_no_mod_reduction(tr, a, b, s, e, c);
    const Big res = fill_reduction_trace(tr, mul(a, b), s, e, c + fp::FP_MULTIPLICATION_TOTAL_COLUMNS);
    fill_range_check_trace(tr, res, e, c + fp::FP_MULTIPLICATION_TOTAL_COLUMNS + fp::REDUCTION_TOTAL);
    return res;
  };
  const Big& P = MODP();
  add_rows(x2, P, col + g1::X2_X1_DIFF);
  const Big x2_x1 = sub(add(x2, P), x1);
  sub_rows(add(x2, P), x1, col + g1::X2_X1_DIFF + fp::FP_ADDITION_TOTAL);
  add_rows(y2, P, col + g1::Y2_Y1_DIFF);
  const Big y2_y1 = sub(add(y2, P), y1);
  sub_rows(add(y2, P), y1, col + g1::Y2_Y1_DIFF + fp::FP_ADDITION_TOTAL);
  const Big x2_x1_sq = mul_red(x2_x1, x2_x1, col + g1::X2_X1_SQ);
  const Big y2_y1_sq = mul_red(y2_y1, y2_y1, col + g1::Y2_Y1_SQ);
  add_rows(x1, x2, col + g1::X1_X2_X3_SUM);
  add_rows(add(x1, x2), x3, col + g1::X1_X2_X3_SUM + fp::FP_ADDITION_TOTAL);
  const Big lhs = mul_red(add(add(x1, x2), x3), x2_x1_sq, col + g1::X1_X2_X3_X2_X1_SQ);
  if (cmp(lhs, y2_y1_sq)) throw std::logic_error("witness: g1 addition, x3 relation does not hold");
  add_rows(y1, y3, col + g1::Y1_Y3);
  add_rows(x1, P, col + g1::X1_X3);
  const Big x1_x3 = sub(add(x1, P), x3);
  sub_rows(add(x1, P), x3, col + g1::X1_X3 + fp::FP_ADDITION_TOTAL);
  const Big a = mul_red(add(y1, y3), x2_x1, col + g1::Y1_Y3_X2_X1);
  const Big b = mul_red(y2_y1, x1_x3, col + g1::Y2_Y1_X1_X3);
  if (cmp(a, b)) throw std::logic_error("witness: g1 addition, y3 relation does not hold");
  return {x3, y3};
}

extern "C" {

const char* sb_witness_last_error(void) { return g_witness_error.c_str(); }

// FP12MulStark::generate_trace (fp12_mul.rs:44-48) and the public inputs fp12_mul_main assembles (aggregate_proof.rs:124-151):
// x, y = Fp12 operands as 12 x 12 little-endian u32 limbs (the reference's Fp12 = [Fp; 12], Fp = [u32; 12]).
// trace_out: [num_rows][60285] uint32_t, row-major (SB_TRACE_ROWMAJOR_U32); public_inputs_out: 432 values x ++ y ++ x*y.
int sb_witness_fp12_mul(const uint32_t* x, const uint32_t* y, uint32_t num_rows, uint32_t* trace_out, uint64_t* public_inputs_out) {
  if (!x || !y || !trace_out || !public_inputs_out) return SB_EINVAL;
  try {
    if (num_rows < 12 || (num_rows & (num_rows - 1))) throw std::invalid_argument("witness: num_rows must be a power of two >= 16");
    const Fp12 X = fp12_from_limbs(x), Y = fp12_from_limbs(y);
    Trace tr = {trace_out, num_rows, woff::fp12_mul::TOTAL_COLUMNS};
    zero_trace(trace_out, 4ull * num_rows * tr.cols);
    fill_trace_fp12_multiplication(tr, X, Y, 0, 11, 0);
    const Fp12 Z = fp12_mul_native(X, Y);
    for (int i = 0; i < 12; i++)
      for (int k = 0; k < 12; k++) {
        public_inputs_out[woff::fp12_mul::PIS_INPUT_X_OFFSET + 12 * i + k] = X.c[i].w[k];
        public_inputs_out[woff::fp12_mul::PIS_INPUT_Y_OFFSET + 12 * i + k] = Y.c[i].w[k];
        public_inputs_out[woff::fp12_mul::PIS_OUTPUT_OFFSET + 12 * i + k] = Z.c[i].w[k];
      }
    return SB_OK;
  } catch (const std::exception& e) {
    g_witness_error = e.what();
    return SB_EINVAL;
  }
}

// ECCAggStark::generate_trace (ecc_aggregate.rs:37-82) + the public inputs of ec_aggregate_main (aggregate_proof.rs:186-227).
// points: 512 affine G1 points as x ++ y, 12 little-endian u32 limbs each ([512][24]); bits: 512 participation flags.
// trace_out: [num_rows][3339] uint32_t row-major; public_inputs_out: 12 824 values (points, bits, aggregate);
// result_out (optional): the aggregate point, 24 limbs.
int sb_witness_ecc_agg(const uint32_t* points, const uint8_t* bits, uint32_t num_rows, uint32_t* trace_out,
                       uint64_t* public_inputs_out, uint32_t* result_out) {
  if (!points || !bits || !trace_out || !public_inputs_out) return SB_EINVAL;
  try {
    namespace E = woff::ecc_aggregate;
    if ((num_rows & (num_rows - 1)) || (size_t)(E::NUM_POINTS - 1) * 12 >= num_rows)
      throw std::invalid_argument("witness: stark doesn't have enough rows (power of two > 12 * 511)");
    Trace tr = {trace_out, num_rows, E::TOTAL_COLUMNS};
    zero_trace(trace_out, 4ull * num_rows * tr.cols);
    std::vector<G1> pts(E::NUM_POINTS);
    for (uint32_t i = 0; i < E::NUM_POINTS; i++) {
      pts[i].x = Big::from_limbs(points + 24 * i, 12);
      pts[i].y = Big::from_limbs(points + 24 * i + 12, 12);
    }
    for (size_t r = 0; r < num_rows; r++) tr.at(r, E::ROW_NUM + r % 12) = 1;
    for (uint32_t i = 0; i < E::NUM_POINTS; i++) {
      const size_t row = i >= 2 ? 12ull * (i - 1) : 0;
      tr.set_rows(row, row + 11, E::PIS_IDX + i, 1);
    }
    size_t row = 0;
    G1 res = fill_trace_g1_addition(tr, pts[0], pts[1], row, E::OP);
    tr.set_rows(row, row + 11, E::A_IS_INF, bits[0] ? 0 : 1);
    tr.set_rows(row, row + 11, E::B_IS_INF, bits[1] ? 0 : 1);
    if (!bits[0]) res = pts[1];
    else if (!bits[1]) res = pts[0];
    for (uint32_t i = 2; i < E::NUM_POINTS; i++) {
      row += 12;
      const G1 tmp = fill_trace_g1_addition(tr, res, pts[i], row, E::OP);
      tr.set_rows(row, row + 11, E::B_IS_INF, bits[i] ? 0 : 1);
      if (bits[i]) res = tmp;
    }
    for (uint32_t i = 0; i < E::NUM_POINTS; i++) {
      for (int k = 0; k < 24; k++) public_inputs_out[E::POINTS + 24 * i + k] = points[24 * i + k];
      public_inputs_out[E::BITS + i] = bits[i] ? 1 : 0;
    }
    for (int k = 0; k < 12; k++) {
      public_inputs_out[E::RES + k] = res.x.w[k];
      public_inputs_out[E::RES + 12 + k] = res.y.w[k];
      if (result_out) { result_out[k] = res.x.w[k]; result_out[12 + k] = res.y.w[k]; }
    }
    return SB_OK;
  } catch (const std::exception& e) {
    g_witness_error = e.what();
    return SB_EINVAL;
  }
}


// PairingPrecompStark::generate_trace (calc_pairing_precomp.rs:150-366) + the public inputs of calc_pairing_precomp_main
// (aggregate_proof.rs:24-69).  q: the projective G2 point x ++ y ++ z, Fp2 each as 2 x 12 little-endian u32 limbs ([3][24]).
// trace_out: [num_rows][29376] uint32_t row-major; public_inputs_out: 4968 values (x, y, z, 68 x 3 line coefficients).
int sb_witness_pairing_precomp(const uint32_t* q, uint32_t num_rows, uint32_t* trace_out, uint64_t* public_inputs_out) {
  if (!q || !trace_out || !public_inputs_out) return SB_EINVAL;
  try {
    namespace A = woff::calc_pairing_precomp;
    if (num_rows < 16 || (num_rows & (num_rows - 1))) throw std::invalid_argument("witness: num_rows must be a power of two >= 16");
    const Fp2 x = fp2_from_limbs(q), y = fp2_from_limbs(q + 24), z = fp2_from_limbs(q + 48);
    Trace tr = {trace_out, num_rows, A::TOTAL_COLUMNS};
    zero_trace(trace_out, 4ull * num_rows * tr.cols);
    const size_t last = num_rows - 1;
    const Fp2 z_inv = fp2_inv(z);
    generate_trace_fp2_mul(tr, z, z_inv, 0, last, A::Z_MULT_Z_INV_OFFSET);
    generate_trace_fp2_mul(tr, x, z_inv, 0, last, A::X_MULT_Z_INV_OFFSET);
    generate_trace_fp2_mul(tr, y, z_inv, 0, last, A::Y_MULT_Z_INV_OFFSET);
    const Fp2 qx = fp2_mul(x, z_inv), qy = fp2_mul(y, z_inv), qz = {{Big(1), Big()}};          // calc_qs (native.rs:277-286)
    put_fp2_rows(tr, 0, last, A::QX_OFFSET, qx);
    put_fp2_rows(tr, 0, last, A::QY_OFFSET, qy);
    put_fp2_rows(tr, 0, last, A::QZ_OFFSET, qz);
    Fp2 rx = qx, ry = qy, rz = qz;
    int bit_pos = 62;
    bool bit1 = false;
    const size_t num_coeffs = 68;
    const Fp three = Big(3), two = Big(2);
    for (size_t n = 0; n < num_rows / 12 + 1; n++) {
      const size_t s = 12 * n, end_row = 12 * (n + 1), be = std::min(end_row, (size_t)num_rows) - 1;
      if (n == 0) tr.set_rows(s, be, A::FIRST_LOOP_SELECTOR_OFFSET, 1);
      put_fp2_rows(tr, s, be, A::RX_OFFSET, rx);
      put_fp2_rows(tr, s, be, A::RY_OFFSET, ry);
      put_fp2_rows(tr, s, be, A::RZ_OFFSET, rz);
      if (bit1) tr.set_rows(s, be, A::BIT1_SELECTOR_OFFSET, 1);
      if (n < num_coeffs) tr.set_rows(s, be, A::ELL_COEFFS_IDX_OFFSET + n, 1);
      tr.at(s, A::FIRST_ROW_SELECTOR_OFFSET) = 1;
      if (end_row > num_rows) break;
      const size_t e = end_row - 1;
      if (!bit1) {
        const Loop0 v = calc_precomp_stuff_loop0(rx, ry, rz);
        generate_trace_fp2_mul(tr, ry, ry, s, e, A::T0_CALC_OFFSET);
        generate_trace_fp2_mul(tr, rz, rz, s, e, A::T1_CALC_OFFSET);
        fill_trace_fp2_fp_mul(tr, v.t1, three, s, e, A::X0_CALC_OFFSET);
        fill_multiply_by_b_trace(tr, v.x0, s, e, A::T2_CALC_OFFSET);
        fill_trace_fp2_fp_mul(tr, v.t2, three, s, e, A::T3_CALC_OFFSET);
        generate_trace_fp2_mul(tr, ry, rz, s, e, A::X1_CALC_OFFSET);
        fill_trace_fp2_fp_mul(tr, v.x1, two, s, e, A::T4_CALC_OFFSET);
        sub_red_rows(tr, v.t2, v.t0, s, e, A::X2_CALC_OFFSET);
        generate_trace_fp2_mul(tr, rx, rx, s, e, A::X3_CALC_OFFSET);
        fill_trace_fp2_fp_mul(tr, v.x3, three, s, e, A::X4_CALC_OFFSET);
        negate_rows(tr, v.t4, s, e, A::X5_CALC_OFFSET);
        sub_red_rows(tr, v.t0, v.t3, s, e, A::X6_CALC_OFFSET);
        generate_trace_fp2_mul(tr, rx, ry, s, e, A::X7_CALC_OFFSET);
        generate_trace_fp2_mul(tr, v.x6, v.x7, s, e, A::X8_CALC_OFFSET);
        add_red_rows(tr, v.t0, v.t3, s, e, A::X9_CALC_OFFSET);
        fill_trace_fp2_fp_mul(tr, v.x9, HALF(), s, e, A::X10_CALC_OFFSET);
        generate_trace_fp2_mul(tr, v.x10, v.x10, s, e, A::X11_CALC_OFFSET);
        generate_trace_fp2_mul(tr, v.t2, v.t2, s, e, A::X12_CALC_OFFSET);
        fill_trace_fp2_fp_mul(tr, v.x12, three, s, e, A::X13_CALC_OFFSET);
        fill_trace_fp2_fp_mul(tr, v.x8, HALF(), s, e, A::NEW_RX_OFFSET);
        sub_red_rows(tr, v.x11, v.x13, s, e, A::NEW_RY_OFFSET);
        generate_trace_fp2_mul(tr, v.t0, v.t4, s, e, A::NEW_RZ_OFFSET);
        rx = v.new_rx; ry = v.new_ry; rz = v.new_rz;
        bit1 = (BLS_X >> bit_pos) & 1;
        if (!bit1) bit_pos = std::max(bit_pos - 1, 0);
      } else {
        const Loop1 w = calc_precomp_stuff_loop1(rx, ry, rz, qx, qy);
        const Fp2* t = w.t;
        generate_trace_fp2_mul(tr, qy, rz, s, e, A::BIT1_T0_CALC_OFFSET);
        sub_red_rows(tr, ry, t[0], s, e, A::BIT1_T1_CALC_OFFSET);
        generate_trace_fp2_mul(tr, qx, rz, s, e, A::BIT1_T2_CALC_OFFSET);
        sub_red_rows(tr, rx, t[2], s, e, A::BIT1_T3_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[1], qx, s, e, A::BIT1_T4_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[3], qy, s, e, A::BIT1_T5_CALC_OFFSET);
        sub_red_rows(tr, t[4], t[5], s, e, A::BIT1_T6_CALC_OFFSET);
        negate_rows(tr, t[1], s, e, A::BIT1_T7_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[3], t[3], s, e, A::BIT1_T8_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[8], t[3], s, e, A::BIT1_T9_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[8], rx, s, e, A::BIT1_T10_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[1], t[1], s, e, A::BIT1_T11_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[11], rz, s, e, A::BIT1_T12_CALC_OFFSET);
        fill_trace_fp2_fp_mul(tr, t[10], two, s, e, A::BIT1_T13_CALC_OFFSET);
        sub_red_rows(tr, t[9], t[13], s, e, A::BIT1_T14_CALC_OFFSET);
        add_red_rows(tr, t[14], t[12], s, e, A::BIT1_T15_CALC_OFFSET);
        sub_red_rows(tr, t[10], t[15], s, e, A::BIT1_T16_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[16], t[1], s, e, A::BIT1_T17_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[9], ry, s, e, A::BIT1_T18_CALC_OFFSET);
        generate_trace_fp2_mul(tr, t[3], t[15], s, e, A::BIT1_RX_CALC_OFFSET);
        sub_red_rows(tr, t[17], t[18], s, e, A::BIT1_RY_CALC_OFFSET);
        generate_trace_fp2_mul(tr, rz, t[9], s, e, A::BIT1_RZ_CALC_OFFSET);
        rx = w.new_rx; ry = w.new_ry; rz = w.new_rz;
        bit1 = false;
        bit_pos = std::max(bit_pos - 1, 0);
      }
    }
    pis_fp2(public_inputs_out, x); pis_fp2(public_inputs_out + 24, y); pis_fp2(public_inputs_out + 48, z);
    const std::vector<Ell> ell = pairing_precomp_native(x, y, z);
    if (72 + 72 * ell.size() != A::PUBLIC_INPUTS) throw std::logic_error("witness: pairing precomputation, coefficient count");
    for (size_t i = 0; i < ell.size(); i++)
      for (int k = 0; k < 3; k++) pis_fp2(public_inputs_out + A::ELL_COEFFS_PUBLIC_INPUTS_OFFSET + 72 * i + 24 * k, ell[i].c[k]);
    return SB_OK;
  } catch (const std::exception& e) {
    g_witness_error = e.what();
    return SB_EINVAL;
  }
}

// MillerLoopStark::generate_trace (miller_loop.rs:157-160, fill_trace_miller_loop :87-146) + the public inputs of
// miller_loop_main (aggregate_proof.rs:71-121).  p: the G1 point x ++ y ([24] limbs); q: the projective G2 point ([3][24]).
// trace_out: [num_rows][97330] uint32_t row-major; public_inputs_out: 5064 values (x, y, line coefficients, result).
int sb_witness_miller_loop(const uint32_t* p, const uint32_t* q, uint32_t num_rows, uint32_t* trace_out, uint64_t* public_inputs_out) {
  if (!p || !q || !trace_out || !public_inputs_out) return SB_EINVAL;
  try {
    namespace M = woff::miller_loop;
    if (num_rows < 16 || (num_rows & (num_rows - 1))) throw std::invalid_argument("witness: num_rows must be a power of two >= 16");
    const Fp x = Big::from_limbs(p, 12), y = Big::from_limbs(p + 12, 12);
    if (cmp(x, MODP()) >= 0 || cmp(y, MODP()) >= 0) throw std::invalid_argument("witness: Fp coefficient is not reduced modulo p");
    const Fp2 qx = fp2_from_limbs(q), qy = fp2_from_limbs(q + 24), qz = fp2_from_limbs(q + 48);
    const std::vector<Ell> ell = pairing_precomp_native(qx, qy, qz);
    const Fp12 res = miller_loop_native(x, y, ell);
    Trace tr = {trace_out, num_rows, M::TOTAL_COLUMNS};
    zero_trace(trace_out, 4ull * num_rows * tr.cols);
    fill_trace_miller_loop(tr, x, y, ell, 0, num_rows - 1, 0);
    if (24 + 72 * ell.size() + 144 != M::PUBLIC_INPUTS) throw std::logic_error("witness: miller loop, coefficient count");
    for (int k = 0; k < 12; k++) { public_inputs_out[M::PIS_PX_OFFSET + k] = x.w[k]; public_inputs_out[M::PIS_PY_OFFSET + k] = y.w[k]; }
    for (size_t i = 0; i < ell.size(); i++)
      for (int k = 0; k < 3; k++) pis_fp2(public_inputs_out + M::PIS_ELL_COEFFS_OFFSET + 72 * i + 24 * k, ell[i].c[k]);
    for (int i = 0; i < 12; i++)
      for (int k = 0; k < 12; k++) public_inputs_out[M::PIS_RES_OFFSET + 12 * i + k] = res.c[i].w[k];
    return SB_OK;
  } catch (const std::exception& e) {
    g_witness_error = e.what();
    return SB_EINVAL;
  }
}

// FinalExponentiateStark::generate_trace (final_exponentiate.rs:137-281) + the public inputs of final_exponentiate_main
// (aggregate_proof.rs:153-184).  x: the Fp12 input ([12][12] limbs).  trace_out: [8192][73527] uint32_t row-major (2.4 GB);
// public_inputs_out: 288 values (x, x^((p^12 - 1) / r)).
int sb_witness_final_exp(const uint32_t* x, uint32_t num_rows, uint32_t* trace_out, uint64_t* public_inputs_out) {
  if (!x || !trace_out || !public_inputs_out) return SB_EINVAL;
  try {
    namespace E = woff::final_exponentiate;
    if (num_rows != 8192) throw std::invalid_argument("witness: FinalExponentiateStark has 8192 rows (one row selector column per row)");
    const Fp12 X = fp12_from_limbs(x);
    Trace tr = {trace_out, num_rows, E::TOTAL_COLUMNS};
    zero_trace(trace_out, 4ull * num_rows * tr.cols);
    const size_t last = num_rows - 1, OP = E::FINAL_EXP_OP_OFFSET;
    for (size_t r = 0; r < num_rows; r++) tr.at(r, E::FINAL_EXP_ROW_SELECTORS + r) = 1;
    put_fp12_rows(tr, 0, last, E::FINAL_EXP_INPUT_OFFSET, X);
    const u32 ROW[33] = {E::T0_ROW, E::T1_ROW, E::T2_ROW, E::T3_ROW, E::T4_ROW, E::T5_ROW, E::T6_ROW, E::T7_ROW, E::T8_ROW, E::T9_ROW, E::T10_ROW,
                         E::T11_ROW, E::T12_ROW, E::T13_ROW, E::T14_ROW, E::T15_ROW, E::T16_ROW, E::T17_ROW, E::T18_ROW, E::T19_ROW, E::T20_ROW,
                         E::T21_ROW, E::T22_ROW, E::T23_ROW, E::T24_ROW, E::T25_ROW, E::T26_ROW, E::T27_ROW, E::T28_ROW, E::T29_ROW, E::T30_ROW,
                         E::T31_ROW, E::TOTAL_ROW};
    const u32 OFF[32] = {E::FINAL_EXP_T0_OFFSET, E::FINAL_EXP_T1_OFFSET, E::FINAL_EXP_T2_OFFSET, E::FINAL_EXP_T3_OFFSET, E::FINAL_EXP_T4_OFFSET,
                         E::FINAL_EXP_T5_OFFSET, E::FINAL_EXP_T6_OFFSET, E::FINAL_EXP_T7_OFFSET, E::FINAL_EXP_T8_OFFSET, E::FINAL_EXP_T9_OFFSET,
                         E::FINAL_EXP_T10_OFFSET, E::FINAL_EXP_T11_OFFSET, E::FINAL_EXP_T12_OFFSET, E::FINAL_EXP_T13_OFFSET, E::FINAL_EXP_T14_OFFSET,
                         E::FINAL_EXP_T15_OFFSET, E::FINAL_EXP_T16_OFFSET, E::FINAL_EXP_T17_OFFSET, E::FINAL_EXP_T18_OFFSET, E::FINAL_EXP_T19_OFFSET,
                         E::FINAL_EXP_T20_OFFSET, E::FINAL_EXP_T21_OFFSET, E::FINAL_EXP_T22_OFFSET, E::FINAL_EXP_T23_OFFSET, E::FINAL_EXP_T24_OFFSET,
                         E::FINAL_EXP_T25_OFFSET, E::FINAL_EXP_T26_OFFSET, E::FINAL_EXP_T27_OFFSET, E::FINAL_EXP_T28_OFFSET, E::FINAL_EXP_T29_OFFSET,
                         E::FINAL_EXP_T30_OFFSET, E::FINAL_EXP_T31_OFFSET};
    // The 32 steps of native.rs:1311-1345.  Their values first (a few ms), then the trace blocks: step k fills rows
    // ROW[k] .. ROW[k+1]-1 of the operation columns and its own result columns, so the steps are filled by independent
    // host threads (the five cyclotomic exponentiations, 841 rows each, are most of the 2.4 GB).
    enum Kind { FROB, MUL, DIV, CEXP, CONJ, CSQ };
    struct Step { Kind kind; int a, b; unsigned pw; };                          // operands: index into t, -1 = the input x
    static const Step STEPS[32] = {
        {FROB, -1, 0, 6}, {DIV, 0, -1, 0}, {FROB, 1, 0, 2}, {MUL, 2, 1, 0}, {CEXP, 3, 0, 0}, {CONJ, 4, 0, 0}, {CSQ, 3, 0, 0}, {CONJ, 6, 0, 0},
        {MUL, 7, 5, 0}, {CEXP, 8, 0, 0}, {CONJ, 9, 0, 0}, {CEXP, 10, 0, 0}, {CONJ, 11, 0, 0}, {CEXP, 12, 0, 0}, {CONJ, 13, 0, 0}, {CSQ, 5, 0, 0},
        {MUL, 14, 15, 0}, {CEXP, 16, 0, 0}, {CONJ, 17, 0, 0}, {MUL, 5, 12, 0}, {FROB, 19, 0, 2}, {MUL, 10, 3, 0}, {FROB, 21, 0, 3}, {CONJ, 3, 0, 0},
        {MUL, 16, 23, 0}, {FROB, 24, 0, 1}, {CONJ, 8, 0, 0}, {MUL, 18, 26, 0}, {MUL, 27, 3, 0}, {MUL, 20, 22, 0}, {MUL, 29, 25, 0}, {MUL, 30, 28, 0}};
    Fp12 t[32];
    auto val = [&](int i) -> const Fp12& { return i < 0 ? X : t[i]; };
    for (int k = 0; k < 32; k++) {
      const Step& st = STEPS[k];
      const Fp12& a = val(st.a);
      switch (st.kind) {
        case FROB: t[k] = fp12_frobenius(a, st.pw); break;
        case MUL: t[k] = fp12_mul_native(a, val(st.b)); break;
        case DIV: t[k] = fp12_mul_native(a, fp12_inv(val(st.b))); break;       // the trace proves t[k] * b = a
        case CEXP: t[k] = fp12_cyclotomic_exponent(a); break;
        case CONJ: t[k] = fp12_conjugate(a); break;
        case CSQ: t[k] = fp12_cyclotomic_square(a); break;
      }
    }
    auto fill_step = [&](int k) {
      const Step& st = STEPS[k];
      const Fp12& a = val(st.a);
      const size_t s = ROW[k], e = ROW[k + 1] - 1;
      switch (st.kind) {
        case FROB:                                                               // fill_trace_forbenius
          tr.set_rows(s, e, E::FINAL_EXP_FORBENIUS_MAP_SELECTOR, 1);
          fill_trace_fp12_forbenius_map(tr, a, st.pw, s, e, OP);
          break;
        case MUL:                                                                // fill_trace_mul
          tr.set_rows(s, e, E::FINAL_EXP_MUL_SELECTOR, 1);
          fill_trace_fp12_multiplication(tr, a, val(st.b), s, e, OP);
          break;
        case DIV:                                                                // fill_trace_div
          tr.set_rows(s, e, E::FINAL_EXP_MUL_SELECTOR, 1);
          fill_trace_fp12_multiplication(tr, t[k], val(st.b), s, e, OP);
          break;
        case CEXP: {                                                             // fill_trace_cyc_exp
          tr.set_rows(s, e, E::FINAL_EXP_CYCLOTOMIC_EXP_SELECTOR, 1);
          const Fp12 z = fill_trace_cyclotomic_exp(tr, a, s, e, OP);
          for (int i = 0; i < 12; i++) if (cmp(z.c[i], t[k].c[i])) throw std::logic_error("witness: cyclotomic exponentiation, trace and value differ");
          break;
        }
        case CONJ:                                                               // fill_trace_conjugate
          tr.at(s, E::FINAL_EXP_CONJUGATE_SELECTOR) = 1;
          fill_trace_fp12_conjugate(tr, a, s, OP);
          break;
        case CSQ:                                                                // fill_trace_cyc_sq
          tr.set_rows(s, e, E::FINAL_EXP_CYCLOTOMIC_SQ_SELECTOR, 1);
          fill_trace_cyclotomic_sq(tr, a, s, e, OP);
          break;
      }
      put_fp12_rows(tr, 0, last, OFF[k], t[k]);
    };
    {
      // longest first: the five exponentiations, then everything else
      std::vector<int> order;
      for (int k = 0; k < 32; k++) if (STEPS[k].kind == CEXP) order.push_back(k);
      for (int k = 0; k < 32; k++) if (STEPS[k].kind != CEXP) order.push_back(k);
      parallel_blocks(order.size(), [&](size_t i) { fill_step(order[i]); });
    }
    for (int i = 0; i < 12; i++)
      for (int k = 0; k < 12; k++) {
        public_inputs_out[E::PIS_INPUT_OFFSET + 12 * i + k] = X.c[i].w[k];
        public_inputs_out[E::PIS_OUTPUT_OFFSET + 12 * i + k] = t[31].c[i].w[k];
      }
    return SB_OK;
  } catch (const std::exception& e) {
    g_witness_error = e.what();
    return SB_EINVAL;
  }
}

}  // extern "C"
