// Witness generation in C++: the reference's generate_trace / fill_trace_* functions restated for the
// host side of the drop-in, so that a caller can hand the prover a few field elements instead of a multi-GB trace.
//   fp.rs:185-428   addition / subtraction / multiply_single / reduce_single / range check / 12x12-limb multiplication / reduction
//   fp2.rs:187-456  Fp2 addition, subtraction, multiplication, +- with reduction, non-residue multiplication
//   fp6.rs:124-303  Fp6 addition, subtraction, +- with reduction, non-residue multiplication, multiplication
//   fp12.rs:186-232 Fp12 multiplication;   fp12_mul.rs:44-48 FP12MulStark::generate_trace
// over the BLS12-381 tower of native.rs (quirks kept because they show in the trace: add_fp subtracts p at most once,
// -x is p - x).  Column offsets come from the reference's constants (witness_offsets.h, generated from
// witness/offsets.json).  Cells are written as row-major uint32_t -- every cell the reference writes is a u32 limb, a carry
// or a bit (utils.rs:7-19) -- which is SB_TRACE_ROWMAJOR_U32, half the PCIe bytes of the u64 layouts.
// The Python restatement (starky_bls12_381_b200/witness) is the checker: tests/test_witness_cpp.py compares cell for cell.
#include <stdint.h>
#include <string.h>

#include <stdexcept>
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

Fp12 fp12_from_limbs(const u32* l) {
  Fp12 r;
  for (int i = 0; i < 12; i++) {
    r.c[i] = Big::from_limbs(l + 12 * i, 12);
    if (cmp(r.c[i], MODP()) >= 0) throw std::invalid_argument("witness: Fp coefficient is not reduced modulo p");
  }
  return r;
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
    memset(trace_out, 0, 4ull * num_rows * tr.cols);
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
    memset(trace_out, 0, 4ull * num_rows * tr.cols);
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

}  // extern "C"
