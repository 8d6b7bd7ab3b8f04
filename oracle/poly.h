// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of plonky2_field 0.1.1 polynomial routines (fft.rs / polynomial/mod.rs of
// Electron-Labs/plonky2 @ 666f315, un-vendored): ifft, fft, lde, coset_fft, coset_ifft,
// reverse_index_bits.  Spec: SURVEY.md Appendix A.2.  Reached from PolynomialBatch::from_values /
// from_coeffs inside starky::prover::prove (/root/reference/src/aggregate_proof.rs:59 ...).
// Textbook iterative radix-2 with an explicit bit-reversal: every exact algorithm is bit-identical.
#pragma once
#include "gl.h"
#include <algorithm>

namespace orc {

template <class T> static inline void reverse_index_bits(std::vector<T>& v) {
  unsigned bits = 0; while ((size_t(1) << bits) < v.size()) bits++;
  for (size_t i = 0; i < v.size(); i++) { size_t j = bitrev((unsigned)i, bits); if (i < j) std::swap(v[i], v[j]); }
}

// in-place natural-order DFT: out[k] = sum_i v[i] w^(ik), w = root (a primitive len-th root or its inverse)
static inline void dft_inplace(u64* v, unsigned log_n, u64 root) {
  size_t n = size_t(1) << log_n;
  for (size_t i = 0; i < n; i++) { size_t j = bitrev((unsigned)i, log_n); if (i < j) std::swap(v[i], v[j]); }
  for (unsigned s = 1; s <= log_n; s++) {
    size_t m = size_t(1) << s, half = m >> 1;
    u64 wm = gl_pow(root, n >> s);
    for (size_t k = 0; k < n; k += m) {
      u64 w = 1;
      for (size_t j = 0; j < half; j++) {
        u64 t = gl_mul(w, v[k + j + half]), u = v[k + j];
        v[k + j] = gl_add(u, t); v[k + j + half] = gl_sub(u, t);
        w = gl_mul(w, wm);
      }
    }
  }
}
static inline void fft(u64* v, unsigned log_n) { dft_inplace(v, log_n, gl_root(log_n)); }
static inline void ifft(u64* v, unsigned log_n) {
  dft_inplace(v, log_n, gl_inv(gl_root(log_n)));
  u64 ninv = gl_inv(u64(1) << log_n);
  for (size_t i = 0; i < (size_t(1) << log_n); i++) v[i] = gl_mul(v[i], ninv);
}
// coeffs (len 2^log_n) -> values on shift*<w>
static inline void coset_fft(u64* c, unsigned log_n, u64 shift) {
  u64 s = 1;
  for (size_t i = 0; i < (size_t(1) << log_n); i++) { c[i] = gl_mul(c[i], s); s = gl_mul(s, shift); }
  fft(c, log_n);
}
static inline void coset_ifft(u64* v, unsigned log_n, u64 shift) {
  ifft(v, log_n);
  u64 si = gl_inv(shift), s = 1;
  for (size_t i = 0; i < (size_t(1) << log_n); i++) { v[i] = gl_mul(v[i], s); s = gl_mul(s, si); }
}
// coefficient vector (len n) -> LDE values P(7 w_N^i), natural order, len n << rate_bits
static inline std::vector<u64> lde_coset(const std::vector<u64>& coeffs, unsigned log_n, unsigned rate_bits) {
  std::vector<u64> v(size_t(1) << (log_n + rate_bits), 0);
  std::copy(coeffs.begin(), coeffs.end(), v.begin());
  coset_fft(v.data(), log_n + rate_bits, GL_GEN);
  return v;
}

// Extension-valued versions (coefficients in F_p^2, domain in F_p): componentwise.
static inline void split_e2(const std::vector<E2>& v, std::vector<u64>& a, std::vector<u64>& b) {
  a.resize(v.size()); b.resize(v.size());
  for (size_t i = 0; i < v.size(); i++) { a[i] = v[i].a; b[i] = v[i].b; }
}
static inline std::vector<E2> join_e2(const std::vector<u64>& a, const std::vector<u64>& b) {
  std::vector<E2> v(a.size());
  for (size_t i = 0; i < a.size(); i++) v[i] = e2(a[i], b[i]);
  return v;
}
static inline std::vector<E2> coset_fft_e2(const std::vector<E2>& coeffs, unsigned log_len, u64 shift) {
  std::vector<u64> a, b; split_e2(coeffs, a, b);
  a.resize(size_t(1) << log_len, 0); b.resize(size_t(1) << log_len, 0);
  coset_fft(a.data(), log_len, shift); coset_fft(b.data(), log_len, shift);
  return join_e2(a, b);
}

}  // namespace orc
