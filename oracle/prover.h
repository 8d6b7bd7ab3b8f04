// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of starky 0.1.2 `prover::prove` / `verifier::verify_stark_proof` and the plonky2
// 0.1.4 pieces they call (PolynomialBatch::{from_values,from_coeffs,prove_openings}, fri_proof,
// verify_fri_proof) at Electron-Labs/plonky2 @ 666f315 -- an un-vendored git dependency of
// /root/reference (Cargo.toml:9-10), so the algorithm is restated from SURVEY.md Appendix A.7-A.10
// and anchored on the reference's call sites (aggregate_proof.rs:59,67,105,113,138,146,169,177,212,220).
// PARITY STATUS: the field/Poseidon layer is pinned by plonky2's KATs; the transcript order, PoW and
// FRI leaf layout are recalled, not checked against a Rust run ("parity unpinned" at proof-byte level;
// see DESIGN.md).  Loops are OpenMP-parallel over columns / leaves / points exactly where plonky2 uses rayon.
#pragma once
#include "air.h"
#include "hash.h"
#include "poly.h"
#include <stdexcept>

namespace orc {

// Binary-identical to sb_params (include/starky_b200.h) so the tests can pass the same struct.
struct Params {
  uint32_t stark_id, log_n, n_cols, n_public_inputs, constraint_degree, rate_bits, cap_height, num_challenges,
      pow_bits, num_query_rounds, fri_arity_bits, fri_final_poly_bits, flags, reserved;
  uint64_t fixed_pow_witness;
};
enum { FLAG_ALLOW_INVALID = 1u, FLAG_FIXED_POW = 2u, FLAG_OBSERVE_PIS = 4u, FLAG_FRI_MUL_X = 8u };

// Binary-identical to sb_proof_layout.
struct Layout {
  uint32_t log_n, log_lde, n_cols, nq, n_pis, cap_len, n_fri_rounds, final_poly_len, n_queries, arity_bits,
      trace_path_len, reserved;
  uint64_t off_trace_cap, off_quotient_cap, off_local, off_next, off_quot_open, off_fri_caps, off_final_poly,
      off_pow, off_queries, query_stride, q_trace_leaf, q_trace_path, q_quot_leaf, q_quot_path, q_steps, off_pis,
      total_words;
};

static inline unsigned qdf_of(const Params& p) { return p.constraint_degree > 1 ? p.constraint_degree - 1 : 1; }
static inline unsigned log2_ceil_u(unsigned x) { unsigned b = 0; while ((1u << b) < x) b++; return b; }

// FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits) (A.6)
static inline std::vector<unsigned> fri_arities(const Params& p) {
  std::vector<unsigned> r;
  unsigned db = p.log_n;
  if (p.fri_arity_bits == 0) throw std::invalid_argument("fri_arity_bits is 0");
  while (db > p.fri_final_poly_bits) {
    // plonky2: `degree_bits + rate_bits - arity_bits >= cap_height` in usize (a negative difference panics or wraps to
    // "true"), then assert!(degree_bits >= arity_bits) inside the loop
    if (db + p.rate_bits >= p.fri_arity_bits && db + p.rate_bits - p.fri_arity_bits < p.cap_height) break;
    if (db < p.fri_arity_bits) throw std::invalid_argument("FRI reduction: degree_bits < arity_bits");
    r.push_back(p.fri_arity_bits); db -= p.fri_arity_bits;
  }
  return r;
}
static inline unsigned step_path_len(const Layout& l, unsigned round) {
  unsigned log_leaves = l.log_lde - l.arity_bits * (round + 1);
  unsigned cap_h = log2_ceil_u(l.cap_len);
  return log_leaves - cap_h;
}
static inline uint64_t step_offset(const Layout& l, unsigned round) {
  uint64_t o = l.q_steps;
  for (unsigned r = 0; r < round; r++) o += (uint64_t(2) << l.arity_bits) + 4 * step_path_len(l, r);
  return o;
}
static inline Layout layout_for(const Params& p) {
  Layout l; memset(&l, 0, sizeof(l));
  std::vector<unsigned> ar = fri_arities(p);
  l.log_n = p.log_n; l.log_lde = p.log_n + p.rate_bits; l.n_cols = p.n_cols;
  l.nq = p.num_challenges * qdf_of(p); l.n_pis = p.n_public_inputs; l.cap_len = 1u << p.cap_height;
  l.n_fri_rounds = (uint32_t)ar.size(); l.final_poly_len = 1u << (p.log_n - p.fri_arity_bits * ar.size());
  l.n_queries = p.num_query_rounds; l.arity_bits = p.fri_arity_bits; l.trace_path_len = l.log_lde - p.cap_height;
  uint64_t o = 0;
  l.off_trace_cap = o; o += 4ull * l.cap_len;
  l.off_quotient_cap = o; o += 4ull * l.cap_len;
  l.off_local = o; o += 2ull * l.n_cols;
  l.off_next = o; o += 2ull * l.n_cols;
  l.off_quot_open = o; o += 2ull * l.nq;
  l.off_fri_caps = o; o += 4ull * l.cap_len * l.n_fri_rounds;
  l.off_final_poly = o; o += 2ull * l.final_poly_len;
  l.off_pow = o; o += 1;
  l.off_queries = o;
  uint64_t q = 0;
  l.q_trace_leaf = q; q += l.n_cols;
  l.q_trace_path = q; q += 4ull * l.trace_path_len;
  l.q_quot_leaf = q; q += l.nq;
  l.q_quot_path = q; q += 4ull * l.trace_path_len;
  l.q_steps = q;
  for (unsigned r = 0; r < l.n_fri_rounds; r++) q += (uint64_t(2) << l.arity_bits) + 4ull * step_path_len(l, r);
  l.query_stride = q;
  o += q * l.n_queries;
  l.off_pis = o; o += l.n_pis;
  l.total_words = o;
  return l;
}

// PolynomialBatch (A.2/A.4): coefficients + Merkle tree whose leaf j = all LDE values at natural index bitrev(j).
struct Batch {
  unsigned log_n = 0, rate_bits = 0;
  std::vector<std::vector<u64>> coeffs;  // [polys][n]
  MerkleTree tree;
  size_t n_polys() const { return coeffs.size(); }
  const u64* lde_row(size_t natural_index) const { return tree.leaf(bitrev((unsigned)natural_index, log_n + rate_bits)); }
};

static inline void batch_commit(Batch& b, unsigned cap_height) {
  unsigned log_lde = b.log_n + b.rate_bits;
  size_t N = size_t(1) << log_lde, P = b.coeffs.size();
  b.tree.n_leaves = N; b.tree.leaf_len = P; b.tree.cap_height = cap_height;
  b.tree.leaves.assign(N * P, 0);
#pragma omp parallel for schedule(dynamic, 8)
  for (long c = 0; c < (long)P; c++) {
    std::vector<u64> v = lde_coset(b.coeffs[c], b.log_n, b.rate_bits);
    for (size_t i = 0; i < N; i++) b.tree.leaves[size_t(bitrev((unsigned)i, log_lde)) * P + c] = v[i];
  }
  b.tree.build();
}
// from_values: values column-major [P][n]
static inline void batch_from_values(Batch& b, const u64* values, size_t P, unsigned log_n, unsigned rate_bits,
                                     unsigned cap_height) {
  size_t n = size_t(1) << log_n;
  b.log_n = log_n; b.rate_bits = rate_bits; b.coeffs.assign(P, std::vector<u64>());
#pragma omp parallel for schedule(dynamic, 8)
  for (long c = 0; c < (long)P; c++) {
    b.coeffs[c].assign(values + c * n, values + (c + 1) * n);
    ifft(b.coeffs[c].data(), log_n);
  }
  batch_commit(b, cap_height);
}

// compute_quotient_polys up to the per-point values q_j(x_i) (A.8); out[j][i], natural order, size n << qbits.
static inline void quotient_values(const Air& air, const Batch& trace, const Params& p, const u64* pis,
                                   const u64* alphas, std::vector<std::vector<u64>>& out) {
  unsigned qbits = log2_ceil_u(qdf_of(p));
  if (qbits > p.rate_bits) throw std::runtime_error("constraint degree higher than the rate is not supported");
  unsigned log_size = p.log_n + qbits;
  size_t size = size_t(1) << log_size, n = size_t(1) << p.log_n;
  size_t step = size_t(1) << (p.rate_bits - qbits), next_step = size_t(1) << qbits;
  u64 g = gl_root(p.log_n), g_inv = gl_inv(g), w = gl_root(log_size), n_f = gl_canon(n);
  out.assign(p.num_challenges, std::vector<u64>(size));
#pragma omp parallel
  {
    std::vector<u64> scratch(air.nodes.size());
    std::vector<u64> acc(p.num_challenges);
#pragma omp for schedule(dynamic, 4)
    for (long i = 0; i < (long)size; i++) {
      u64 x = gl_mul(GL_GEN, gl_pow(w, (u64)i));
      const u64* lv = trace.lde_row(size_t(i) * step);
      const u64* nv = trace.lde_row(((size_t(i) + next_step) % size) * step);
      u64 zh = gl_sub(gl_pow(x, n), 1);
      u64 z_last = gl_sub(x, g_inv);
      u64 l_first = gl_mul(zh, gl_inv(gl_mul(n_f, gl_sub(x, 1))));
      u64 l_last = gl_mul(zh, gl_inv(gl_mul(n_f, gl_sub(gl_mul(g, x), 1))));
      air.eval_base(lv, nv, pis, z_last, l_first, l_last, alphas, p.num_challenges, acc.data(), scratch.data());
      u64 zh_inv = gl_inv(zh);
      for (unsigned j = 0; j < p.num_challenges; j++) out[j][i] = gl_mul(acc[j], zh_inv);
    }
  }
}

static inline E2 eval_poly_base_at_ext(const std::vector<u64>& c, E2 z) {
  E2 acc = e2(0);
  for (size_t i = c.size(); i-- > 0;) acc = e2_add(e2_mul(acc, z), e2(c[i]));
  return acc;
}
static inline E2 eval_poly_ext(const E2* c, size_t n, E2 z) {
  E2 acc = e2(0);
  for (size_t i = n; i-- > 0;) acc = e2_add(e2_mul(acc, z), c[i]);
  return acc;
}

static inline void put_hash(u64* w, const Hash& h) { memcpy(w, h.e, 32); }
static inline void put_cap(u64* w, const std::vector<Hash>& cap) { for (size_t i = 0; i < cap.size(); i++) put_hash(w + 4 * i, cap[i]); }

// PoW (A.9): smallest witness whose duplex response has >= pow_bits leading zeros.
static inline bool pow_ok(const u64 inter[12], int pos, u64 cand, unsigned bits) {
  u64 s[12]; memcpy(s, inter, sizeof(s)); s[pos] = cand; poseidon_permute(s);
  return bits == 0 || (s[7] >> (64 - bits)) == 0;
}
static inline u64 pow_grind(const Challenger& ch, unsigned bits) {
  u64 inter[12]; memcpy(inter, ch.state, sizeof(inter));
  for (int i = 0; i < ch.n_in; i++) inter[i] = ch.in_buf[i];
  int pos = ch.n_in;
  const long BLOCK = 1 << 14;
  for (u64 base = 0;; base += BLOCK) {
    u64 best = ~u64(0);
#pragma omp parallel for schedule(static) reduction(min : best)
    for (long k = 0; k < BLOCK; k++)
      if (pow_ok(inter, pos, base + k, bits) && base + k < best) best = base + k;
    if (best != ~u64(0)) return best;
  }
}

// fri_committed_trees (A.9): values of the (zero-padded) polynomial on the current coset, arity-sized leaves in
// bit-reversed order, Merkle tree to the cap, then `next_beta(round, cap)` and the fold in coefficient form.
// Returns the final polynomial's coefficients (length n >> sum of arity bits).
template <class BetaFn>
static inline std::vector<E2> fri_commit_phase(const Params& p, const std::vector<E2>& final_poly, std::vector<MerkleTree>& fri_trees,
                                               BetaFn&& next_beta) {
  const unsigned r = p.rate_bits, log_lde = p.log_n + r;
  const size_t N = size_t(1) << log_lde;
  std::vector<E2> coeffs = final_poly; coeffs.resize(N, e2(0));
  std::vector<E2> values = coset_fft_e2(coeffs, log_lde, GL_GEN);
  std::vector<unsigned> ar = fri_arities(p);
  fri_trees.assign(ar.size(), MerkleTree());
  u64 shift = GL_GEN;
  unsigned cur_log = log_lde;
  for (size_t round = 0; round < ar.size(); round++) {
    unsigned ab = ar[round]; size_t arity = size_t(1) << ab;
    reverse_index_bits(values);
    MerkleTree& t = fri_trees[round];
    t.n_leaves = values.size() / arity; t.leaf_len = 2 * arity; t.cap_height = p.cap_height;
    t.leaves.resize(values.size() * 2);
    for (size_t i = 0; i < values.size(); i++) { t.leaves[2 * i] = values[i].a; t.leaves[2 * i + 1] = values[i].b; }
    t.build();
    E2 beta = next_beta(round, t.cap());
    std::vector<E2> folded(coeffs.size() / arity);
    for (size_t m = 0; m < folded.size(); m++) {
      E2 acc = e2(0);
      for (size_t i = arity; i-- > 0;) acc = e2_add(e2_mul(acc, beta), coeffs[arity * m + i]);
      folded[m] = acc;
    }
    coeffs.swap(folded);
    shift = gl_pow(shift, arity);
    cur_log -= ab;
    values = coset_fft_e2(coeffs, cur_log, shift);
  }
  coeffs.resize(coeffs.size() >> r);
  return coeffs;
}

struct ProofOut { Layout layout; std::vector<u64> words; };

static inline int prove(const Air& air, const Params& p, const u64* trace_colmajor, const u64* pis, ProofOut& out,
                        std::string* err) {
  const unsigned log_n = p.log_n, r = p.rate_bits, log_lde = log_n + r, qdf = qdf_of(p);
  const size_t n = size_t(1) << log_n, N = size_t(1) << log_lde, C = p.n_cols;
  Layout L = layout_for(p);
  out.layout = L; out.words.assign(L.total_words, 0);
  u64* W = out.words.data();

  // 1. trace commitment
  Batch trace; batch_from_values(trace, trace_colmajor, C, log_n, r, p.cap_height);
  Challenger ch;
  if (p.flags & FLAG_OBSERVE_PIS) for (unsigned i = 0; i < p.n_public_inputs; i++) ch.observe(pis[i]);
  ch.observe_cap(trace.tree.cap());
  put_cap(W + L.off_trace_cap, trace.tree.cap());
  // 2. alphas
  std::vector<u64> alphas(p.num_challenges);
  for (unsigned j = 0; j < p.num_challenges; j++) alphas[j] = ch.challenge();
  // 3. quotient
  std::vector<std::vector<u64>> qv;
  quotient_values(air, trace, p, pis, alphas.data(), qv);
  unsigned qbits = log2_ceil_u(qdf);
  Batch quot; quot.log_n = log_n; quot.rate_bits = r;
  for (unsigned j = 0; j < p.num_challenges; j++) {
    coset_ifft(qv[j].data(), log_n + qbits, GL_GEN);
    for (size_t i = qdf * n; i < qv[j].size(); i++)
      if (qv[j][i] != 0 && !(p.flags & FLAG_ALLOW_INVALID)) {
        if (err) *err = "Quotient has failed, the vanishing polynomial is not divisible by Z_H";
        return -4;
      }
    for (unsigned c = 0; c < qdf; c++) quot.coeffs.emplace_back(qv[j].begin() + c * n, qv[j].begin() + (c + 1) * n);
  }
  batch_commit(quot, p.cap_height);
  ch.observe_cap(quot.tree.cap());
  put_cap(W + L.off_quotient_cap, quot.tree.cap());
  // 4. zeta
  E2 zeta = ch.ext_challenge();
  { E2 t = zeta; for (unsigned i = 0; i < log_n; i++) t = e2_mul(t, t);
    if (e2_eq(t, e2(1))) { if (err) *err = "Opening point is in the subgroup."; return -5; } }
  u64 g = gl_root(log_n);
  E2 zeta_next = e2_scale(zeta, g);
  // 5. openings
  std::vector<E2> local(C), next(C), qopen(L.nq);
#pragma omp parallel for schedule(dynamic, 16)
  for (long c = 0; c < (long)C; c++) {
    local[c] = eval_poly_base_at_ext(trace.coeffs[c], zeta);
    next[c] = eval_poly_base_at_ext(trace.coeffs[c], zeta_next);
  }
  for (unsigned q = 0; q < L.nq; q++) qopen[q] = eval_poly_base_at_ext(quot.coeffs[q], zeta);
  for (size_t c = 0; c < C; c++) { W[L.off_local + 2 * c] = local[c].a; W[L.off_local + 2 * c + 1] = local[c].b;
                                   W[L.off_next + 2 * c] = next[c].a; W[L.off_next + 2 * c + 1] = next[c].b; }
  for (unsigned q = 0; q < L.nq; q++) { W[L.off_quot_open + 2 * q] = qopen[q].a; W[L.off_quot_open + 2 * q + 1] = qopen[q].b; }
  for (size_t c = 0; c < C; c++) ch.observe_ext(local[c]);          // batch 0 = local ++ quotient
  for (unsigned q = 0; q < L.nq; q++) ch.observe_ext(qopen[q]);
  for (size_t c = 0; c < C; c++) ch.observe_ext(next[c]);           // batch 1 = next
  // 6. prove_openings (A.9)
  E2 alpha = ch.ext_challenge();
  std::vector<E2> final_poly(n, e2(0));
  for (int batch = 0; batch < 2; batch++) {
    E2 z = batch == 0 ? zeta : zeta_next;
    size_t n_polys = batch == 0 ? C + L.nq : C;
    // composition F = sum_j alpha^j f_j
    std::vector<E2> apow(n_polys);
    { E2 a = e2(1); for (size_t j = 0; j < n_polys; j++) { apow[j] = a; a = e2_mul(a, alpha); } }
    std::vector<E2> comp(n, e2(0));
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)n; k++) {
      E2 s = e2(0);
      for (size_t j = 0; j < n_polys; j++) {
        u64 cf = j < C ? trace.coeffs[j][k] : quot.coeffs[j - C][k];
        s = e2_add(s, e2_scale(apow[j], cf));
      }
      comp[k] = s;
    }
    // divide_by_linear(z), padded back to n coefficients
    std::vector<E2> q(n, e2(0));
    { E2 acc = e2(0); for (size_t k = n; k-- > 1;) { acc = e2_add(e2_mul(acc, z), comp[k]); q[k - 1] = acc; } }
    // final <- final * alpha^(#polys in this batch) + q
    E2 shift = e2(1); for (size_t j = 0; j < n_polys; j++) shift = e2_mul(shift, alpha);
    for (size_t k = 0; k < n; k++) final_poly[k] = e2_add(e2_mul(final_poly[k], shift), q[k]);
  }
  if (p.flags & FLAG_FRI_MUL_X) { final_poly.insert(final_poly.begin(), e2(0)); final_poly.pop_back(); }
  // commit phase
  std::vector<unsigned> ar = fri_arities(p);
  std::vector<MerkleTree> fri_trees;
  std::vector<E2> coeffs = fri_commit_phase(p, final_poly, fri_trees, [&](size_t round, const std::vector<Hash>& cap) {
    ch.observe_cap(cap);
    put_cap(W + L.off_fri_caps + round * 4ull * L.cap_len, cap);
    return ch.ext_challenge();
  });
  if (coeffs.size() != L.final_poly_len) { if (err) *err = "internal: final poly length"; return -1; }
  for (size_t i = 0; i < coeffs.size(); i++) { W[L.off_final_poly + 2 * i] = coeffs[i].a; W[L.off_final_poly + 2 * i + 1] = coeffs[i].b; ch.observe_ext(coeffs[i]); }
  // PoW
  u64 witness = (p.flags & FLAG_FIXED_POW) ? p.fixed_pow_witness : pow_grind(ch, p.pow_bits);
  ch.observe(witness);
  u64 response = ch.challenge();
  if (p.pow_bits && (response >> (64 - p.pow_bits)) != 0) { if (err) *err = "Proof of work failed"; return -6; }
  W[L.off_pow] = witness;
  // queries
  for (unsigned qi = 0; qi < p.num_query_rounds; qi++) {
    size_t x = ch.challenge() % N;
    u64* Q = W + L.off_queries + qi * L.query_stride;
    memcpy(Q + L.q_trace_leaf, trace.tree.leaf(x), 8 * C);
    { std::vector<Hash> s = trace.tree.prove(x); for (size_t i = 0; i < s.size(); i++) put_hash(Q + L.q_trace_path + 4 * i, s[i]); }
    memcpy(Q + L.q_quot_leaf, quot.tree.leaf(x), 8 * L.nq);
    { std::vector<Hash> s = quot.tree.prove(x); for (size_t i = 0; i < s.size(); i++) put_hash(Q + L.q_quot_path + 4 * i, s[i]); }
    for (size_t round = 0; round < ar.size(); round++) {
      size_t leaf = x >> ar[round];
      u64* S = Q + step_offset(L, (unsigned)round);
      memcpy(S, fri_trees[round].leaf(leaf), 8 * (size_t(2) << ar[round]));
      std::vector<Hash> s = fri_trees[round].prove(leaf);
      for (size_t i = 0; i < s.size(); i++) put_hash(S + (size_t(2) << ar[round]) + 4 * i, s[i]);
      x = leaf;
    }
  }
  memcpy(W + L.off_pis, pis, 8 * p.n_public_inputs);
  return 0;
}

// verify_stark_proof (A.10).  Returns 0 if accepted, otherwise a negative code and *err.
static inline int verify(const Air& air, const Params& p, const u64* W, size_t n_words, std::string* err) {
#define VFAIL(code, msg) do { if (err) *err = (msg); return (code); } while (0)
  Layout L = layout_for(p);
  if (n_words != L.total_words) VFAIL(-1, "proof size mismatch");
  const unsigned log_n = p.log_n, log_lde = L.log_lde, qdf = qdf_of(p);
  const size_t N = size_t(1) << log_lde, C = p.n_cols;
  for (size_t i = 0; i < n_words; i++) if (W[i] >= GL_P) VFAIL(-1, "non-canonical field element in proof");
  const u64* pis = W + L.off_pis;
  std::vector<Hash> trace_cap(L.cap_len), quot_cap(L.cap_len);
  memcpy(trace_cap.data(), W + L.off_trace_cap, 32 * L.cap_len);
  memcpy(quot_cap.data(), W + L.off_quotient_cap, 32 * L.cap_len);
  // challenges
  Challenger ch;
  if (p.flags & FLAG_OBSERVE_PIS) for (unsigned i = 0; i < p.n_public_inputs; i++) ch.observe(pis[i]);
  ch.observe_cap(trace_cap);
  std::vector<u64> alphas(p.num_challenges);
  for (unsigned j = 0; j < p.num_challenges; j++) alphas[j] = ch.challenge();
  ch.observe_cap(quot_cap);
  E2 zeta = ch.ext_challenge();
  std::vector<E2> local(C), next(C), qopen(L.nq);
  for (size_t c = 0; c < C; c++) { local[c] = e2(W[L.off_local + 2 * c], W[L.off_local + 2 * c + 1]);
                                   next[c] = e2(W[L.off_next + 2 * c], W[L.off_next + 2 * c + 1]); }
  for (unsigned q = 0; q < L.nq; q++) qopen[q] = e2(W[L.off_quot_open + 2 * q], W[L.off_quot_open + 2 * q + 1]);
  for (size_t c = 0; c < C; c++) ch.observe_ext(local[c]);
  for (unsigned q = 0; q < L.nq; q++) ch.observe_ext(qopen[q]);
  for (size_t c = 0; c < C; c++) ch.observe_ext(next[c]);
  E2 alpha = ch.ext_challenge();
  std::vector<unsigned> ar = fri_arities(p);
  std::vector<E2> betas(ar.size());
  std::vector<std::vector<Hash>> fri_caps(ar.size(), std::vector<Hash>(L.cap_len));
  for (size_t round = 0; round < ar.size(); round++) {
    memcpy(fri_caps[round].data(), W + L.off_fri_caps + round * 4ull * L.cap_len, 32 * L.cap_len);
    ch.observe_cap(fri_caps[round]);
    betas[round] = ch.ext_challenge();
  }
  std::vector<E2> final_poly(L.final_poly_len);
  for (size_t i = 0; i < final_poly.size(); i++) { final_poly[i] = e2(W[L.off_final_poly + 2 * i], W[L.off_final_poly + 2 * i + 1]); ch.observe_ext(final_poly[i]); }
  ch.observe(W[L.off_pow]);
  u64 pow_response = ch.challenge();
  std::vector<size_t> indices(p.num_query_rounds);
  for (unsigned qi = 0; qi < p.num_query_rounds; qi++) indices[qi] = ch.challenge() % N;

  // constraint check at zeta
  u64 g = gl_root(log_n);
  E2 zeta_pow = zeta; for (unsigned i = 0; i < log_n; i++) zeta_pow = e2_mul(zeta_pow, zeta_pow);
  E2 zh = e2_sub(zeta_pow, e2(1));
  u64 n_f = u64(1) << log_n;
  E2 l_first = e2_mul(zh, e2_inv(e2_scale(e2_sub(zeta, e2(1)), n_f)));
  E2 l_last = e2_mul(zh, e2_inv(e2_scale(e2_sub(e2_scale(zeta, g), e2(1)), n_f)));
  E2 z_last = e2_sub(zeta, e2(gl_inv(g)));
  {
    std::vector<E2> scratch(air.nodes.size()), acc(p.num_challenges);
    air.eval_ext(local.data(), next.data(), pis, z_last, l_first, l_last, alphas.data(), p.num_challenges, acc.data(), scratch.data());
    for (unsigned j = 0; j < p.num_challenges; j++) {
      E2 t = e2(0);
      for (unsigned i = qdf; i-- > 0;) t = e2_add(e2_mul(t, zeta_pow), qopen[j * qdf + i]);
      if (!e2_eq(acc[j], e2_mul(zh, t))) VFAIL(-10, "Mismatch between evaluation and opening of quotient polynomial");
    }
  }
  // FRI
  if (p.pow_bits && (pow_response >> (64 - p.pow_bits)) != 0) VFAIL(-11, "Invalid proof of work witness.");
  E2 zeta_next = e2_scale(zeta, g);
  // precomputed reduced openings: sum_j alpha^j opening_j per batch
  E2 red0 = e2(0), red1 = e2(0);
  for (unsigned q = L.nq; q-- > 0;) red0 = e2_add(e2_mul(red0, alpha), qopen[q]);
  for (size_t c = C; c-- > 0;) red0 = e2_add(e2_mul(red0, alpha), local[c]);
  for (size_t c = C; c-- > 0;) red1 = e2_add(e2_mul(red1, alpha), next[c]);
  E2 alpha_pow_c = e2_pow(alpha, C);
  for (unsigned qi = 0; qi < p.num_query_rounds; qi++) {
    size_t x_index = indices[qi];
    const u64* Q = W + L.off_queries + qi * L.query_stride;
    if (!merkle_verify(Q + L.q_trace_leaf, C, x_index, trace_cap.data(), L.cap_len, (const Hash*)(Q + L.q_trace_path), L.trace_path_len))
      VFAIL(-12, "Invalid Merkle proof (trace).");
    if (!merkle_verify(Q + L.q_quot_leaf, L.nq, x_index, quot_cap.data(), L.cap_len, (const Hash*)(Q + L.q_quot_path), L.trace_path_len))
      VFAIL(-12, "Invalid Merkle proof (quotient).");
    u64 subgroup_x = gl_mul(GL_GEN, gl_pow(gl_root(log_lde), bitrev((unsigned)x_index, log_lde)));
    // fri_combine_initial
    E2 e0 = e2(0), e1 = e2(0);
    for (unsigned q = L.nq; q-- > 0;) e0 = e2_add(e2_mul(e0, alpha), e2(Q[L.q_quot_leaf + q]));
    for (size_t c = C; c-- > 0;) e0 = e2_add(e2_mul(e0, alpha), e2(Q[L.q_trace_leaf + c]));
    for (size_t c = C; c-- > 0;) e1 = e2_add(e2_mul(e1, alpha), e2(Q[L.q_trace_leaf + c]));
    E2 sum = e2_mul(e2_sub(e0, red0), e2_inv(e2_sub(e2(subgroup_x), zeta)));
    sum = e2_add(e2_mul(sum, alpha_pow_c), e2_mul(e2_sub(e1, red1), e2_inv(e2_sub(e2(subgroup_x), zeta_next))));
    E2 old_eval = sum;
    if (p.flags & FLAG_FRI_MUL_X) old_eval = e2_scale(old_eval, subgroup_x);
    for (size_t round = 0; round < ar.size(); round++) {
      unsigned ab = ar[round]; size_t arity = size_t(1) << ab;
      const u64* S = Q + step_offset(L, (unsigned)round);
      size_t coset_index = x_index >> ab, within = x_index & (arity - 1);
      std::vector<E2> evals(arity);
      for (size_t i = 0; i < arity; i++) evals[i] = e2(S[2 * i], S[2 * i + 1]);
      if (!e2_eq(evals[within], old_eval)) VFAIL(-13, "FRI fold consistency check failed");
      // compute_evaluation: interpolate the coset at beta
      u64 ga = gl_root(ab);
      std::vector<E2> ev = evals; reverse_index_bits(ev);
      unsigned rev_within = bitrev((unsigned)within, ab);
      u64 coset_start = gl_mul(subgroup_x, gl_pow(ga, arity - rev_within));
      std::vector<u64> xs(arity);
      { u64 y = 1; for (size_t i = 0; i < arity; i++) { xs[i] = gl_mul(coset_start, y); y = gl_mul(y, ga); } }
      E2 res = e2(0);
      for (size_t i = 0; i < arity; i++) {   // Lagrange interpolation at beta
        E2 num = e2(1); u64 den = 1;
        for (size_t k = 0; k < arity; k++) if (k != i) { num = e2_mul(num, e2_sub(betas[round], e2(xs[k]))); den = gl_mul(den, gl_sub(xs[i], xs[k])); }
        res = e2_add(res, e2_mul(ev[i], e2_scale(num, gl_inv(den))));
      }
      old_eval = res;
      if (!merkle_verify(S, 2 * arity, coset_index, fri_caps[round].data(), L.cap_len, (const Hash*)(S + 2 * arity), step_path_len(L, (unsigned)round)))
        VFAIL(-12, "Invalid Merkle proof (FRI step).");
      for (unsigned i = 0; i < ab; i++) subgroup_x = gl_mul(subgroup_x, subgroup_x);
      x_index = coset_index;
    }
    if (!e2_eq(eval_poly_ext(final_poly.data(), final_poly.size(), e2(subgroup_x)), old_eval))
      VFAIL(-14, "Final polynomial evaluation is invalid.");
  }
  return 0;
#undef VFAIL
}

}  // namespace orc
