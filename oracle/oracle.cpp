// ORACLE -- TEST INFRASTRUCTURE ONLY.  C entry points (ctypes) over the CPU restatement in gl.h / hash.h /
// poly.h / air.h / prover.h.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (starky_bls12_381_b200/) never does.
// PARITY UNPINNED at proof-byte level: the reference's prove() lives in un-vendored Rust git dependencies (plonky2 @
// 666f315) and cannot be built here; this restatement follows SURVEY.md Appendix A.  Pinned: Poseidon-12 (plonky2's
// published KATs), Goldilocks constants, the five starks' constraint counts / fingerprints (tests/test_oracle_kat.py,
// tests/test_air_programs.py, tests/golden/); the constraint programs it interprets, by the reference's own witness
// logic: every constraint vanishes on every row of generated traces of all five starks, and the BLS12-381 arithmetic
// under those generators passes the reference's KATs (tests/test_witness.py, tests/golden/bls_kats.json).
#include "prover.h"
#include <map>
#include <memory>
#include <mutex>
#include <omp.h>

using namespace orc;

static thread_local std::string g_err;
static std::mutex g_air_mu;
static std::map<std::string, std::shared_ptr<Air>> g_airs;

static std::shared_ptr<Air> get_air(const char* path) {
  std::lock_guard<std::mutex> lk(g_air_mu);
  auto it = g_airs.find(path);
  if (it != g_airs.end()) return it->second;
  auto a = std::make_shared<Air>();
  if (!a->load(path, &g_err)) return nullptr;
  g_airs[path] = a;
  return a;
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

void orc_poseidon_permute(uint64_t* state) { poseidon_permute(state); }
void orc_hash_no_pad(const uint64_t* in, size_t n, uint64_t* out4) { Hash h = hash_no_pad(in, n); memcpy(out4, h.e, 32); }
void orc_hash_or_noop(const uint64_t* in, size_t n, uint64_t* out4) { Hash h = hash_or_noop(in, n); memcpy(out4, h.e, 32); }
void orc_two_to_one(const uint64_t* l, const uint64_t* r, uint64_t* out4) {
  Hash a, b; memcpy(a.e, l, 32); memcpy(b.e, r, 32); Hash h = two_to_one(a, b); memcpy(out4, h.e, 32);
}
uint64_t orc_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }
uint64_t orc_gl_mul_slow(uint64_t a, uint64_t b) { return gl_mul_slow(a, b); }
uint64_t orc_gl_root(unsigned log_n) { return gl_root(log_n); }

// natural-order forward / inverse NTT of `count` vectors
void orc_ntt_batch(uint64_t* data, unsigned log_n, unsigned count, int inverse) {
#pragma omp parallel for schedule(dynamic, 4)
  for (long c = 0; c < (long)count; c++) { if (inverse) ifft(data + ((size_t)c << log_n), log_n); else fft(data + ((size_t)c << log_n), log_n); }
}

// Challenger replay: observe `n_obs` elements then draw `n_out` challenges.
void orc_challenger_run(const uint64_t* obs, size_t n_obs, uint64_t* out, size_t n_out) {
  Challenger ch; for (size_t i = 0; i < n_obs; i++) ch.observe(obs[i]);
  for (size_t i = 0; i < n_out; i++) out[i] = ch.challenge();
}

int orc_layout(const Params* p, Layout* out) {
  try { *out = layout_for(*p); return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// PolynomialBatch::from_values.  leaves_out [N][C] (plonky2 leaf order), digests_out [N][4], cap_out [2^cap][4];
// coeffs_out [C][n].  Any output may be NULL.
int orc_lde_commit(const Params* p, const uint64_t* trace_colmajor, uint64_t* leaves_out, uint64_t* digests_out,
                   uint64_t* cap_out, uint64_t* coeffs_out) {
  Batch b; batch_from_values(b, trace_colmajor, p->n_cols, p->log_n, p->rate_bits, p->cap_height);
  size_t N = b.tree.n_leaves, n = size_t(1) << p->log_n;
  if (leaves_out) memcpy(leaves_out, b.tree.leaves.data(), 8 * b.tree.leaves.size());
  if (digests_out) memcpy(digests_out, b.tree.levels[0].data(), 32 * N);
  if (cap_out) memcpy(cap_out, b.tree.cap().data(), 32 * b.tree.cap().size());
  if (coeffs_out) for (size_t c = 0; c < p->n_cols; c++) memcpy(coeffs_out + c * n, b.coeffs[c].data(), 8 * n);
  return 0;
}

// Merkle tree over explicit leaves [n_leaves][leaf_len]; returns cap and (optionally) all leaf digests.
int orc_merkle(const uint64_t* leaves, size_t n_leaves, size_t leaf_len, unsigned cap_height, uint64_t* digests_out,
               uint64_t* cap_out) {
  MerkleTree t; t.n_leaves = n_leaves; t.leaf_len = leaf_len; t.cap_height = cap_height;
  t.leaves.assign(leaves, leaves + n_leaves * leaf_len); t.build();
  if (digests_out) memcpy(digests_out, t.levels[0].data(), 32 * n_leaves);
  if (cap_out) memcpy(cap_out, t.cap().data(), 32 * t.cap().size());
  return 0;
}

// q_j(x_i) before coset_ifft; out [num_challenges][n << qbits] natural order.
int orc_quotient_values(const char* air_path, const Params* p, const uint64_t* trace_colmajor, const uint64_t* pis,
                        const uint64_t* alphas, uint64_t* out) {
  auto air = get_air(air_path); if (!air) return -8;
  if (air->h.n_cols != p->n_cols || air->h.n_pis != p->n_public_inputs) { g_err = "AIR shape mismatch"; return -1; }
  try {
    Batch b; batch_from_values(b, trace_colmajor, p->n_cols, p->log_n, p->rate_bits, p->cap_height);
    std::vector<std::vector<u64>> q; quotient_values(*air, b, *p, pis, alphas, q);
    for (size_t j = 0; j < q.size(); j++) memcpy(out + j * q[j].size(), q[j].data(), 8 * q[j].size());
  } catch (std::exception& e) { g_err = e.what(); return -1; }
  return 0;
}

// Evaluate every constraint at one (local,next) row pair; out[k] = c_k (no class factor, no folding).
int orc_eval_constraints_row(const char* air_path, const uint64_t* local, const uint64_t* next, const uint64_t* pis,
                             uint64_t* out) {
  auto air = get_air(air_path); if (!air) return -8;
  std::vector<u64> scratch(air->nodes.size());
  u64 alpha = 0, acc = 0;
  air->eval_base(local, next, pis, 1, 1, 1, &alpha, 1, &acc, scratch.data());
  for (size_t k = 0; k < air->cons.size(); k++) out[k] = scratch[air->cons[k].node];
  return 0;
}
int orc_air_info(const char* air_path, uint32_t* out6) {
  auto air = get_air(air_path); if (!air) return -8;
  out6[0] = air->h.n_cols; out6[1] = air->h.n_pis; out6[2] = air->h.degree; out6[3] = air->h.n_consts;
  out6[4] = air->h.n_nodes; out6[5] = air->h.n_constraints;
  return 0;
}

int orc_prove(const char* air_path, const Params* p, const uint64_t* trace_colmajor, const uint64_t* pis,
              uint64_t* words_out, size_t capacity) {
  auto air = get_air(air_path); if (!air) return -8;
  if (air->h.n_cols != p->n_cols || air->h.n_pis != p->n_public_inputs) { g_err = "AIR shape mismatch"; return -1; }
  ProofOut po; int rc;
  try { rc = prove(*air, *p, trace_colmajor, pis, po, &g_err); } catch (std::exception& e) { g_err = e.what(); return -1; }
  if (rc) return rc;
  if (capacity < po.words.size()) { g_err = "output buffer too small"; return -1; }
  memcpy(words_out, po.words.data(), 8 * po.words.size());
  return 0;
}

// stage export for the parity test of sb_fri_commit: commit phase with injected folding challenges
int orc_fri_commit(const Params* p, const uint64_t* coeffs, const uint64_t* betas, uint64_t* caps_out, uint64_t* final_poly_out) {
  try {
    const size_t n = size_t(1) << p->log_n, cap_len = size_t(1) << p->cap_height;
    std::vector<E2> fp(n);
    for (size_t i = 0; i < n; i++) fp[i] = e2(coeffs[2 * i], coeffs[2 * i + 1]);
    std::vector<MerkleTree> trees;
    std::vector<E2> out = fri_commit_phase(*p, fp, trees, [&](size_t round, const std::vector<Hash>& cap) {
      put_cap(caps_out + round * 4 * cap_len, cap);
      return e2(betas[2 * round], betas[2 * round + 1]);
    });
    for (size_t i = 0; i < out.size(); i++) { final_poly_out[2 * i] = out[i].a; final_poly_out[2 * i + 1] = out[i].b; }
    return (int)out.size();
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int orc_verify(const char* air_path, const Params* p, const uint64_t* words, size_t n_words) {
  auto air = get_air(air_path); if (!air) return -8;
  try { return verify(*air, *p, words, n_words, &g_err); } catch (std::exception& e) { g_err = e.what(); return -1; }
}

}  // extern "C"
