// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of plonky2 0.1.4 (Electron-Labs/plonky2 @ 666f315, un-vendored) PoseidonHash over
// Goldilocks, MerkleTree and Challenger as reached from starky::prover::prove
// (/root/reference/src/aggregate_proof.rs:59 etc. with C = PoseidonGoldilocksConfig,
// aggregate_proof.rs:236).  Spec: SURVEY.md Appendix A.3-A.5.  The permutation is the *naive*
// round function (add constants, S-box, full MDS) -- plonky2's own unit test asserts the fast
// partial-round schedule equals it.  Pinned by plonky2's two permutation KATs (tests/test_oracle_kat.py).
#pragma once
#include "gl.h"
#include "poseidon_rc.h"
#include <string.h>
#include <assert.h>

namespace orc {

static const u64 POSEIDON_RC[POSEIDON_RC_COUNT] = POSEIDON_RC_TABLE;
static const u64 MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const u64 MDS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

static inline u64 sbox7(u64 x) {
  u64 x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x2, x);
  return gl_mul(x3, x4);
}

static inline void poseidon_permute(u64 s[12]) {
  for (int r = 0; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], POSEIDON_RC[12 * r + i]);
    if (r < 4 || r >= 26) { for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]); }
    else s[0] = sbox7(s[0]);
    u64 t[12];
    for (int row = 0; row < 12; row++) {
      u128 acc = 0;
      for (int i = 0; i < 12; i++) acc += (u128)MDS_CIRC[i] * s[(i + row) % 12];
      acc += (u128)MDS_DIAG[row] * s[row];
      t[row] = gl_reduce128(acc);
    }
    memcpy(s, t, sizeof(t));
  }
}

struct Hash { u64 e[4]; };

// hash_no_pad: overwrite-mode sponge, rate 8, no padding (A.3).
static inline Hash hash_no_pad(const u64* in, size_t n) {
  u64 s[12] = {0};
  for (size_t off = 0; off < n; off += 8) {
    size_t len = n - off < 8 ? n - off : 8;
    for (size_t i = 0; i < len; i++) s[i] = in[off + i];
    poseidon_permute(s);
  }
  Hash h; memcpy(h.e, s, 32); return h;
}
static inline Hash hash_or_noop(const u64* in, size_t n) {
  if (n <= 4) { Hash h = {{0, 0, 0, 0}}; for (size_t i = 0; i < n; i++) h.e[i] = in[i]; return h; }
  return hash_no_pad(in, n);
}
static inline Hash two_to_one(const Hash& l, const Hash& r) {
  u64 s[12] = {l.e[0], l.e[1], l.e[2], l.e[3], r.e[0], r.e[1], r.e[2], r.e[3], 0, 0, 0, 0};
  poseidon_permute(s);
  Hash h; memcpy(h.e, s, 32); return h;
}

// MerkleTree::new(leaves, cap_height) (A.4).  levels[0] = leaf digests, levels[k] has L>>k nodes,
// the last level is the cap (2^cap_height nodes).
struct MerkleTree {
  size_t n_leaves = 0, leaf_len = 0;
  unsigned cap_height = 0;
  std::vector<u64> leaves;                 // [n_leaves][leaf_len]
  std::vector<std::vector<Hash>> levels;
  const u64* leaf(size_t i) const { return &leaves[i * leaf_len]; }
  const std::vector<Hash>& cap() const { return levels.back(); }
  void build() {
    unsigned log_l = 0; while ((size_t(1) << log_l) < n_leaves) log_l++;
    assert((size_t(1) << log_l) == n_leaves && cap_height <= log_l);
    levels.clear(); levels.emplace_back(n_leaves);
    std::vector<Hash>& d = levels[0];
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n_leaves; i++) d[i] = hash_or_noop(leaf(i), leaf_len);
    for (unsigned l = log_l; l > cap_height; l--) {
      const std::vector<Hash>& prev = levels.back();
      std::vector<Hash> cur(prev.size() / 2);
#pragma omp parallel for schedule(static)
      for (long i = 0; i < (long)cur.size(); i++) cur[i] = two_to_one(prev[2 * i], prev[2 * i + 1]);
      levels.push_back(std::move(cur));
    }
  }
  // siblings bottom-up, length log2(L) - cap_height
  std::vector<Hash> prove(size_t idx) const {
    std::vector<Hash> sib;
    for (size_t l = 0; l + 1 < levels.size(); l++) sib.push_back(levels[l][(idx >> l) ^ 1]);
    return sib;
  }
};

static inline bool merkle_verify(const u64* leaf, size_t leaf_len, size_t idx, const Hash* cap, size_t cap_len,
                                 const Hash* sib, size_t n_sib) {
  Hash cur = hash_or_noop(leaf, leaf_len);
  for (size_t i = 0; i < n_sib; i++) {
    cur = (idx & 1) ? two_to_one(sib[i], cur) : two_to_one(cur, sib[i]);
    idx >>= 1;
  }
  if (idx >= cap_len) return false;
  return memcmp(cur.e, cap[idx].e, 32) == 0;
}

// Duplex-sponge Challenger (A.5).
struct Challenger {
  u64 state[12];
  u64 in_buf[8]; int n_in;
  u64 out_buf[8]; int n_out;
  Challenger() { memset(state, 0, sizeof(state)); n_in = 0; n_out = 0; }
  void duplexing() {
    for (int i = 0; i < n_in; i++) state[i] = in_buf[i];
    n_in = 0;
    poseidon_permute(state);
    memcpy(out_buf, state, 64); n_out = 8;
  }
  void observe(u64 x) { n_out = 0; in_buf[n_in++] = x; if (n_in == 8) duplexing(); }
  void observe_hash(const Hash& h) { for (int i = 0; i < 4; i++) observe(h.e[i]); }
  void observe_cap(const std::vector<Hash>& cap) { for (size_t i = 0; i < cap.size(); i++) observe_hash(cap[i]); }
  void observe_ext(E2 x) { observe(x.a); observe(x.b); }
  u64 challenge() { if (n_in > 0 || n_out == 0) duplexing(); return out_buf[--n_out]; }
  E2 ext_challenge() { u64 a = challenge(); u64 b = challenge(); return e2(a, b); }
};

}  // namespace orc
