// Proof wire formats: how a proof leaves the C ABI for a consumer that does not link the library --
// the reference's verifier and recursion circuit consume a starky::proof::StarkProofWithPublicInputs<F, C, 2>
// (/root/reference/src/aggregate_proof.rs:67,113,146,177,220 verify_stark_proof; :435-439 recursive verifier).
//
//   SB_WIRE_POD             "SBPROOF1" | sb_params (64 B) | total_words u64 | words (little-endian u64): the flat POD of
//                           include/starky_b200.h with the parameters that define its layout; self-describing
//   SB_WIRE_PLONKY2_BUFFER  the byte stream plonky2::util::serialization::Write produces when the fields of
//                           StarkProofWithPublicInputs are written in declaration order with its primitives:
//                             write_merkle_cap    : cap_len x write_hash (4 x u64 LE), no length prefix
//                             write_field_ext_vec : elements x (c0, c1) u64 LE, no length prefix (lengths come from the
//                                                   stark / config on the reading side, as in read_*(…, len))
//                             write_fri_proof     : commit-phase caps; per query round: per oracle write_field_vec(leaf) +
//                                                   write_merkle_proof (u8 length, siblings); per step
//                                                   write_field_ext_vec(evals) + write_merkle_proof; final_poly
//                                                   coefficients; pow_witness
//                             public_inputs       : write_field_vec
//                           (Option fields permutation_zs_cap / permutation_zs / permutation_zs_next are None for the
//                           five starks -- no permutation argument -- and occupy no bytes)
//   SB_WIRE_SERDE_JSON      the serde_json shape of the same struct: MerkleCap = [ {"elements":[u64;4]}, .. ], extension
//                           elements = [c0, c1], MerkleProof = {"siblings":[..]}, FriProof exactly as its
//                           #[derive(Serialize)] prints it, so `serde_json::from_value::<FriProof<F, PoseidonHash, 2>>`
//                           reads "opening_proof" as is (rust/starky_gpu/src/wire.rs)
// All host code; works in a process without a GPU.
#include <inttypes.h>
#include <string.h>

#include <string>

#include "prover.cuh"


namespace {

struct Sink {                      // counts when buf == nullptr
  unsigned char* buf; size_t cap, len = 0; bool overflow = false;
  void put(const void* p, size_t n) {
    if (buf) { if (len + n <= cap) memcpy(buf + len, p, n); else overflow = true; }
    len += n;
  }
  void u8(uint8_t x) { put(&x, 1); }
  void u64s(const u64* w, size_t n) { put(w, 8 * n); }      // little-endian host
  void str(const char* s) { put(s, strlen(s)); }
  void num(u64 x) { char b[24]; int n = snprintf(b, sizeof(b), "%" PRIu64, x); put(b, (size_t)n); }
};

void merkle_proof_bytes(Sink& s, const u64* sib, uint32_t len) {
  if (len > 255) SB_THROW(SB_EINVAL, "Merkle proof length must fit in u8");
  s.u8((uint8_t)len);
  s.u64s(sib, 4ull * len);
}

void to_buffer(Sink& s, const sb_proof* pr) {
  const sb_proof_layout& l = pr->layout;
  const u64* W = pr->words;
  s.u64s(W + l.off_trace_cap, 4ull * l.cap_len);
  s.u64s(W + l.off_quotient_cap, 4ull * l.cap_len);
  s.u64s(W + l.off_local_values, 2ull * l.n_cols);
  s.u64s(W + l.off_next_values, 2ull * l.n_cols);
  s.u64s(W + l.off_quotient_polys, 2ull * l.n_quotient_polys);
  s.u64s(W + l.off_fri_caps, 4ull * l.cap_len * l.n_fri_rounds);
  for (uint32_t q = 0; q < l.n_queries; q++) {
    const u64* Q = W + l.off_queries + (uint64_t)q * l.query_stride;
    s.u64s(Q + l.q_off_trace_leaf, l.n_cols);
    merkle_proof_bytes(s, Q + l.q_off_trace_path, l.trace_path_len);
    s.u64s(Q + l.q_off_quot_leaf, l.n_quotient_polys);
    merkle_proof_bytes(s, Q + l.q_off_quot_path, l.trace_path_len);
    for (uint32_t r = 0; r < l.n_fri_rounds; r++) {
      const u64* S = Q + fri_step_offset(l, r);
      s.u64s(S, 2ull << l.arity_bits);
      merkle_proof_bytes(s, S + (2ull << l.arity_bits), fri_step_path_len(l, r));
    }
  }
  s.u64s(W + l.off_final_poly, 2ull * l.final_poly_len);
  s.u64s(W + l.off_pow_witness, 1);
  s.u64s(W + l.off_public_inputs, l.n_public_inputs);
}

struct Source {
  const unsigned char* buf; size_t len, pos = 0;
  void get(void* p, size_t n) {
    if (pos + n > len) SB_THROW(SB_EINVAL, "serialized proof is truncated (need %zu bytes at offset %zu of %zu)", n, pos, len);
    memcpy(p, buf + pos, n);
    pos += n;
  }
  void u64s(u64* w, size_t n) {
    get(w, 8 * n);
    for (size_t i = 0; i < n; i++) if (w[i] >= GL_P) SB_THROW(SB_EINVAL, "serialized proof holds a non-canonical field element at byte %zu", pos - 8 * (n - i));
  }
  void merkle_proof(u64* sib, uint32_t len_expected) {
    uint8_t n;
    get(&n, 1);
    if (n != len_expected) SB_THROW(SB_EINVAL, "Merkle proof of length %u where %u is expected", (unsigned)n, len_expected);
    u64s(sib, 4ull * n);
  }
};

void from_buffer(Source& s, sb_proof* pr) {
  const sb_proof_layout& l = pr->layout;
  u64* W = pr->words;
  s.u64s(W + l.off_trace_cap, 4ull * l.cap_len);
  s.u64s(W + l.off_quotient_cap, 4ull * l.cap_len);
  s.u64s(W + l.off_local_values, 2ull * l.n_cols);
  s.u64s(W + l.off_next_values, 2ull * l.n_cols);
  s.u64s(W + l.off_quotient_polys, 2ull * l.n_quotient_polys);
  s.u64s(W + l.off_fri_caps, 4ull * l.cap_len * l.n_fri_rounds);
  for (uint32_t q = 0; q < l.n_queries; q++) {
    u64* Q = W + l.off_queries + (uint64_t)q * l.query_stride;
    s.u64s(Q + l.q_off_trace_leaf, l.n_cols);
    s.merkle_proof(Q + l.q_off_trace_path, l.trace_path_len);
    s.u64s(Q + l.q_off_quot_leaf, l.n_quotient_polys);
    s.merkle_proof(Q + l.q_off_quot_path, l.trace_path_len);
    for (uint32_t r = 0; r < l.n_fri_rounds; r++) {
      u64* S = Q + fri_step_offset(l, r);
      s.u64s(S, 2ull << l.arity_bits);
      s.merkle_proof(S + (2ull << l.arity_bits), fri_step_path_len(l, r));
    }
  }
  s.u64s(W + l.off_final_poly, 2ull * l.final_poly_len);
  s.u64s(W + l.off_pow_witness, 1);
  s.u64s(W + l.off_public_inputs, l.n_public_inputs);
  if (s.pos != s.len) SB_THROW(SB_EINVAL, "%zu trailing bytes after the serialized proof", s.len - s.pos);
}

// ---- serde_json shapes ----
void j_fields(Sink& s, const u64* w, size_t n) {
  s.str("[");
  for (size_t i = 0; i < n; i++) { if (i) s.str(","); s.num(w[i]); }
  s.str("]");
}
void j_exts(Sink& s, const u64* w, size_t n) {
  s.str("[");
  for (size_t i = 0; i < n; i++) { if (i) s.str(","); s.str("["); s.num(w[2 * i]); s.str(","); s.num(w[2 * i + 1]); s.str("]"); }
  s.str("]");
}
void j_hashes(Sink& s, const u64* w, size_t n) {
  s.str("[");
  for (size_t i = 0; i < n; i++) { if (i) s.str(","); s.str("{\"elements\":"); j_fields(s, w + 4 * i, 4); s.str("}"); }
  s.str("]");
}
void j_merkle_proof(Sink& s, const u64* sib, uint32_t len) { s.str("{\"siblings\":"); j_hashes(s, sib, len); s.str("}"); }

void to_json(Sink& s, const sb_proof* pr, const sb_params* p) {
  const sb_proof_layout& l = pr->layout;
  const u64* W = pr->words;
  static const char* names[] = {"FP12MulStark", "PairingPrecompStark", "MillerLoopStark", "FinalExponentiateStark", "ECCAggStark"};
  s.str("{\"stark\":\"");
  s.str(p && p->stark_id <= SB_STARK_ECC_AGG ? names[p->stark_id] : "custom");
  s.str("\",\"degree_bits\":"); s.num(l.log_n);
  if (p) {
    s.str(",\"config\":{\"security_bits\":100,\"num_challenges\":"); s.num(p->num_challenges);
    s.str(",\"fri_config\":{\"rate_bits\":"); s.num(p->rate_bits);
    s.str(",\"cap_height\":"); s.num(p->cap_height);
    s.str(",\"proof_of_work_bits\":"); s.num(p->pow_bits);
    s.str(",\"reduction_strategy\":{\"ConstantArityBits\":["); s.num(p->fri_arity_bits); s.str(","); s.num(p->fri_final_poly_bits);
    s.str("]},\"num_query_rounds\":"); s.num(p->num_query_rounds); s.str("}}");
  }
  s.str(",\"proof\":{\"trace_cap\":"); j_hashes(s, W + l.off_trace_cap, l.cap_len);
  s.str(",\"permutation_zs_cap\":null,\"quotient_polys_cap\":"); j_hashes(s, W + l.off_quotient_cap, l.cap_len);
  s.str(",\"openings\":{\"local_values\":"); j_exts(s, W + l.off_local_values, l.n_cols);
  s.str(",\"next_values\":"); j_exts(s, W + l.off_next_values, l.n_cols);
  s.str(",\"permutation_zs\":null,\"permutation_zs_next\":null,\"quotient_polys\":"); j_exts(s, W + l.off_quotient_polys, l.n_quotient_polys);
  s.str("},\"opening_proof\":{\"commit_phase_merkle_caps\":[");
  for (uint32_t r = 0; r < l.n_fri_rounds; r++) { if (r) s.str(","); j_hashes(s, W + l.off_fri_caps + 4ull * l.cap_len * r, l.cap_len); }
  s.str("],\"query_round_proofs\":[");
  for (uint32_t q = 0; q < l.n_queries; q++) {
    const u64* Q = W + l.off_queries + (uint64_t)q * l.query_stride;
    if (q) s.str(",");
    s.str("{\"initial_trees_proof\":{\"evals_proofs\":[[");
    j_fields(s, Q + l.q_off_trace_leaf, l.n_cols); s.str(","); j_merkle_proof(s, Q + l.q_off_trace_path, l.trace_path_len);
    s.str("],[");
    j_fields(s, Q + l.q_off_quot_leaf, l.n_quotient_polys); s.str(","); j_merkle_proof(s, Q + l.q_off_quot_path, l.trace_path_len);
    s.str("]]},\"steps\":[");
    for (uint32_t r = 0; r < l.n_fri_rounds; r++) {
      const u64* S = Q + fri_step_offset(l, r);
      if (r) s.str(",");
      s.str("{\"evals\":"); j_exts(s, S, 1ull << l.arity_bits);
      s.str(",\"merkle_proof\":"); j_merkle_proof(s, S + (2ull << l.arity_bits), fri_step_path_len(l, r));
      s.str("}");
    }
    s.str("]}");
  }
  s.str("],\"final_poly\":{\"coeffs\":"); j_exts(s, W + l.off_final_poly, l.final_poly_len);
  s.str("},\"pow_witness\":"); s.num(W[l.off_pow_witness]);
  s.str("}},\"public_inputs\":"); j_fields(s, W + l.off_public_inputs, l.n_public_inputs);
  s.str("}");
}

bool same_layout(const sb_proof_layout& a, const sb_proof_layout& b) { return memcmp(&a, &b, sizeof(a)) == 0; }

}  // namespace

extern "C" {

int sb_proof_serialize(const sb_proof* proof, const sb_params* p, int format, void* buf, size_t cap, size_t* len_out) {
  if (!proof || !proof->words || !len_out) return SB_EINVAL;
  try {
    if (p && !same_layout(proof_layout(*p), proof->layout)) SB_THROW(SB_EINVAL, "params do not describe this proof's layout");
    Sink s{(unsigned char*)buf, buf ? cap : 0};
    switch (format) {
      case SB_WIRE_POD: {
        if (!p) SB_THROW(SB_EINVAL, "SB_WIRE_POD needs the params that define the layout");
        s.put("SBPROOF1", 8);
        s.put(p, sizeof(*p));
        const u64 n = proof->layout.total_words;
        s.put(&n, 8);
        s.u64s(proof->words, n);
        break;
      }
      case SB_WIRE_PLONKY2_BUFFER: to_buffer(s, proof); break;
      case SB_WIRE_SERDE_JSON: to_json(s, proof, p); break;
      default: SB_THROW(SB_EINVAL, "unknown wire format %d", format);
    }
    *len_out = s.len;
    if (s.overflow) SB_THROW(SB_EINVAL, "buffer of %zu bytes is too small for %zu", cap, s.len);
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(nullptr, e); }
}

int sb_proof_deserialize(const void* buf, size_t len, int format, const sb_params* p, sb_params* params_out, sb_proof** out) {
  if (!buf || !out) return SB_EINVAL;
  sb_proof* pr = nullptr;
  try {
    Source s{(const unsigned char*)buf, len};
    sb_params q;
    if (format == SB_WIRE_POD) {
      char magic[8];
      s.get(magic, 8);
      if (memcmp(magic, "SBPROOF1", 8)) SB_THROW(SB_EINVAL, "not an SBPROOF1 image");
      s.get(&q, sizeof(q));
      check_params(&q);
      if (p && memcmp(p, &q, offsetof(sb_params, flags)) != 0) SB_THROW(SB_EINVAL, "the image was made with other parameters");
      u64 n;
      s.get(&n, 8);
      pr = proof_alloc(q);
      if (n != pr->layout.total_words) SB_THROW(SB_EINVAL, "image holds %" PRIu64 " words, the layout has %" PRIu64, n, (u64)pr->layout.total_words);
      s.u64s(pr->words, n);
      if (s.pos != s.len) SB_THROW(SB_EINVAL, "%zu trailing bytes after the serialized proof", s.len - s.pos);
    } else if (format == SB_WIRE_PLONKY2_BUFFER) {
      if (!p) SB_THROW(SB_EINVAL, "SB_WIRE_PLONKY2_BUFFER carries no lengths: params are required to read it");
      q = *p;
      check_params(&q);
      pr = proof_alloc(q);
      from_buffer(s, pr);
    } else SB_THROW(SB_EINVAL, "wire format %d cannot be read back", format);
    if (params_out) *params_out = q;
    *out = pr;
    return SB_OK;
  } catch (const SbError& e) { sb_proof_free(pr); return sb_fail(nullptr, e); }
}

// A proof object from raw POD words (e.g. produced by another prover with the same layout): copies `n_words` words.
int sb_proof_from_words(const sb_params* p, const uint64_t* words, size_t n_words, sb_proof** out) {
  if (!p || !words || !out) return SB_EINVAL;
  sb_proof* pr = nullptr;
  try {
    check_params(p);
    pr = proof_alloc(*p);
    if (n_words != pr->layout.total_words) SB_THROW(SB_EINVAL, "%zu words given, the layout has %" PRIu64, n_words, (u64)pr->layout.total_words);
    memcpy(pr->words, words, 8 * n_words);
    *out = pr;
    return SB_OK;
  } catch (const SbError& e) { sb_proof_free(pr); return sb_fail(nullptr, e); }
}

}  // extern "C"
