// Host orchestration of one proof: the device replacement of starky::prover::prove
// (/root/reference/src/aggregate_proof.rs:59,105,138,169,212).
#include "prover.cuh"

std::vector<unsigned> fri_arities(const sb_params& p) {
  // FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits)
  std::vector<unsigned> r;
  unsigned db = p.log_n;
  while (db > p.fri_final_poly_bits && db + p.rate_bits - p.fri_arity_bits >= p.cap_height) {
    r.push_back(p.fri_arity_bits);
    db -= p.fri_arity_bits;
  }
  return r;
}

uint32_t fri_step_path_len(const sb_proof_layout& l, uint32_t round) {
  unsigned log_leaves = l.log_lde - l.arity_bits * (round + 1);
  return log_leaves - ilog2(l.cap_len);
}
uint64_t fri_step_offset(const sb_proof_layout& l, uint32_t round) {
  uint64_t o = l.q_off_steps;
  for (uint32_t r = 0; r < round; r++) o += (uint64_t(2) << l.arity_bits) + 4ull * fri_step_path_len(l, r);
  return o;
}

sb_proof_layout proof_layout(const sb_params& p) {
  if (p.fri_arity_bits == 0) SB_THROW(SB_EINVAL, "fri_arity_bits is 0");
  sb_proof_layout l = {};
  std::vector<unsigned> ar = fri_arities(p);
  l.log_n = p.log_n; l.log_lde = p.log_n + p.rate_bits; l.n_cols = p.n_cols;
  l.n_quotient_polys = p.num_challenges * quotient_degree_factor(p);
  l.n_public_inputs = p.n_public_inputs; l.cap_len = 1u << p.cap_height;
  l.n_fri_rounds = (uint32_t)ar.size();
  l.final_poly_len = 1u << (p.log_n - p.fri_arity_bits * (uint32_t)ar.size());
  l.n_queries = p.num_query_rounds; l.arity_bits = p.fri_arity_bits; l.trace_path_len = l.log_lde - p.cap_height;
  uint64_t o = 0;
  l.off_trace_cap = o; o += 4ull * l.cap_len;
  l.off_quotient_cap = o; o += 4ull * l.cap_len;
  l.off_local_values = o; o += 2ull * l.n_cols;
  l.off_next_values = o; o += 2ull * l.n_cols;
  l.off_quotient_polys = o; o += 2ull * l.n_quotient_polys;
  l.off_fri_caps = o; o += 4ull * l.cap_len * l.n_fri_rounds;
  l.off_final_poly = o; o += 2ull * l.final_poly_len;
  l.off_pow_witness = o; o += 1;
  l.off_queries = o;
  uint64_t q = 0;
  l.q_off_trace_leaf = q; q += l.n_cols;
  l.q_off_trace_path = q; q += 4ull * l.trace_path_len;
  l.q_off_quot_leaf = q; q += l.n_quotient_polys;
  l.q_off_quot_path = q; q += 4ull * l.trace_path_len;
  l.q_off_steps = q;
  for (uint32_t r = 0; r < l.n_fri_rounds; r++) q += (uint64_t(2) << l.arity_bits) + 4ull * fri_step_path_len(l, r);
  l.query_stride = q;
  o += q * l.n_queries;
  l.off_public_inputs = o; o += l.n_public_inputs;
  l.total_words = o;
  return l;
}

// ---- temporary stubs (filled in by quotient.cu / fri.cu as those stages land) ----
void air_release_all(sb_ctx*) {}

__global__ void bitrev_permute_kernel(const u64* __restrict__ in, u64* __restrict__ out, unsigned log_size, uint64_t total) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  uint64_t vec = i >> log_size;
  uint32_t k = (uint32_t)(i & ((1ull << log_size) - 1));
  out[(vec << log_size) + bitrev32(k, log_size)] = in[i];
}
void sb_bitrev_permute_device(sb_ctx* ctx, const u64* d_in, u64* d_out, unsigned log_size, uint32_t count) {
  uint64_t total = (uint64_t)count << log_size;
  LAUNCH(ctx, bitrev_permute_kernel, (unsigned)((total + 255) / 256), 256, 0, d_in, d_out, log_size, total);
}

extern "C" {
int sb_air_load(sb_ctx* ctx, uint32_t, const char*) { return sb_fail(ctx, SbError{SB_EAIR, "sb_air_load: not built yet"}); }
int sb_prove(sb_ctx* ctx, const sb_params*, const void*, int, const uint64_t*, sb_proof**) { return sb_fail(ctx, SbError{SB_EINVAL, "sb_prove: not built yet"}); }
void sb_proof_free(sb_proof* p) { if (p) { delete[] p->words; delete p; } }
int sb_quotient_values(sb_ctx* ctx, const sb_params*, const uint64_t*, const uint64_t*, uint64_t*) { return sb_fail(ctx, SbError{SB_EINVAL, "sb_quotient_values: not built yet"}); }
}
