// Host orchestration of one proof: the device replacement of starky::prover::prove
// (/root/reference/src/aggregate_proof.rs:59,105,138,169,212).  The Fiat-Shamir transcript
// (plonky2 Challenger, ~10^2 permutations) runs on the host between kernels; only caps, openings, the two combined
// polynomials and the final proof cross PCIe.
#include <functional>
#include <memory>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <map>
#include <set>

#include "poseidon.cuh"
#include "prover.cuh"

// ---- stage functions from the other translation units ----
struct AirProgram;
AirProgram* air_get(sb_ctx* ctx, const sb_params* p);
void sb_quotient_device(sb_ctx* ctx, const sb_params* p, const u64* d_pis, const u64* alphas, u64* d_out);
void sb_openings_device(sb_ctx* ctx, const u64* d_coeffs, unsigned log_n, uint32_t n_polys, e2_t za, const e2_t* zb,
                        e2_t* d_tab_a, e2_t* d_tab_b, e2_t* d_out_a, e2_t* d_out_b);
void sb_combine_device(sb_ctx* ctx, const u64* d_coeffs, unsigned log_n, uint32_t n_polys, e2_t alpha, uint32_t j0,
                       e2_t* d_apow, e2_t* d_partial, size_t partial_capacity_elems, e2_t* d_out);
void sb_fri_leaf_hash_device(sb_ctx* ctx, const u64* d_re, const u64* d_im, uint32_t n_leaves, unsigned arity_bits, u64* d_digests);
u64 sb_pow_device(sb_ctx* ctx, const u64 state[12], int pos, unsigned bits, unsigned long long* d_best);
void sb_gather_leaf_device(sb_ctx* ctx, u64* dst, uint64_t stride, uint64_t off, const u64* cols, uint32_t N, uint32_t n_cols,
                           const uint32_t* d_positions, uint32_t n_queries);
void sb_gather_path_device(sb_ctx* ctx, u64* dst, uint64_t stride, uint64_t off, const u64* tree, uint32_t n_leaves,
                           uint32_t path_len, const uint32_t* d_leaf_idx, unsigned shift, uint32_t n_queries);
void sb_gather_fri_evals_device(sb_ctx* ctx, u64* dst, uint64_t stride, uint64_t off, const u64* re, const u64* im,
                                unsigned arity_bits, const uint32_t* d_leaf_idx, unsigned shift, uint32_t n_queries);
void sb_coset_to_bitrev_device(sb_ctx* ctx, const u64* d_in, u64* d_out, unsigned log_n, unsigned log_N, uint32_t count);
void sb_coset_shift_device(sb_ctx* ctx, u64* d, unsigned log_size, uint32_t count, u64 s);
void sb_tail_nonzero_device(sb_ctx* ctx, const u64* d, unsigned log_size, uint32_t count, uint32_t keep, int* d_flag);

// ---------------------------------------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------------------------------------
std::vector<unsigned> fri_arities(const sb_params& p) {
  // FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits)
  // (plonky2 asserts degree_bits >= arity_bits inside the loop and would underflow in usize otherwise: SB_EINVAL here)
  std::vector<unsigned> r;
  unsigned db = p.log_n;
  if (p.fri_arity_bits == 0) SB_THROW(SB_EINVAL, "fri_arity_bits is 0");
  while (db > p.fri_final_poly_bits) {
    // `degree_bits + rate_bits - arity_bits >= cap_height` in usize: a negative difference panics (debug) or wraps to "true"
    // and then trips the assert (release) -- an error either way, never "stop here"
    if (db + p.rate_bits >= p.fri_arity_bits && db + p.rate_bits - p.fri_arity_bits < p.cap_height) break;
    if (db < p.fri_arity_bits)
      SB_THROW(SB_EINVAL, "FRI reduction: degree_bits %u < arity_bits %u (log_n %u, final_poly_bits %u, cap_height %u)", db,
               p.fri_arity_bits, p.log_n, p.fri_final_poly_bits, p.cap_height);
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
  check_params(&p);
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

// ---------------------------------------------------------------------------------------------------------
// plonky2 Challenger (duplex sponge over Poseidon-12), host side
// ---------------------------------------------------------------------------------------------------------
void sb_host_poseidon_permute(u64 s[12]);   // host_poseidon.cpp

struct HostChallenger {
  u64 state[12];
  u64 in_buf[8]; int n_in = 0;
  u64 out_buf[8]; int n_out = 0;
  HostChallenger() { memset(state, 0, sizeof(state)); }
  void duplexing() {
    for (int i = 0; i < n_in; i++) state[i] = in_buf[i];
    n_in = 0;
    sb_host_poseidon_permute(state);
    memcpy(out_buf, state, 64);
    n_out = 8;
  }
  void observe(u64 x) { n_out = 0; in_buf[n_in++] = x; if (n_in == 8) duplexing(); }
  void observe_many(const u64* x, size_t n) { for (size_t i = 0; i < n; i++) observe(x[i]); }
  u64 challenge() { if (n_in > 0 || n_out == 0) duplexing(); return out_buf[--n_out]; }
  e2_t ext_challenge() { u64 a = challenge(); u64 b = challenge(); return e2_make(a, b); }
};

// The first step of the transcript, for callers that run the stages themselves (the sharded path: every rank derives
// the same alphas from the gathered cap): observe the trace cap, draw `num_challenges` challenges.
extern "C" int sb_transcript_alphas(const uint64_t* trace_cap, uint32_t cap_len, uint32_t num_challenges, uint64_t* alphas_out) {
  if (!trace_cap || !alphas_out || !cap_len || num_challenges > 8) return SB_EINVAL;
  HostChallenger ch;
  ch.observe_many(trace_cap, 4ull * cap_len);
  for (uint32_t j = 0; j < num_challenges; j++) alphas_out[j] = ch.challenge();
  return SB_OK;
}

struct Arena {
  char* base; size_t off = 0, cap;
  Arena(void* b, size_t c) : base((char*)b), cap(c) {}
  template <class T> T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* r = (T*)(base + off);
    off += sizeof(T) * count;
    if (off > cap) SB_THROW(SB_ENOMEM, "internal: work arena overflow");
    return r;
  }
};

static float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

// fri_committed_trees (plonky2 fri::prover): per round the coefficients go to values on the current coset
// (shift^m scaling on the host, zero padding to the LDE size, forward transform on the device, bit-reversed order), the
// arity-sized leaves are hashed and the tree built to the cap; `next_beta(round, cap)` supplies the folding challenge
// (the transcript in a proof, injected values in the sb_fri_commit stage export) and the polynomial is folded in
// coefficient form.  caps_out: host, [rounds][cap_len][4].
struct FriRound { u64* re; u64* im; u64* tree; uint32_t n_leaves; unsigned log_size; };
template <class BetaFn>
static void fri_commit_phase(sb_ctx* ctx, const sb_params* p, Arena& ar, std::vector<e2_t>& coeffs, u64* caps_out,
                             std::vector<FriRound>& rounds, BetaFn&& next_beta) {
  const std::vector<unsigned> arities = fri_arities(*p);
  const uint32_t cap_len = 1u << p->cap_height;
  cudaStream_t st = ctx->stream;
  u64 shift = 7;
  unsigned cur_log = p->log_n + p->rate_bits;
  for (size_t round = 0; round < arities.size(); round++) {
    const unsigned ab = arities[round];
    const size_t size = size_t(1) << cur_log;
    u64* d_vals = ar.take<u64>(2 * size);
    std::vector<u64> host(2 * size, 0);
    u64 s = 1;
    for (size_t m = 0; m < coeffs.size(); m++) {
      host[m] = gl_mul(coeffs[m].a, s);
      host[size + m] = gl_mul(coeffs[m].b, s);
      s = gl_mul(s, shift);
    }
    CUDA_CHECK(cudaMemcpyAsync(d_vals, host.data(), 16ull * size, cudaMemcpyHostToDevice, st));
    sb_ntt_device(ctx, d_vals, cur_log, 2, false, /*dif=*/true);
    FriRound R;
    R.re = d_vals; R.im = d_vals + size; R.log_size = cur_log; R.n_leaves = (uint32_t)(size >> ab);
    R.tree = ar.take<u64>(8ull * R.n_leaves + 64);
    sb_fri_leaf_hash_device(ctx, R.re, R.im, R.n_leaves, ab, R.tree);
    sb_merkle_levels(ctx, R.tree, R.n_leaves, p->cap_height);
    u64* cap_dst = caps_out + round * 4ull * cap_len;
    CUDA_CHECK(cudaMemcpyAsync(cap_dst, tree_cap_ptr(R.tree, R.n_leaves, p->cap_height), 32ull * cap_len, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));   // also keeps `host` alive until the upload is done
    rounds.push_back(R);
    const e2_t beta = next_beta(round, (const u64*)cap_dst);
    const size_t arity = size_t(1) << ab;
    std::vector<e2_t> folded(coeffs.size() / arity);
    for (size_t m = 0; m < folded.size(); m++) {
      e2_t acc = e2_make(0, 0);
      for (size_t i = arity; i-- > 0;) acc = e2_add(e2_mul(acc, beta), coeffs[arity * m + i]);
      folded[m] = acc;
    }
    coeffs.swap(folded);
    shift = gl_pow(shift, arity);
    cur_log -= ab;
  }
}

#define HOOK(call)                                                                              \
  do {                                                                                          \
    int rc_ = (call);                                                                           \
    if (rc_) SB_THROW(rc_, "sharded proof: hook %s failed with code %d", #call, rc_);           \
  } while (0)

// hooks == nullptr: one GPU holds everything.  hooks != nullptr: the trace is sharded over the GPUs of a box and the five
// distributed steps are done by the host (include/starky_b200.h: sb_shard_hooks); every rank runs this same function.
static void prove_impl(sb_ctx* ctx, const sb_params* p, const void* trace, int layout, const uint64_t* public_inputs,
                       sb_proof* proof, const sb_shard_hooks* hooks = nullptr) {
  const unsigned log_n = p->log_n, r = p->rate_bits, log_N = log_n + r, qdf = quotient_degree_factor(*p);
  const uint32_t n = 1u << log_n, N = 1u << log_N, C = p->n_cols;
  const sb_proof_layout L = proof_layout(*p);
  const uint32_t nq = L.n_quotient_polys;
  if (p->n_public_inputs && !public_inputs) SB_THROW(SB_EINVAL, "public_inputs is NULL");
  if (p->num_challenges != 2) SB_THROW(SB_EINVAL, "num_challenges must be 2");
  air_get(ctx, p);  // validates the shape against the constraint program before any work
  u64* W = proof->words;
  cudaStream_t st = ctx->stream;
  cudaEvent_t ev[8];
  for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
  struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 8; i++) cudaEventDestroy(e[i]); } } guard{ev};

  // ---- 0. ingest, 1. trace commitment ----
  CUDA_CHECK(cudaEventRecord(ev[0], st));
  if (hooks) {
    CUDA_CHECK(cudaEventRecord(ev[1], st));
    HOOK(hooks->commit(hooks->user, W + L.off_trace_cap));
  } else {
    ingest_and_commit_trace(ctx, p, trace, layout, ev[1]);   // ev[1] = trace fully on the device (copy stream)
    CUDA_CHECK(cudaMemcpyAsync(W + L.off_trace_cap, tree_cap_ptr(ctx->tree.as<u64>(), N, p->cap_height), 32ull * L.cap_len,
                               cudaMemcpyDeviceToHost, st));
  }
  ctx->pis.ensure(8ull * (p->n_public_inputs + 1));
  if (p->n_public_inputs)
    CUDA_CHECK(cudaMemcpyAsync(ctx->pis.p, public_inputs, 8ull * p->n_public_inputs, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaEventRecord(ev[2], st));
  CUDA_CHECK(cudaStreamSynchronize(st));

  HostChallenger ch;
  if (p->flags & SB_FLAG_OBSERVE_PUBLIC_INPUTS) ch.observe_many(public_inputs, p->n_public_inputs);
  ch.observe_many(W + L.off_trace_cap, 4ull * L.cap_len);
  u64 alphas[2] = {ch.challenge(), ch.challenge()};

  // ---- work arena ----
  const size_t partial_elems = (size_t)(ctx->sm_count * 16 + 8) * 128 + 2 * (size_t)n;
  size_t need = 16ull * N * 2 + 8ull * nq * n * 2 + 16ull * n * 2 + 16ull * (2ull * C + nq + 8) + 16ull * (C + nq + 8) +
                16ull * partial_elems + 16ull * n * 2 + 8ull * L.query_stride * L.n_queries + 16ull * N * 2 + 64ull * N +
                (hooks ? 8ull * L.n_queries * C + 4096 : 0) +
                (1 << 16);
  ctx->scratch2.ensure(need);
  Arena ar(ctx->scratch2.p, ctx->scratch2.cap);

  // ---- 3. quotient values, K5: quotient polynomials and their commitment ----
  ctx->qvals.ensure(16ull * N);
  u64* d_q = ctx->qvals.as<u64>();
  if (hooks) {
    HOOK(hooks->quotient(hooks->user, alphas, d_q));
  } else {
    stage_begin(ctx, "quotient");
    sb_quotient_device(ctx, p, ctx->pis.as<u64>(), alphas, d_q);
    stage_end(ctx, "quotient");
  }
  CUDA_CHECK(cudaEventRecord(ev[3], st));
  u64* d_qtmp = ar.take<u64>(2ull * N);
  int* d_flag = ar.take<int>(4);
  sb_coset_to_bitrev_device(ctx, d_q, d_qtmp, log_n, log_N, 2);          // values in bit-reversed natural order
  sb_ntt_device(ctx, d_qtmp, log_N, 2, /*inverse=*/true, /*dif=*/false);  // -> natural-order coefficients of q(7X)
  sb_coset_shift_device(ctx, d_qtmp, log_N, 2, gl_inv(7));                // coset_ifft(7)
  CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
  sb_tail_nonzero_device(ctx, d_qtmp, log_N, 2, qdf * n, d_flag);
  int h_flag = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  // chunk polys [nq][n], natural coefficient order -> values on H -> the same LDE + Merkle path as the trace
  u64* d_qc_nat = ar.take<u64>((size_t)nq * n);
  u64* d_qc_br = ar.take<u64>((size_t)nq * n);
  for (unsigned j = 0; j < 2; j++)
    CUDA_CHECK(cudaMemcpyAsync(d_qc_nat + (size_t)j * qdf * n, d_qtmp + (size_t)j * N, 8ull * qdf * n, cudaMemcpyDeviceToDevice, st));
  sb_bitrev_permute_device(ctx, d_qc_nat, d_qc_br, log_n, nq);
  sb_ntt_device(ctx, d_qc_br, log_n, nq, false, /*dif=*/false);          // bit-reversed coefficients -> natural values
  ctx->qcoeffs.ensure(8ull * nq * n);
  ctx->qlde.ensure(8ull * nq * N);
  ctx->qtree.ensure(64ull * N);
  sb_lde_trace(ctx, d_qc_br, ctx->qcoeffs.as<u64>(), ctx->qlde.as<u64>(), nq, log_n, r);
  sb_hash_leaves_device(ctx, ctx->qlde.as<u64>(), nq, N, log_n, ctx->qtree.as<u64>());
  sb_merkle_levels(ctx, ctx->qtree.as<u64>(), N, p->cap_height);
  CUDA_CHECK(cudaMemcpyAsync(W + L.off_quotient_cap, tree_cap_ptr(ctx->qtree.as<u64>(), N, p->cap_height), 32ull * L.cap_len,
                             cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaEventRecord(ev[4], st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (h_flag && !(p->flags & SB_FLAG_ALLOW_INVALID_TRACE))
    SB_THROW(SB_EQUOTIENT_NOT_DIVISIBLE, "Quotient has failed, the vanishing polynomial is not divisible by Z_H");
  ch.observe_many(W + L.off_quotient_cap, 4ull * L.cap_len);

  // ---- 4. zeta, 5. openings ----
  const e2_t zeta = ch.ext_challenge();
  {
    e2_t t = zeta;
    for (unsigned i = 0; i < log_n; i++) t = e2_mul(t, t);
    if (e2_eq(t, e2_make(1, 0))) SB_THROW(SB_EZETA_IN_SUBGROUP, "Opening point is in the subgroup.");
  }
  const u64 g = gl_root(log_n);
  const e2_t zeta_next = e2_scale(zeta, g);
  e2_t* d_tab_a = ar.take<e2_t>(n);
  e2_t* d_tab_b = ar.take<e2_t>(n);
  e2_t* d_open = ar.take<e2_t>(2ull * C + nq);
  if (hooks) {
    const u64 z[2] = {zeta.a, zeta.b}, zn[2] = {zeta_next.a, zeta_next.b};
    HOOK(hooks->openings(hooks->user, z, zn, W + L.off_local_values, W + L.off_next_values));
  } else {
    sb_openings_device(ctx, ctx->coeffs.as<u64>(), log_n, C, zeta, &zeta_next, d_tab_a, d_tab_b, d_open, d_open + C);
    CUDA_CHECK(cudaMemcpyAsync(W + L.off_local_values, d_open, 16ull * C, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(W + L.off_next_values, d_open + C, 16ull * C, cudaMemcpyDeviceToHost, st));
  }
  sb_openings_device(ctx, ctx->qcoeffs.as<u64>(), log_n, nq, zeta, nullptr, d_tab_a, d_tab_b, d_open + 2ull * C, nullptr);
  CUDA_CHECK(cudaMemcpyAsync(W + L.off_quotient_polys, d_open + 2ull * C, 16ull * nq, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaEventRecord(ev[5], st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  ch.observe_many(W + L.off_local_values, 2ull * C);        // batch 0 = local_values ++ quotient_polys
  ch.observe_many(W + L.off_quotient_polys, 2ull * nq);
  ch.observe_many(W + L.off_next_values, 2ull * C);         // batch 1 = next_values

  // ---- 6. prove_openings: batch combine on the device, the two synthetic divisions on the host ----
  const e2_t alpha = ch.ext_challenge();
  e2_t* d_apow = ar.take<e2_t>((size_t)C + nq);
  e2_t* d_partial = ar.take<e2_t>(partial_elems);
  e2_t* d_F = ar.take<e2_t>(2ull * n);
  if (hooks) {
    const u64 al[2] = {alpha.a, alpha.b};
    HOOK(hooks->combine(hooks->user, al, (u64*)d_F));
  } else {
    sb_combine_device(ctx, ctx->coeffs.as<u64>(), log_n, C, alpha, 0, d_apow, d_partial, partial_elems, d_F);
  }
  sb_combine_device(ctx, ctx->qcoeffs.as<u64>(), log_n, nq, alpha, C, d_apow, d_partial, partial_elems, d_F + n);
  std::vector<e2_t> hF(2ull * n);
  CUDA_CHECK(cudaMemcpyAsync(hF.data(), d_F, 32ull * n, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<e2_t> coeffs(n);
  {
    std::vector<e2_t> ft(n), fz(n);
    for (uint32_t k = 0; k < n; k++) {                       // bit-reversed coefficient order -> natural
      uint32_t pp = bitrev32(k, log_n);
      ft[k] = hF[pp];
      fz[k] = e2_add(hF[pp], hF[n + pp]);
    }
    // Q_z = (F - F(z)) / (X - z), padded back to n coefficients; final = alpha^C * Q_zeta + Q_{g zeta}
    const e2_t shift = e2_pow(alpha, C);
    e2_t a0 = e2_make(0, 0), a1 = e2_make(0, 0);
    coeffs[n - 1] = e2_make(0, 0);
    for (uint32_t k = n; k-- > 1;) {
      a0 = e2_add(e2_mul(a0, zeta), fz[k]);
      a1 = e2_add(e2_mul(a1, zeta_next), ft[k]);
      coeffs[k - 1] = e2_add(e2_mul(a0, shift), a1);
    }
    if (p->flags & SB_FLAG_FRI_MUL_BY_X) { coeffs.insert(coeffs.begin(), e2_make(0, 0)); coeffs.pop_back(); }
  }

  // ---- FRI commit phase ----
  const std::vector<unsigned> arities = fri_arities(*p);
  std::vector<FriRound> rounds;
  fri_commit_phase(ctx, p, ar, coeffs, W + L.off_fri_caps, rounds, [&](size_t, const u64* cap) {
    ch.observe_many(cap, 4ull * L.cap_len);
    return ch.ext_challenge();
  });
  if (coeffs.size() != L.final_poly_len) SB_THROW(SB_EINVAL, "internal: final polynomial length %zu != %u", coeffs.size(), L.final_poly_len);
  for (size_t i = 0; i < coeffs.size(); i++) { W[L.off_final_poly + 2 * i] = coeffs[i].a; W[L.off_final_poly + 2 * i + 1] = coeffs[i].b; }
  ch.observe_many(W + L.off_final_poly, 2ull * coeffs.size());

  // ---- proof of work ----
  u64 witness;
  if (p->flags & SB_FLAG_FIXED_POW_WITNESS) witness = p->fixed_pow_witness;
  else {
    u64 inter[12];
    memcpy(inter, ch.state, sizeof(inter));
    for (int i = 0; i < ch.n_in; i++) inter[i] = ch.in_buf[i];
    witness = sb_pow_device(ctx, inter, ch.n_in, p->pow_bits, ar.take<unsigned long long>(2));
  }
  ch.observe(witness);
  const u64 response = ch.challenge();
  if (p->pow_bits && (response >> (64 - p->pow_bits)) != 0) SB_THROW(SB_EPOW, "proof-of-work witness does not satisfy %u bits", p->pow_bits);
  W[L.off_pow_witness] = witness;

  // ---- query rounds ----
  const uint32_t nQ = L.n_queries;
  std::vector<uint32_t> h_idx(2ull * nQ);
  for (uint32_t q = 0; q < nQ; q++) {
    uint32_t x = (uint32_t)(ch.challenge() % N);
    h_idx[q] = x;                                                             // plonky2 leaf index
    h_idx[nQ + q] = (x & ~(n - 1)) | bitrev32(x & (n - 1), log_n);            // device LDE position of that leaf
  }
  uint32_t* d_idx = ar.take<uint32_t>(2ull * nQ);
  u64* d_queries = ar.take<u64>(L.query_stride * nQ);
  CUDA_CHECK(cudaMemcpyAsync(d_idx, h_idx.data(), 8ull * nQ, cudaMemcpyHostToDevice, st));
  if (hooks) {
    u64* d_rows = ar.take<u64>((size_t)nQ * C);
    HOOK(hooks->query_rows(hooks->user, h_idx.data() + nQ, nQ, d_rows));
    CUDA_CHECK(cudaMemcpy2DAsync(d_queries + L.q_off_trace_leaf, 8ull * L.query_stride, d_rows, 8ull * C, 8ull * C, nQ,
                                 cudaMemcpyDeviceToDevice, st));
  } else {
    sb_gather_leaf_device(ctx, d_queries, L.query_stride, L.q_off_trace_leaf, ctx->lde.as<u64>(), N, C, d_idx + nQ, nQ);
  }
  sb_gather_path_device(ctx, d_queries, L.query_stride, L.q_off_trace_path, ctx->tree.as<u64>(), N, L.trace_path_len, d_idx, 0, nQ);
  sb_gather_leaf_device(ctx, d_queries, L.query_stride, L.q_off_quot_leaf, ctx->qlde.as<u64>(), N, nq, d_idx + nQ, nQ);
  sb_gather_path_device(ctx, d_queries, L.query_stride, L.q_off_quot_path, ctx->qtree.as<u64>(), N, L.trace_path_len, d_idx, 0, nQ);
  unsigned sh = 0;
  for (size_t round = 0; round < rounds.size(); round++) {
    const unsigned ab = arities[round];
    sh += ab;
    const uint64_t off = fri_step_offset(L, (uint32_t)round);
    sb_gather_fri_evals_device(ctx, d_queries, L.query_stride, off, rounds[round].re, rounds[round].im, ab, d_idx, sh, nQ);
    sb_gather_path_device(ctx, d_queries, L.query_stride, off + (uint64_t(2) << ab), rounds[round].tree, rounds[round].n_leaves,
                          fri_step_path_len(L, (uint32_t)round), d_idx, sh, nQ);
  }
  CUDA_CHECK(cudaEventRecord(ev[6], st));
  CUDA_CHECK(cudaMemcpyAsync(W + L.off_queries, d_queries, 8ull * L.query_stride * nQ, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaEventRecord(ev[7], st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (p->n_public_inputs) memcpy(W + L.off_public_inputs, public_inputs, 8ull * p->n_public_inputs);

  proof->ms_h2d = ev_ms(ev[0], ev[1]);
  proof->ms_trace_commit = ev_ms(ev[1], ev[2]);
  proof->ms_quotient = ev_ms(ev[2], ev[3]);
  proof->ms_quotient_commit = ev_ms(ev[3], ev[4]);
  proof->ms_openings = ev_ms(ev[4], ev[5]);
  proof->ms_fri = ev_ms(ev[5], ev[6]);
  proof->ms_d2h = ev_ms(ev[6], ev[7]);
  proof->ms_total = ev_ms(ev[0], ev[7]);
  stage_collect(ctx);
}

// Pinned host buffers for proofs are expensive to create (cudaMallocHost of tens of MB) and cudaMallocHost / cudaFreeHost
// synchronise the whole device: with two proofs in flight a proof that allocates its buffer waits for the OTHER proof's
// kernels (measured: a 120 ms proof intermittently taking 450 ms).  sb_proof_free parks the buffers here, the next
// sb_prove takes the smallest one that fits, and nothing is freed until the pool holds 16 of them.
#include <mutex>
static std::mutex g_pool_mu;
static std::vector<std::pair<size_t, void*>> g_pool;
static std::map<void*, size_t> g_pinned_cap;     // capacity of every buffer handed out
static std::set<void*> g_unpinned;               // buffers that came from malloc (no CUDA device in this process)
static void* pinned_take(size_t bytes, size_t* got) {
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    size_t best = g_pool.size();
    for (size_t i = 0; i < g_pool.size(); i++)
      if (g_pool[i].first >= bytes && (best == g_pool.size() || g_pool[i].first < g_pool[best].first)) best = i;
    if (best != g_pool.size()) {
      void* p = g_pool[best].second;
      *got = g_pool[best].first;
      g_pool.erase(g_pool.begin() + best);
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, bytes);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
    // host-only use of the library (sb_proof_deserialize / sb_proof_from_words on a box without a GPU): pageable memory
    cudaGetLastError();
    p = malloc(bytes);
    if (!p) SB_THROW(SB_ENOMEM, "malloc(proof, %zu bytes) failed", bytes);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_unpinned.insert(p);
  } else if (e != cudaSuccess) SB_THROW(SB_ENOMEM, "cudaMallocHost(proof, %zu bytes): %s", bytes, cudaGetErrorString(e));
  *got = bytes;
  return p;
}
static void* pinned_take(size_t bytes) {
  size_t got = 0;
  void* p = pinned_take(bytes, &got);
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_pinned_cap[p] = got;
  return p;
}
static void pinned_give(void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto it = g_pinned_cap.find(p);
  if (it != g_pinned_cap.end()) { bytes = it->second; g_pinned_cap.erase(it); }
  if (g_unpinned.erase(p)) { free(p); return; }
  if (g_pool.size() >= 16) {
    size_t small = 0;
    for (size_t i = 1; i < g_pool.size(); i++) if (g_pool[i].first < g_pool[small].first) small = i;
    cudaFreeHost(g_pool[small].second);
    g_pool.erase(g_pool.begin() + small);
  }
  g_pool.push_back({bytes, p});
}

// error path: nothing of this proof may still be reading the caller's trace buffer or writing the proof buffer
static void quiesce(sb_ctx* ctx) {
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
}

// an empty proof object of the right shape (wire.cu: deserialisation)
sb_proof* proof_alloc(const sb_params& p) {
  sb_proof* proof = new sb_proof();
  memset(proof, 0, sizeof(*proof));
  try {
    proof->layout = proof_layout(p);
    proof->words = (u64*)pinned_take(8ull * proof->layout.total_words);
  } catch (...) { delete proof; throw; }
  return proof;
}

extern "C" {

int sb_prove(sb_ctx* ctx, const sb_params* p, const void* trace, int layout, const uint64_t* public_inputs, sb_proof** out) {
  if (!ctx || !p || !out) return SB_EINVAL;
  sb_proof* proof = nullptr;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    check_params(p);
    if (ctx->multi) return multi_prove(ctx, p, trace, layout, public_inputs, out);
    proof = new sb_proof();
    memset(proof, 0, sizeof(*proof));
    proof->layout = proof_layout(*p);
    proof->words = (u64*)pinned_take(8ull * proof->layout.total_words);
    prove_impl(ctx, p, trace, layout, public_inputs, proof);
    *out = proof;
    return SB_OK;
  } catch (const SbError& e) {
    quiesce(ctx);
    sb_proof_free(proof);
    return sb_fail(ctx, e);
  } catch (const std::exception& e) {
    quiesce(ctx);
    sb_proof_free(proof);
    return sb_fail(ctx, SbError{SB_EINVAL, e.what()});
  }
}

int sb_prove_sharded(sb_ctx* ctx, const sb_params* p, const sb_shard_hooks* hooks, const uint64_t* public_inputs, sb_proof** out) {
  if (!ctx || !p || !out || !hooks || !hooks->commit || !hooks->quotient || !hooks->openings || !hooks->combine || !hooks->query_rows)
    return SB_EINVAL;
  sb_proof* proof = nullptr;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    check_params(p);
    proof = new sb_proof();
    memset(proof, 0, sizeof(*proof));
    proof->layout = proof_layout(*p);
    proof->words = (u64*)pinned_take(8ull * proof->layout.total_words);
    prove_impl(ctx, p, nullptr, SB_TRACE_DEVICE_COLMAJOR_U64, public_inputs, proof, hooks);
    *out = proof;
    return SB_OK;
  } catch (const SbError& e) {
    quiesce(ctx);
    sb_proof_free(proof);
    return sb_fail(ctx, e);
  } catch (const std::exception& e) {
    quiesce(ctx);
    sb_proof_free(proof);
    return sb_fail(ctx, SbError{SB_EINVAL, e.what()});
  }
}

// P_c(zeta) and P_c(g zeta) for a column slice of coefficients (this rank's share of StarkOpeningSet::new)
int sb_openings_cols_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_coeffs, uint32_t n_cols_local, const uint64_t* zeta,
                            const uint64_t* zeta_next, uint64_t* local_out, uint64_t* next_out) {
  if (!ctx || !p || !d_coeffs || !zeta || !zeta_next || !local_out || !next_out) return SB_EINVAL;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    if (n_cols_local == 0) return SB_OK;
    const uint32_t n = 1u << p->log_n;
    ctx->scratch3.ensure(32ull * n + 32ull * n_cols_local + 1024);
    Arena ar(ctx->scratch3.p, ctx->scratch3.cap);
    e2_t* d_tab_a = ar.take<e2_t>(n);
    e2_t* d_tab_b = ar.take<e2_t>(n);
    e2_t* d_open = ar.take<e2_t>(2ull * n_cols_local);
    const e2_t za = e2_make(zeta[0], zeta[1]), zb = e2_make(zeta_next[0], zeta_next[1]);
    sb_openings_device(ctx, d_coeffs, p->log_n, n_cols_local, za, &zb, d_tab_a, d_tab_b, d_open, d_open + n_cols_local);
    CUDA_CHECK(cudaMemcpyAsync(local_out, d_open, 16ull * n_cols_local, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(next_out, d_open + n_cols_local, 16ull * n_cols_local, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

// sum_c alpha^(first_col + c) coeffs_c over a column slice (this rank's share of the batch reduce in prove_openings)
int sb_combine_cols_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_coeffs, uint32_t n_cols_local, const uint64_t* alpha,
                           uint32_t first_col, uint64_t* d_out) {
  if (!ctx || !p || !d_coeffs || !alpha || !d_out) return SB_EINVAL;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    const uint32_t n = 1u << p->log_n;
    if (n_cols_local == 0) { CUDA_CHECK(cudaMemsetAsync(d_out, 0, 16ull * n, ctx->stream)); return SB_OK; }
    const size_t partial_elems = (size_t)(ctx->sm_count * 16 + 8) * 128 + 2 * (size_t)n;
    ctx->scratch3.ensure(16ull * (n_cols_local + 8) + 16ull * partial_elems + 1024);
    Arena ar(ctx->scratch3.p, ctx->scratch3.cap);
    e2_t* d_apow = ar.take<e2_t>((size_t)n_cols_local + 8);
    e2_t* d_partial = ar.take<e2_t>(partial_elems);
    sb_combine_device(ctx, d_coeffs, p->log_n, n_cols_local, e2_make(alpha[0], alpha[1]), first_col, d_apow, d_partial, partial_elems,
                      (e2_t*)d_out);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

int sb_memcpy_device(sb_ctx* ctx, void* d_dst, const void* d_src, uint64_t bytes) {
  if (!ctx || !d_dst || !d_src) return SB_EINVAL;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    CUDA_CHECK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

// ---- stage exports: openings of the committed trace, FRI commit phase with injected challenges ----
// StarkOpeningSet::new on the trace committed by the preceding sb_lde_commit / sb_prove on this ctx:
// local_out[c] = P_c(zeta), next_out[c] = P_c(g zeta), extension elements as (c0, c1).
int sb_openings(sb_ctx* ctx, const sb_params* p, const uint64_t* zeta, uint64_t* local_out, uint64_t* next_out) {
  if (!ctx || !p || !zeta || !local_out || !next_out) return SB_EINVAL;
  try {
    check_params(p);
    CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!ctx->have_lde || ctx->cur.n_cols != p->n_cols || ctx->cur.log_n != p->log_n)
      SB_THROW(SB_EINVAL, "sb_openings needs a preceding sb_lde_commit with the same shape on this ctx");
    const e2_t z = e2_make(zeta[0], zeta[1]), zn = e2_scale(z, gl_root(p->log_n));
    const u64 zn_w[2] = {zn.a, zn.b};
    return sb_openings_cols_device(ctx, p, ctx->coeffs.as<u64>(), p->n_cols, zeta, zn_w, local_out, next_out);
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

// fri_committed_trees on a given polynomial: coeffs = [n][2] extension coefficients in natural order (the `final_poly` of
// prove_openings), betas = [rounds][2] folding challenges in place of the transcript.  caps_out: [rounds][2^cap_height][4],
// final_poly_out: [final_poly_len][2] (sb_proof_layout_for gives both counts).
int sb_fri_commit(sb_ctx* ctx, const sb_params* p, const uint64_t* coeffs, const uint64_t* betas, uint64_t* caps_out,
                  uint64_t* final_poly_out) {
  if (!ctx || !p || !coeffs || !caps_out || !final_poly_out) return SB_EINVAL;
  try {
    check_params(p);
    CUDA_CHECK(cudaSetDevice(ctx->device));
    const sb_proof_layout L = proof_layout(*p);
    if (L.n_fri_rounds && !betas) SB_THROW(SB_EINVAL, "betas is NULL");
    const uint32_t n = 1u << p->log_n, N = n << p->rate_bits;
    ctx->scratch2.ensure(16ull * N * 2 + 64ull * N + (1 << 16));
    Arena ar(ctx->scratch2.p, ctx->scratch2.cap);
    std::vector<e2_t> c(n);
    for (uint32_t i = 0; i < n; i++) c[i] = e2_make(coeffs[2 * i], coeffs[2 * i + 1]);
    std::vector<FriRound> rounds;
    fri_commit_phase(ctx, p, ar, c, caps_out, rounds, [&](size_t round, const u64*) { return e2_make(betas[2 * round], betas[2 * round + 1]); });
    if (c.size() != L.final_poly_len) SB_THROW(SB_EINVAL, "internal: final polynomial length %zu != %u", c.size(), L.final_poly_len);
    for (size_t i = 0; i < c.size(); i++) { final_poly_out[2 * i] = c[i].a; final_poly_out[2 * i + 1] = c[i].b; }
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

// FP12MulStark from its operands: witness generation on the host (witness.cpp), then the proof
int sb_prove_fp12_mul(sb_ctx* ctx, const sb_params* p, const uint32_t* x, const uint32_t* y, sb_proof** out) {
  if (!ctx || !p || !x || !y || !out) return SB_EINVAL;
  try {
    check_params(p);
    if (p->stark_id != SB_STARK_FP12_MUL || p->n_cols != 60285 || p->n_public_inputs != 432)
      SB_THROW(SB_EINVAL, "sb_prove_fp12_mul needs the FP12MulStark parameters (60285 columns, 432 public inputs)");
    const uint32_t rows = 1u << p->log_n;
    std::vector<uint32_t> trace((size_t)rows * p->n_cols);
    std::vector<uint64_t> pis(p->n_public_inputs);
    if (sb_witness_fp12_mul(x, y, rows, trace.data(), pis.data())) SB_THROW(SB_EINVAL, "%s", sb_witness_last_error());
    return sb_prove(ctx, p, trace.data(), SB_TRACE_ROWMAJOR_U32, pis.data(), out);
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

// ECCAggStark from the 512 points and participation bits (witness.cpp), then the proof
int sb_prove_ecc_agg(sb_ctx* ctx, const sb_params* p, const uint32_t* points, const uint8_t* bits, sb_proof** out) {
  if (!ctx || !p || !points || !bits || !out) return SB_EINVAL;
  try {
    check_params(p);
    if (p->stark_id != SB_STARK_ECC_AGG || p->n_cols != 3339 || p->n_public_inputs != 12824)
      SB_THROW(SB_EINVAL, "sb_prove_ecc_agg needs the ECCAggStark parameters (3339 columns, 12824 public inputs)");
    const uint32_t rows = 1u << p->log_n;
    std::vector<uint32_t> trace((size_t)rows * p->n_cols);
    std::vector<uint64_t> pis(p->n_public_inputs);
    if (sb_witness_ecc_agg(points, bits, rows, trace.data(), pis.data(), nullptr)) SB_THROW(SB_EINVAL, "%s", sb_witness_last_error());
    return sb_prove(ctx, p, trace.data(), SB_TRACE_ROWMAJOR_U32, pis.data(), out);
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}


// the three larger starks from their operands (witness.cpp), then the proof.  The trace is generated into a malloc'd
// row-major u32 buffer (FinalExp: 2.4 GB) and crosses PCIe once, in slabs, behind the commitment kernels.
static int prove_from_witness(sb_ctx* ctx, const sb_params* p, uint32_t stark_id, uint32_t n_cols, uint32_t n_pis, const char* what,
                              const std::function<int(uint32_t rows, uint32_t* trace, uint64_t* pis)>& gen, sb_proof** out) {
  try {
    check_params(p);
    if (p->stark_id != stark_id || p->n_cols != n_cols || p->n_public_inputs != n_pis)
      SB_THROW(SB_EINVAL, "%s needs that stark's parameters (%u columns, %u public inputs)", what, n_cols, n_pis);
    const uint32_t rows = 1u << p->log_n;
    std::unique_ptr<uint32_t, void (*)(void*)> trace((uint32_t*)malloc(4ull * rows * n_cols), free);
    if (!trace) SB_THROW(SB_EINVAL, "%s: out of host memory for the trace", what);
    std::vector<uint64_t> pis(n_pis);
    if (gen(rows, trace.get(), pis.data())) SB_THROW(SB_EINVAL, "%s", sb_witness_last_error());
    return sb_prove(ctx, p, trace.get(), SB_TRACE_ROWMAJOR_U32, pis.data(), out);
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}
int sb_prove_pairing_precomp(sb_ctx* ctx, const sb_params* p, const uint32_t* q, sb_proof** out) {
  if (!ctx || !p || !q || !out) return SB_EINVAL;
  return prove_from_witness(ctx, p, SB_STARK_PAIRING_PRECOMP, 29376, 4968, "sb_prove_pairing_precomp",
                            [&](uint32_t rows, uint32_t* t, uint64_t* pi) { return sb_witness_pairing_precomp(q, rows, t, pi); }, out);
}
int sb_prove_miller_loop(sb_ctx* ctx, const sb_params* p, const uint32_t* g1, const uint32_t* q, sb_proof** out) {
  if (!ctx || !p || !g1 || !q || !out) return SB_EINVAL;
  return prove_from_witness(ctx, p, SB_STARK_MILLER_LOOP, 97330, 5064, "sb_prove_miller_loop",
                            [&](uint32_t rows, uint32_t* t, uint64_t* pi) { return sb_witness_miller_loop(g1, q, rows, t, pi); }, out);
}
int sb_prove_final_exp(sb_ctx* ctx, const sb_params* p, const uint32_t* x, sb_proof** out) {
  if (!ctx || !p || !x || !out) return SB_EINVAL;
  return prove_from_witness(ctx, p, SB_STARK_FINAL_EXP, 73527, 288, "sb_prove_final_exp",
                            [&](uint32_t rows, uint32_t* t, uint64_t* pi) { return sb_witness_final_exp(x, rows, t, pi); }, out);
}

void sb_proof_free(sb_proof* proof) {
  if (!proof) return;
  if (proof->words) pinned_give(proof->words, 8ull * proof->layout.total_words);
  delete proof;
}

}  // extern "C"
