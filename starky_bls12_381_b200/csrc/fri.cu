// K6-K10: openings at zeta / g*zeta, the batched FRI combine, FRI commit-phase hashing, proof-of-work grinding and
// the query gathers, for sm_100a.  Replaces StarkOpeningSet::new, PolynomialBatch::prove_openings,
// fri_committed_trees, fri_proof_of_work and fri_prover_query_rounds (plonky2 / starky),
// reached from starky::prover::prove (/root/reference/src/aggregate_proof.rs:59,105,138,169,212).
#include <algorithm>

#include "poseidon.cuh"
#include "prover.cuh"

// ---------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 warp_sum_gl(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = gl_add(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// tab[p] = z^{bitrev_n(p)} for p < n   (coefficients are stored in bit-reversed coefficient order, ntt.cu)
__global__ void ext_pow_table_kernel(e2_t* tab, e2_t z, unsigned log_n) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (1u << log_n)) return;
  tab[p] = e2_pow(z, bitrev32(p, log_n));
}
// tab[j] = a^(j + j0)
__global__ void ext_pow_seq_kernel(e2_t* tab, e2_t a, uint32_t count, uint32_t j0) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  tab[j] = e2_pow(a, (u64)j + j0);
}

// ---------------------------------------------------------------------------------------------------------
// K6: out_a[c] = P_c(za), out_b[c] = P_c(zb) in F_p^2, P_c given by coeffs[c][p] = c_{bitrev(p)}.
// One 128-thread block per polynomial; HBM: one pass over the coefficients.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) openings_kernel(const u64* __restrict__ coeffs, uint32_t n, uint32_t n_polys,
                                                       const e2_t* __restrict__ tab_a, const e2_t* __restrict__ tab_b,
                                                       e2_t* __restrict__ out_a, e2_t* __restrict__ out_b) {
  __shared__ u64 red[4][4];
  const uint32_t c = blockIdx.x;
  if (c >= n_polys) return;
  const u64* co = coeffs + (size_t)c * n;
  u64 a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  for (uint32_t p = threadIdx.x; p < n; p += blockDim.x) {
    u64 v = co[p];
    e2_t ta = tab_a[p];
    a0 = gl_add(a0, gl_mul(v, ta.a));
    a1 = gl_add(a1, gl_mul(v, ta.b));
    if (tab_b) {
      e2_t tb = tab_b[p];
      b0 = gl_add(b0, gl_mul(v, tb.a));
      b1 = gl_add(b1, gl_mul(v, tb.b));
    }
  }
  a0 = warp_sum_gl(a0); a1 = warp_sum_gl(a1); b0 = warp_sum_gl(b0); b1 = warp_sum_gl(b1);
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { red[warp][0] = a0; red[warp][1] = a1; red[warp][2] = b0; red[warp][3] = b1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned w = 1; w < nw; w++) {
      a0 = gl_add(a0, red[w][0]); a1 = gl_add(a1, red[w][1]); b0 = gl_add(b0, red[w][2]); b1 = gl_add(b1, red[w][3]);
    }
    out_a[c] = e2_make(a0, a1);
    if (tab_b) out_b[c] = e2_make(b0, b1);
  }
}

void sb_openings_device(sb_ctx* ctx, const u64* d_coeffs, unsigned log_n, uint32_t n_polys, e2_t za, const e2_t* zb,
                        e2_t* d_tab_a, e2_t* d_tab_b, e2_t* d_out_a, e2_t* d_out_b) {
  const uint32_t n = 1u << log_n;
  LAUNCH(ctx, ext_pow_table_kernel, (n + 127) / 128, 128, 0, d_tab_a, za, log_n);
  if (zb) LAUNCH(ctx, ext_pow_table_kernel, (n + 127) / 128, 128, 0, d_tab_b, *zb, log_n);
  unsigned block = n < 128 ? (n < 32 ? 32 : n) : 128;
  LAUNCH(ctx, openings_kernel, n_polys, block, 0, d_coeffs, n, n_polys, d_tab_a, zb ? d_tab_b : nullptr, d_out_a,
         zb ? d_out_b : nullptr);
}

// ---------------------------------------------------------------------------------------------------------
// K7: batch combine  F[p] = sum_{c < n_polys} apow[c] * coeffs[c][p]   (base coefficient x extension weight).
// grid (n / 128, column chunks): partial[chunk][p], then a small reduce.  One pass over the coefficients.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) combine_kernel(const u64* __restrict__ coeffs, uint32_t n, uint32_t n_polys,
                                                      const e2_t* __restrict__ apow, uint32_t cols_per_chunk,
                                                      e2_t* __restrict__ partial) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t c0 = blockIdx.y * cols_per_chunk;
  const uint32_t c1 = min(n_polys, c0 + cols_per_chunk);
  u64 s0 = 0, s1 = 0;
  for (uint32_t c = c0; c < c1; c++) {
    u64 v = coeffs[(size_t)c * n + p];
    e2_t a = apow[c];
    s0 = gl_add(s0, gl_mul(v, a.a));
    s1 = gl_add(s1, gl_mul(v, a.b));
  }
  partial[(size_t)blockIdx.y * n + p] = e2_make(s0, s1);
}
__global__ void combine_reduce_kernel(const e2_t* __restrict__ partial, uint32_t n, uint32_t n_chunks, e2_t* __restrict__ out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  e2_t s = e2_make(0, 0);
  for (uint32_t c = 0; c < n_chunks; c++) s = e2_add(s, partial[(size_t)c * n + p]);
  out[p] = s;
}

void sb_combine_device(sb_ctx* ctx, const u64* d_coeffs, unsigned log_n, uint32_t n_polys, e2_t alpha, uint32_t j0,
                       e2_t* d_apow, e2_t* d_partial, size_t partial_capacity_elems, e2_t* d_out) {
  const uint32_t n = 1u << log_n;
  LAUNCH(ctx, ext_pow_seq_kernel, (n_polys + 127) / 128, 128, 0, d_apow, alpha, n_polys, j0);
  unsigned block = n < 128 ? n : 128;
  uint32_t xt = (n + block - 1) / block;
  uint32_t want_chunks = (uint32_t)((ctx->sm_count * 16 + xt - 1) / xt);
  if (want_chunks > n_polys) want_chunks = n_polys;
  if ((size_t)want_chunks * n > partial_capacity_elems) want_chunks = (uint32_t)(partial_capacity_elems / n);
  if (want_chunks < 1) want_chunks = 1;
  uint32_t cpc = (n_polys + want_chunks - 1) / want_chunks;
  uint32_t n_chunks = (n_polys + cpc - 1) / cpc;
  LAUNCH(ctx, combine_kernel, dim3(xt, n_chunks), block, 0, d_coeffs, n, n_polys, d_apow, cpc, d_partial);
  LAUNCH(ctx, combine_reduce_kernel, xt, block, 0, d_partial, n, n_chunks, d_out);
}

// ---------------------------------------------------------------------------------------------------------
// K8: FRI commit-phase leaves.  values: re[], im[] in bit-reversed order; leaf l = the 2^arity_bits consecutive
// extension values starting at l << arity_bits, flattened [re, im, re, im, ...] (A.9).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) fri_leaf_hash_kernel(const u64* __restrict__ re, const u64* __restrict__ im,
                                                           uint32_t n_leaves, unsigned arity_bits, u64* __restrict__ digests) {
  uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_leaves) return;
  const uint32_t arity = 1u << arity_bits;
  const u64* r = re + ((size_t)l << arity_bits);
  const u64* m = im + ((size_t)l << arity_bits);
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  if (2 * arity <= 4) {  // hash_or_noop
    s[0] = r[0]; s[1] = m[0];
    if (arity == 2) { s[2] = r[1]; s[3] = m[1]; }
  } else {
    for (uint32_t e = 0; e < arity; e += 4) {   // 4 extension values = 8 field elements = one absorb
#pragma unroll
      for (int i = 0; i < 4; i++) { s[2 * i] = r[e + i]; s[2 * i + 1] = m[e + i]; }
      poseidon_permute(s);
    }
  }
  u64* d = digests + 4ull * l;
  d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}
void sb_fri_leaf_hash_device(sb_ctx* ctx, const u64* d_re, const u64* d_im, uint32_t n_leaves, unsigned arity_bits, u64* d_digests) {
  if (arity_bits == 0) SB_THROW(SB_EINVAL, "fri arity bits must be >= 1");
  LAUNCH(ctx, fri_leaf_hash_kernel, (n_leaves + 63) / 64, 64, 0, d_re, d_im, n_leaves, arity_bits, d_digests);
}

// ---------------------------------------------------------------------------------------------------------
// K9: proof-of-work.  state = challenger sponge state with the pending inputs written in; candidate goes to
// slot `pos`; accept when the duplex response state[7] has >= bits leading zeros.  Returns the SMALLEST witness in
// the scanned window (the reference's rayon find_any returns any; smallest is deterministic).
// ---------------------------------------------------------------------------------------------------------
struct PowState { u64 s[12]; };
__global__ void __launch_bounds__(128) pow_kernel(PowState st, int pos, unsigned bits, u64 base, unsigned long long* best) {
  u64 cand = base + (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (cand >= GL_P) return;
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = (i == pos) ? cand : st.s[i];
  poseidon_permute(s);
  if (bits == 0 || (s[7] >> (64 - bits)) == 0) atomicMin(best, (unsigned long long)cand);
}
u64 sb_pow_device(sb_ctx* ctx, const u64 state[12], int pos, unsigned bits, unsigned long long* d_best) {
  PowState st;
  for (int i = 0; i < 12; i++) st.s[i] = state[i];
  // windows are scanned in order, so the SMALLEST witness is returned whatever their size; the first window is sized for
  // the expected 2^bits candidates (x4: found in it with probability 1 - e^-4), later ones grow to keep the machine full
  u64 window = std::max<u64>(1ull << 14, std::min<u64>(1ull << 20, 4ull << bits));
  for (u64 base = 0; base < (1ull << 40); base += window, window = 1ull << 20) {
    unsigned long long init = ~0ull, got = ~0ull;
    CUDA_CHECK(cudaMemcpyAsync(d_best, &init, 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, pow_kernel, (unsigned)(window / 128), 128, 0, st, pos, bits, base, d_best);
    CUDA_CHECK(cudaMemcpyAsync(&got, d_best, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (got != ~0ull) return got;
  }
  SB_THROW(SB_EPOW, "Proof of work failed. This is highly unlikely!");
}

// ---------------------------------------------------------------------------------------------------------
// K10: query gathers, straight into the (device copy of the) query records of the proof.
// ---------------------------------------------------------------------------------------------------------
// dst[q * stride + off + c] = cols[c * N + position[q]]
__global__ void gather_leaf_kernel(u64* __restrict__ dst, uint64_t stride, uint64_t off, const u64* __restrict__ cols,
                                   uint32_t N, uint32_t n_cols, const uint32_t* __restrict__ positions) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t q = blockIdx.y;
  if (c >= n_cols) return;
  dst[q * stride + off + c] = cols[(size_t)c * N + positions[q]];
}
// Merkle path of leaf_idx[q] >> shift: dst[q*stride + off + 4*l .. +4] = tree level l sibling
__global__ void gather_path_kernel(u64* __restrict__ dst, uint64_t stride, uint64_t off, const u64* __restrict__ tree,
                                   uint32_t n_leaves, uint32_t path_len, const uint32_t* __restrict__ leaf_idx, unsigned shift) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t q = blockIdx.y;
  if (t >= path_len * 4) return;
  uint32_t l = t >> 2, w = t & 3;
  uint32_t idx = leaf_idx[q] >> shift;
  size_t level_off = 2 * (size_t)n_leaves - ((2 * (size_t)n_leaves) >> l);
  dst[q * stride + off + t] = tree[4 * (level_off + ((idx >> l) ^ 1)) + w];
}
// FRI step evals: 2^arity_bits extension values of leaf (leaf_idx[q] >> shift), interleaved re/im
__global__ void gather_fri_evals_kernel(u64* __restrict__ dst, uint64_t stride, uint64_t off, const u64* __restrict__ re,
                                        const u64* __restrict__ im, unsigned arity_bits, const uint32_t* __restrict__ leaf_idx,
                                        unsigned shift) {
  uint32_t t = threadIdx.x;
  uint32_t q = blockIdx.x;
  if (t >= (2u << arity_bits)) return;
  uint32_t leaf = leaf_idx[q] >> shift;
  size_t v = ((size_t)leaf << arity_bits) + (t >> 1);
  dst[q * stride + off + t] = (t & 1) ? im[v] : re[v];
}

void sb_gather_leaf_device(sb_ctx* ctx, u64* dst, uint64_t stride, uint64_t off, const u64* cols, uint32_t N, uint32_t n_cols,
                           const uint32_t* d_positions, uint32_t n_queries) {
  LAUNCH(ctx, gather_leaf_kernel, dim3((n_cols + 127) / 128, n_queries), 128, 0, dst, stride, off, cols, N, n_cols, d_positions);
}
void sb_gather_path_device(sb_ctx* ctx, u64* dst, uint64_t stride, uint64_t off, const u64* tree, uint32_t n_leaves,
                           uint32_t path_len, const uint32_t* d_leaf_idx, unsigned shift, uint32_t n_queries) {
  if (path_len == 0) return;
  LAUNCH(ctx, gather_path_kernel, dim3((path_len * 4 + 63) / 64, n_queries), 64, 0, dst, stride, off, tree, n_leaves, path_len,
         d_leaf_idx, shift);
}
void sb_gather_fri_evals_device(sb_ctx* ctx, u64* dst, uint64_t stride, uint64_t off, const u64* re, const u64* im,
                                unsigned arity_bits, const uint32_t* d_leaf_idx, unsigned shift, uint32_t n_queries) {
  LAUNCH(ctx, gather_fri_evals_kernel, n_queries, 2u << arity_bits, 0, dst, stride, off, re, im, arity_bits, d_leaf_idx, shift);
}

// ---------------------------------------------------------------------------------------------------------
// permutations between the orders used on the device
// ---------------------------------------------------------------------------------------------------------
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
// coset-major positions (J*n + k) -> fully bit-reversed natural order (J*n + bitrev_n(k)), `count` vectors of N
__global__ void coset_to_bitrev_kernel(const u64* __restrict__ in, u64* __restrict__ out, unsigned log_n, unsigned log_N,
                                       uint64_t total) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  uint64_t vec = i >> log_N;
  uint32_t pos = (uint32_t)(i & ((1ull << log_N) - 1));
  uint32_t mask = (1u << log_n) - 1;
  out[(vec << log_N) + ((pos & ~mask) | bitrev32(pos & mask, log_n))] = in[i];
}
void sb_coset_to_bitrev_device(sb_ctx* ctx, const u64* d_in, u64* d_out, unsigned log_n, unsigned log_N, uint32_t count) {
  uint64_t total = (uint64_t)count << log_N;
  LAUNCH(ctx, coset_to_bitrev_kernel, (unsigned)((total + 255) / 256), 256, 0, d_in, d_out, log_n, log_N, total);
}
// d[v][m] *= s^m  (undo / apply a coset shift on coefficient vectors)
__global__ void coset_shift_kernel(u64* d, unsigned log_size, uint64_t total, u64 s) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  uint64_t m = i & ((1ull << log_size) - 1);
  d[i] = gl_mul(d[i], gl_pow(s, m));
}
void sb_coset_shift_device(sb_ctx* ctx, u64* d, unsigned log_size, uint32_t count, u64 s) {
  uint64_t total = (uint64_t)count << log_size;
  LAUNCH(ctx, coset_shift_kernel, (unsigned)((total + 255) / 256), 256, 0, d, log_size, total, s);
}
// flag = 1 if any d[v][m] != 0 for m >= keep
__global__ void tail_nonzero_kernel(const u64* __restrict__ d, unsigned log_size, uint64_t total, uint32_t keep, int* flag) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  uint64_t m = i & ((1ull << log_size) - 1);
  if (m >= keep && d[i] != 0) *flag = 1;
}
void sb_tail_nonzero_device(sb_ctx* ctx, const u64* d, unsigned log_size, uint32_t count, uint32_t keep, int* d_flag) {
  uint64_t total = (uint64_t)count << log_size;
  LAUNCH(ctx, tail_nonzero_kernel, (unsigned)((total + 255) / 256), 256, 0, d, log_size, total, keep, d_flag);
}
