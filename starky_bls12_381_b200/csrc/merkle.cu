// K2 / K3: Poseidon-12 Merkle commitment of an LDE batch for sm_100a.
// Replaces plonky2::hash::merkle_tree::MerkleTree::new(leaves, cap_height) with PoseidonHash::hash_or_noop leaves
// and two_to_one inner nodes, as called from PolynomialBatch::from_values / from_coeffs
// inside starky::prover::prove (reference call sites /root/reference/src/aggregate_proof.rs:59,105,138,169,212).
//
// The LDE is never transposed into plonky2's Vec<Vec<F>> leaves: a leaf is "all columns at one position" of the
// column-major [C][N] batch, so the thread(s) owning a position stream down the columns (coalesced across positions)
// through the rate-8 overwrite sponge and scatter only the 32-byte digest to plonky2's leaf index.
//
// Tree storage: one buffer of 4-word digests, level 0 (N leaf digests, plonky2 leaf order) first, then N/2, ...
// down to the cap (2^cap_height digests).  Level l starts at digest offset 2N - (2N >> l).
#include <stdlib.h>

#include "common.cuh"
#include "poseidon.cuh"

// position (coset-major, see ntt.cu) -> plonky2 leaf index:  J*n + k  ->  J*n + bitrev_n(k)
__device__ __forceinline__ uint32_t leaf_index_of(uint32_t pos, unsigned log_block) {
  uint32_t mask = (1u << log_block) - 1;
  return (pos & ~mask) | bitrev32(pos & mask, log_block);
}

// One thread per leaf.  cols: [leaf_len][n_leaves].
__global__ void __launch_bounds__(128) leaf_hash_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                        uint32_t n_leaves, unsigned log_block,
                                                        u64* __restrict__ digests) {
  uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n_leaves) return;
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  const u64* p = cols + pos;
  if (leaf_len <= 4) {  // hash_or_noop: short leaves are copied, not hashed
#pragma unroll
    for (int i = 0; i < 4; i++) if ((uint32_t)i < leaf_len) s[i] = p[(size_t)i * n_leaves];
  } else {
    uint32_t full = leaf_len / 8, rem = leaf_len % 8;
    u64 nx[8];
    if (full) {
#pragma unroll
      for (int i = 0; i < 8; i++) nx[i] = p[(size_t)i * n_leaves];
    }
    for (uint32_t c = 0; c < full; c++) {
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = nx[i];
      if (c + 1 < full) {  // prefetch the next 8 columns under the permutation
        const u64* q = p + (size_t)(c + 1) * 8 * n_leaves;
#pragma unroll
        for (int i = 0; i < 8; i++) nx[i] = q[(size_t)i * n_leaves];
      }
      poseidon_permute(s);
    }
    if (rem) {
      const u64* q = p + (size_t)full * 8 * n_leaves;
#pragma unroll
      for (int i = 0; i < 8; i++) if ((uint32_t)i < rem) s[i] = q[(size_t)i * n_leaves];
      poseidon_permute(s);
    }
  }
  u64* d = digests + 4ull * leaf_index_of(pos, log_block);
  d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

// ---------------------------------------------------------------------------------------------------------
// Cooperative leaf sponge: L lanes share one leaf's 12-word state (12/L words each).  A leaf is a strictly sequential
// chain of ceil(C/8) permutations, so with only N = 2048..32768 leaves one thread per leaf leaves the machine idle
// (and is latency-bound at ~31k dependent-ish instructions per permutation); splitting the state over lanes cuts the
// chain latency ~L-fold.  Per round each lane applies constants + S-box to its own words, publishes them through a
// double-buffered shared-memory exchange (one __syncwarp per round), reads back all 12 words with 128-bit loads and
// forms its own MDS rows with per-lane rotated coefficient registers.  Input columns are staged through shared memory
// by the whole block ([8 columns][leaves] tiles, coalesced along the leaf axis, prefetched one permutation ahead).
// ---------------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(256) leaf_hash_coop_kernel(const u64* __restrict__ cols, uint32_t leaf_len,
                                                             uint32_t n_leaves, unsigned log_block,
                                                             u64* __restrict__ digests) {
  constexpr int G = (L == 12) ? 16 : L;   // lanes reserved per leaf
  constexpr int W = 12 / L;               // state words per lane
  constexpr int LPB = 256 / G;            // leaves per block
  constexpr int EPT = (8 * LPB + 255) / 256;
  __shared__ __align__(16) u64 tile[2][8][LPB];
  __shared__ __align__(16) u64 xch[2][LPB][12];
  __shared__ u64 rc[POSEIDON_RC_COUNT];
  const unsigned tid = threadIdx.x;
  const unsigned g = tid / G, r = tid % G;
  const bool active = r < (unsigned)L;
  const unsigned w0 = r * W;
  const uint32_t pos0 = blockIdx.x * LPB;
  for (unsigned i = tid; i < POSEIDON_RC_COUNT; i += 256) rc[i] = c_poseidon_rc[i];

  // rotated MDS coefficients of this lane's rows: out[row] = sum_j CIRC[(j - row) mod 12] * t[j] (+ 8 t[0] on row 0)
  const u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u32 coef[W][12];
#pragma unroll
  for (int k = 0; k < W; k++)
#pragma unroll
    for (int j = 0; j < 12; j++) {
      unsigned row = (w0 + k) % 12;
      u32 c = 0;
#pragma unroll
      for (int i = 0; i < 12; i++) c = ((j + 12 - row) % 12 == (unsigned)i) ? CIRC[i] : c;
      coef[k][j] = c + ((row == 0 && j == 0) ? 8u : 0u);
    }

  u64 s[W];
#pragma unroll
  for (int k = 0; k < W; k++) s[k] = 0;

  const uint32_t n_chunks = (leaf_len + 7) / 8;
  // stage chunk 0
  u64 pre[EPT];
  auto fetch = [&](uint32_t chunk) {
#pragma unroll
    for (int q = 0; q < EPT; q++) {
      unsigned e = tid + q * 256;
      unsigned col = e / LPB, leaf = e % LPB;
      uint32_t c = chunk * 8 + col, pos = pos0 + leaf;
      pre[q] = (e < 8 * LPB && c < leaf_len && pos < n_leaves) ? cols[(size_t)c * n_leaves + pos] : 0;
    }
  };
  auto stash = [&](unsigned buf) {
#pragma unroll
    for (int q = 0; q < EPT; q++) {
      unsigned e = tid + q * 256;
      if (e < 8 * LPB) tile[buf][e / LPB][e % LPB] = pre[q];
    }
  };
  fetch(0);
  stash(0);
  unsigned xb = 0;
  for (uint32_t m = 0; m < n_chunks; m++) {
    __syncthreads();
    const unsigned take = min(8u, leaf_len - 8 * m);
    if (active) {
#pragma unroll
      for (int k = 0; k < W; k++)
        if (w0 + k < take) s[k] = tile[m & 1][w0 + k][g];
    }
    if (m + 1 < n_chunks) fetch(m + 1);
    // ---- one permutation: 4 full, 22 partial, 4 full rounds ----
    auto linear_layer = [&]() {
      __syncwarp();
      if (active) {
        u32 lo[12], hi[12];
        const ulonglong2* src = (const ulonglong2*)&xch[xb][g][0];
#pragma unroll
        for (int j = 0; j < 6; j++) {
          ulonglong2 v = src[j];
          lo[2 * j] = (u32)v.x; hi[2 * j] = (u32)(v.x >> 32);
          lo[2 * j + 1] = (u32)v.y; hi[2 * j + 1] = (u32)(v.y >> 32);
        }
#pragma unroll
        for (int k = 0; k < W; k++) {
          u64 al = 0, ah = 0;
#pragma unroll
          for (int j = 0; j < 12; j++) {
            al += (u64)coef[k][j] * lo[j];
            ah += (u64)coef[k][j] * hi[j];
          }
          u64 c = (ah >> 32) * GL_EPS, b = (ah & GL_EPS) << 32, t = al + c, v = b + t;
          if (v < t) v += GL_EPS;
          s[k] = v;
        }
      }
      xb ^= 1;
    };
    auto full_round = [&](int rd) {
      if (active) {
#pragma unroll
        for (int k = 0; k < W; k++) xch[xb][g][w0 + k] = poseidon_sbox(gl_add_lazy_canon(s[k], rc[12 * rd + w0 + k]));
      }
      linear_layer();
    };
#pragma unroll 1
    for (int rd = 0; rd < 4; rd++) full_round(rd);
#pragma unroll 1
    for (int rd = 4; rd < 26; rd++) {
      if (active) {
        u64 v0 = gl_add_lazy_canon(s[0], rc[12 * rd + w0]);
        if (w0 == 0) v0 = poseidon_sbox(v0);       // only the lane holding word 0
        xch[xb][g][w0] = v0;
#pragma unroll
        for (int k = 1; k < W; k++) xch[xb][g][w0 + k] = gl_add_lazy_canon(s[k], rc[12 * rd + w0 + k]);
      }
      linear_layer();
    }
#pragma unroll 1
    for (int rd = 26; rd < 30; rd++) full_round(rd);
    if (m + 1 < n_chunks) stash((m + 1) & 1);
  }
  // digest = state words 0..3, canonical, scattered to plonky2's leaf index
  const uint32_t pos = pos0 + g;
  if (active && pos < n_leaves) {
    u64* d = digests + 4ull * leaf_index_of(pos, log_block);
#pragma unroll
    for (int k = 0; k < W; k++)
      if (w0 + k < 4) d[w0 + k] = gl_canon(s[k]);
  }
}

void sb_hash_leaves_device(sb_ctx* ctx, const u64* d_cols, uint32_t leaf_len, uint32_t n_leaves, unsigned log_block,
                           u64* d_digests) {
  const uint32_t perms = (leaf_len + 7) / 8;
  // lanes per leaf: enough threads to fill the machine, but never wider than the chain needs.
  // SB_LEAF_LANES=1|4|12 overrides the choice (profiling).
  int lanes = 1;
  if (leaf_len > 4 && perms >= 8) {
    if ((uint64_t)n_leaves * 12 <= (uint64_t)ctx->sm_count * 512) lanes = 12;
    else if ((uint64_t)n_leaves * 4 <= (uint64_t)ctx->sm_count * 1024) lanes = 4;
  }
  if (const char* e = getenv("SB_LEAF_LANES")) { int v = atoi(e); if (v == 1 || v == 4 || v == 12) lanes = v; }
  if (lanes == 12) {
    LAUNCH(ctx, leaf_hash_coop_kernel<12>, (n_leaves + 15) / 16, 256, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else if (lanes == 4) {
    LAUNCH(ctx, leaf_hash_coop_kernel<4>, (n_leaves + 63) / 64, 256, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else {
    unsigned block = n_leaves >= 128 * (unsigned)ctx->sm_count ? 128 : (n_leaves >= 64 * (unsigned)ctx->sm_count ? 64 : 32);
    LAUNCH(ctx, leaf_hash_kernel, (n_leaves + block - 1) / block, block, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  }
}

// one level: out[i] = two_to_one(in[2i], in[2i+1])
__global__ void merkle_level_kernel(const u64* __restrict__ in, u64* __restrict__ out, uint32_t n_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const ulonglong2* src = (const ulonglong2*)(in + 8ull * i);
  ulonglong2 a = src[0], b = src[1], c = src[2], d = src[3];
  u64 l[4] = {a.x, a.y, b.x, b.y}, r[4] = {c.x, c.y, d.x, d.y}, o[4];
  poseidon_two_to_one(l, r, o);
  ulonglong2* dst = (ulonglong2*)(out + 4ull * i);
  dst[0] = make_ulonglong2(o[0], o[1]);
  dst[1] = make_ulonglong2(o[2], o[3]);
}

void sb_merkle_levels(sb_ctx* ctx, u64* d_tree, uint32_t n_leaves, unsigned cap_height) {
  uint32_t cap = 1u << cap_height;
  u64* in = d_tree;
  for (uint32_t cnt = n_leaves; cnt > cap; cnt >>= 1) {
    u64* out = in + 4ull * cnt;
    uint32_t n_out = cnt >> 1;
    unsigned block = n_out >= 4096 ? 64 : 32;
    LAUNCH(ctx, merkle_level_kernel, (n_out + block - 1) / block, block, 0, in, out, n_out);
    in = out;
  }
}

__global__ void permute_kernel(u64* states, uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 s[12];
#pragma unroll
  for (int k = 0; k < 12; k++) s[k] = states[12ull * i + k];
  poseidon_permute(s);
#pragma unroll
  for (int k = 0; k < 12; k++) states[12ull * i + k] = s[k];
}
void sb_poseidon_permute_device(sb_ctx* ctx, u64* d_states, uint32_t count) {
  LAUNCH(ctx, permute_kernel, (count + 63) / 64, 64, 0, d_states, count);
}
