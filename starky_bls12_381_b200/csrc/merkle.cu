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
#include "leafhash.cuh"
#include "leafhash_mm.cuh"

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

// leaves per warp of the matrix-instruction sponge: 8 up to 64 leaves per SM (latency-bound), 32 once that still leaves
// ~1.5 warps per scheduler, 16 in between (measured, lab "mm")
// 9 = one block per SM that gives every scheduler three 16-leaf warps and one 8-leaf warp (leaf_sponge_mm_het_kernel), when
// the leaves fill that shape to more than 6/7 (FinalExp / ECCAgg: 32768 leaves on 148 SMs, 33152 places)
typedef MmHet<0, 3, 1> HetShape;
// 10 = the 8-leaf form with helper lanes (four leaves per warp, the other half of the warp takes a third of the full-round
// S-boxes) while that still leaves at most one warp per scheduler (MillerLoop 2048, FP12Mul 32 leaves)
static int sb_mm_kind(sb_ctx* ctx, uint32_t n_leaves) {
  if ((uint64_t)n_leaves <= 16ull * ctx->sm_count) return 10;
  if ((uint64_t)n_leaves <= 64ull * ctx->sm_count) return 8;
  const uint64_t places = (uint64_t)HetShape::LEAVES * ctx->sm_count;
  if (n_leaves <= places && 7ull * n_leaves > 6 * places) return 9;
  return (uint64_t)n_leaves >= 160ull * ctx->sm_count ? 7 : 6;
}
static void launch_mm_het(sb_ctx* ctx, const u64* d_cols, uint32_t leaf_len, uint32_t n_leaves, unsigned log_block, u64* d_digests,
                          const u64* in, u64* out) {
  // 120 KB of (unused) dynamic shared memory: two blocks never share an SM.  Per device: function attributes are per context.
  CUDA_CHECK(cudaFuncSetAttribute(leaf_sponge_mm_het_kernel<0, 3, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 << 10));
  LAUNCH(ctx, (leaf_sponge_mm_het_kernel<0, 3, 1, 0>), (n_leaves + HetShape::LEAVES - 1) / HetShape::LEAVES, HetShape::THREADS, 120 << 10,
         d_cols, leaf_len, n_leaves, log_block, d_digests, in, out);
}

void sb_hash_leaves_device(sb_ctx* ctx, const u64* d_cols, uint32_t leaf_len, uint32_t n_leaves, unsigned log_block,
                           u64* d_digests) {
  const uint32_t perms = (leaf_len + 7) / 8;
  // Kernel choice (leafhash.cuh, leafhash_mm.cuh; measured on B200 with tools/perf/poseidon_lab.cu, profiles/r2_poseidon_lab_mm.txt):
  //   short chains (quotient / FRI leaves, <= 2 permutations): one thread per leaf, nothing to split;
  //   everything else: the dense MDS layer as an integer matrix instruction over the lanes of one warp (leafhash_mm.cuh),
  //     8 = 8 leaves per warp for the latency-bound shapes (PairingPrecomp 4096, MillerLoop 2048 leaves: every leaf is in
  //         flight at once and wall time = chain length x latency of one permutation), 10 = the same with 4 leaves per
  //         warp and helper lanes when there are at most 4 x 4 leaves per SM (MillerLoop, FP12Mul),
  //     6 = 16 and 7 = 32 leaves per warp for the throughput-bound ones, 9 = one block per SM with the same mix of 16- and
  //         8-leaf warps on every scheduler when the leaves fill that shape (FinalExp / ECCAgg: 32768 leaves on 148 SMs);
  //   round 1 / 2 kernels, still selectable: 13 = one state word per warp, sparse partial rounds with a reducer warp (the
  //     former latency kernel), 12 = its dense two-barrier variant, 4 = three words per thread, four warps per 32 leaves,
  //     dense MDS on dp2a (the former throughput kernel), 3 = the same on IMAD.WIDE, 5 = dp2a + sparse partial rounds.
  // SB_LEAF_KERNEL=1|3|4|5|6|7|8|9|10|12|13 overrides the choice (profiling, tests).
  int kind = 1;
  if (leaf_len > 4 && perms > 2) kind = sb_mm_kind(ctx, n_leaves);
  if (const char* e = getenv("SB_LEAF_KERNEL")) { int v = atoi(e); if (v == 1 || ((v == 3 || v == 4 || v == 5 || v == 6 || v == 7 || v == 8 || v == 9 || v == 10 || v == 12 || v == 13) && leaf_len > 4)) kind = v; }
  const uint32_t groups = (n_leaves + 31) / 32;
  if (kind == 13) {
    // SB_SP_VARIANT: lab variants of the sp kernel (leafhash.cuh); 0 = the round-1 kernel
    static const int spv = [] { const char* e = getenv("SB_SP_VARIANT"); return e ? atoi(e) : 0; }();
    switch (spv) {
      case 1: LAUNCH(ctx, leaf_sponge_sp_kernel<1>, groups, 512, 0, d_cols, leaf_len, n_leaves, log_block, d_digests); break;
      case 2: LAUNCH(ctx, leaf_sponge_sp_kernel<2>, groups, 416, 0, d_cols, leaf_len, n_leaves, log_block, d_digests); break;
      case 3: LAUNCH(ctx, leaf_sponge_sp_kernel<3>, groups, 512, 0, d_cols, leaf_len, n_leaves, log_block, d_digests); break;
      case 4: LAUNCH(ctx, leaf_sponge_sp_kernel<4>, groups, 416, 0, d_cols, leaf_len, n_leaves, log_block, d_digests); break;
      case 6: LAUNCH(ctx, leaf_sponge_sp_kernel<6>, groups, 416, 0, d_cols, leaf_len, n_leaves, log_block, d_digests); break;
      case 7: LAUNCH(ctx, leaf_sponge_sp_kernel<7>, groups, 512, 0, d_cols, leaf_len, n_leaves, log_block, d_digests); break;
      default: LAUNCH(ctx, leaf_sponge_sp_kernel<0>, groups, 416, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
    }
  } else if (kind == 12) {
    LAUNCH(ctx, (leaf_sponge_w12_kernel<0, 1>), groups, 384, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else if (kind == 4) {
    LAUNCH(ctx, leaf_sponge_dp_kernel<0>, groups, 128, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else if (kind == 9) {
    launch_mm_het(ctx, d_cols, leaf_len, n_leaves, log_block, d_digests, nullptr, nullptr);
  } else if (kind == 10) {
    LAUNCH(ctx, (leaf_sponge_mm_kernel<1, 256>), (n_leaves + 3) / 4, 32, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else if (kind == 8) {
    LAUNCH(ctx, leaf_sponge_mm_kernel<1>, (n_leaves + 7) / 8, 32, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else if (kind == 6) {
    LAUNCH(ctx, leaf_sponge_mm_kernel<2>, groups, 64, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else if (kind == 7) {
    LAUNCH(ctx, leaf_sponge_mm_kernel<4>, (n_leaves + 63) / 64, 64, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else if (kind == 5) {
    LAUNCH(ctx, leaf_sponge_ds_kernel, groups, 128, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else if (kind == 3) {
    LAUNCH(ctx, leaf_sponge_ws_kernel<4>, groups, 128, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  } else {
    unsigned block = n_leaves >= 128 * (unsigned)ctx->sm_count ? 128 : (n_leaves >= 64 * (unsigned)ctx->sm_count ? 64 : 32);
    LAUNCH(ctx, leaf_hash_kernel, (n_leaves + block - 1) / block, block, 0, d_cols, leaf_len, n_leaves, log_block, d_digests);
  }
}

// The same sponge fed one slab of columns at a time (capi.cu ingest_and_commit_trace: the trace arrives over PCIe in
// column slabs and the sponge absorbs columns in order, so slab k is hashed while slab k+1 is copied and extended).
// d_state = [12][n_leaves] u64 carried between launches; n_cols % 8 == 0 on every launch but the last; the last launch
// writes the digests.  The matrix-instruction kernels take part (SB_STREAM_OLD=1: round 2's sp / dp kernels).
bool sb_hash_leaves_streamable(uint32_t leaf_len_total) { return leaf_len_total > 64 && !getenv("SB_LEAF_KERNEL"); }
void sb_hash_leaves_stream(sb_ctx* ctx, const u64* d_cols, uint32_t n_cols, uint32_t n_leaves, unsigned log_block, u64* d_state,
                           bool first, bool last, u64* d_digests) {
  if (!last && (n_cols % 8)) SB_THROW(SB_EINVAL, "internal: a %u-column slab inside a streamed leaf sponge", n_cols);
  const uint32_t groups = (n_leaves + 31) / 32;
  const u64* in = first ? nullptr : d_state;
  u64* out = last ? nullptr : d_state;
  const char* e_old = getenv("SB_STREAM_OLD");                     // round-2 kernels (A/B runs, tests)
  const int old = e_old ? atoi(e_old) : 0;
  const int kind = sb_mm_kind(ctx, n_leaves);
  if (old && (kind == 8 || kind == 10))
    LAUNCH(ctx, leaf_sponge_sp_kernel<0>, groups, 416, 0, d_cols, n_cols, n_leaves, log_block, d_digests, in, out);
  else if (old)
    LAUNCH(ctx, leaf_sponge_dp_kernel<0>, groups, 128, 0, d_cols, n_cols, n_leaves, log_block, d_digests, in, out);
  else if (kind == 9)
    launch_mm_het(ctx, d_cols, n_cols, n_leaves, log_block, d_digests, in, out);
  else if (kind == 10)
    LAUNCH(ctx, (leaf_sponge_mm_kernel<1, 256>), (n_leaves + 3) / 4, 32, 0, d_cols, n_cols, n_leaves, log_block, d_digests, in, out);
  else if (kind == 8)
    LAUNCH(ctx, leaf_sponge_mm_kernel<1>, (n_leaves + 7) / 8, 32, 0, d_cols, n_cols, n_leaves, log_block, d_digests, in, out);
  else if (kind == 7)
    LAUNCH(ctx, leaf_sponge_mm_kernel<4>, (n_leaves + 63) / 64, 64, 0, d_cols, n_cols, n_leaves, log_block, d_digests, in, out);
  else
    LAUNCH(ctx, leaf_sponge_mm_kernel<2>, groups, 64, 0, d_cols, n_cols, n_leaves, log_block, d_digests, in, out);
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

// digests in device position order -> plonky2 leaf order (row-sharded leaf hashing gathers position-ordered digests)
__global__ void digests_to_leaf_order_kernel(const ulonglong2* __restrict__ in, ulonglong2* __restrict__ out, uint32_t n_leaves,
                                             unsigned log_block) {
  uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n_leaves) return;
  const size_t d = leaf_index_of(pos, log_block);
  out[2 * d] = in[2 * (size_t)pos];
  out[2 * d + 1] = in[2 * (size_t)pos + 1];
}
void sb_digests_to_leaf_order(sb_ctx* ctx, const u64* d_pos_order, u64* d_leaf_order, uint32_t n_leaves, unsigned log_block) {
  LAUNCH(ctx, digests_to_leaf_order_kernel, (n_leaves + 127) / 128, 128, 0, (const ulonglong2*)d_pos_order, (ulonglong2*)d_leaf_order,
         n_leaves, log_block);
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
