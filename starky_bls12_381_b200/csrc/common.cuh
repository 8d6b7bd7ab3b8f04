// Shared host-side plumbing of libstarkyb200: context, device buffers, error handling, stage timers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/starky_b200.h"
#include "gl.cuh"

struct SbError {
  int code;
  std::string msg;
};

#define SB_THROW(code_, ...)                         \
  do {                                               \
    char _b[512];                                    \
    snprintf(_b, sizeof(_b), __VA_ARGS__);           \
    throw SbError{(code_), std::string(_b)};         \
  } while (0)

#define CUDA_CHECK(x)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (x);                                                                         \
    if (_e != cudaSuccess) SB_THROW(SB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(_e)); \
  } while (0)

// grow-only device buffer (reused across the proofs of one ctx)
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) SB_THROW(SB_ENOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    cap = bytes;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

struct Twiddles {
  unsigned log_size = 0;
  DevBuf fwd;  // w^k,  k < size/2
  DevBuf inv;  // w^-k, k < size/2
};

struct AirProgram;  // quotient.cu
struct sb_multi;    // group.cu: the per-device rank contexts of a multi-GPU ctx

struct StageTimer {
  cudaEvent_t a = nullptr, b = nullptr;
};

struct sb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;    // H2D of trace slabs, overlapped with K1 (capi.cu ingest_and_commit_trace)
  cudaEvent_t slab_ev[4] = {nullptr, nullptr, nullptr, nullptr}, fork_ev = nullptr;
  std::string err;
  uint64_t launches = 0;
  int sm_count = 148;
  std::map<unsigned, Twiddles> tw;                 // per log_size
  std::map<uint64_t, DevBuf> coset_scale;          // key (log_n<<8 | rate_bits): [2^r][n] scale factors
  std::map<uint32_t, AirProgram*> airs;
  std::map<std::string, float> stage_ms;
  std::map<std::string, StageTimer> timers;
  // resident state of the current proof
  sb_params cur = {};
  bool have_trace = false, have_lde = false;
  bool yield_slabs = false;   // set by sb_prove_batch: commit traces of every layout group by group (capi.cu)
  DevBuf trace;        // [C][n] u64 column-major values
  DevBuf staging;      // raw host layout before transposition
  DevBuf coeffs;       // [C][n] coefficients, bit-reversed coefficient order
  DevBuf lde;          // [C][N] values, coset-major order (see ntt.cu)
  DevBuf tree;         // trace Merkle tree: all levels, [sum_l N>>l][4]
  DevBuf qvals;        // [num_challenges][N] quotient values / coefficients
  DevBuf qcoeffs;      // [nq][n]
  DevBuf qlde;         // [nq][N]
  DevBuf qtree;
  DevBuf pis;          // public inputs
  DevBuf weights;      // alpha powers
  DevBuf scratch0, scratch1, scratch2, scratch3, peer_tab;
  DevBuf sponge;       // [12][N] leaf-sponge state between the column slabs of a streamed trace commitment
  void* pinned = nullptr; size_t pinned_cap = 0;
  sb_multi* multi = nullptr;   // sb_init(devices, n > 1): one rank context per device behind this ctx (group.cu)
};

const Twiddles& sb_twiddles(sb_ctx* ctx, unsigned log_size);

static inline void stage_begin(sb_ctx* ctx, const char* name) {
  StageTimer& t = ctx->timers[name];
  if (!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
  cudaEventRecord(t.a, ctx->stream);
}
static inline void stage_end(sb_ctx* ctx, const char* name) {
  StageTimer& t = ctx->timers[name];
  cudaEventRecord(t.b, ctx->stream);
}
// call after a stream synchronize
static inline void stage_collect(sb_ctx* ctx) {
  for (auto& kv : ctx->timers) {
    float ms = 0;
    if (kv.second.a && cudaEventElapsedTime(&ms, kv.second.a, kv.second.b) == cudaSuccess) ctx->stage_ms[kv.first] = ms;
  }
}

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                    \
  do {                                                                 \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);   \
    (ctx)->launches++;                                                 \
    CUDA_CHECK(cudaGetLastError());                                    \
  } while (0)

static inline unsigned ilog2(uint64_t x) { unsigned b = 0; while ((uint64_t(1) << b) < x) b++; return b; }

// ---- stage functions implemented across the .cu files ----
// ntt.cu
void sb_lde_trace(sb_ctx* ctx, const u64* d_values, u64* d_coeffs, u64* d_lde, uint32_t n_cols, unsigned log_n, unsigned rate_bits,
                  unsigned log_row_blocks = 0, u64* const* d_dst_tab = nullptr, uint32_t col0 = 0);
void sb_ntt_device(sb_ctx* ctx, u64* d_data, unsigned log_size, uint32_t count, bool inverse, bool dif);
void sb_transpose_rows_to_cols(sb_ctx* ctx, const void* d_rows, u64* d_cols, uint32_t n_rows, uint32_t n_cols, bool is_u32);
// merkle.cu
void sb_hash_leaves_device(sb_ctx* ctx, const u64* d_cols, uint32_t leaf_len, uint32_t n_leaves, unsigned log_block, u64* d_digests);
void sb_merkle_levels(sb_ctx* ctx, u64* d_tree, uint32_t n_leaves, unsigned cap_height);
bool sb_hash_leaves_streamable(uint32_t leaf_len_total);
void sb_hash_leaves_stream(sb_ctx* ctx, const u64* d_cols, uint32_t n_cols, uint32_t n_leaves, unsigned log_block, u64* d_state,
                           bool first, bool last, u64* d_digests);
void sb_poseidon_permute_device(sb_ctx* ctx, u64* d_states, uint32_t count);
void sb_digests_to_leaf_order(sb_ctx* ctx, const u64* d_pos_order, u64* d_leaf_order, uint32_t n_leaves, unsigned log_block);
