// C ABI of libstarkyb200.so (include/starky_b200.h): lifecycle, parameter / layout helpers, trace ingest and the
// stage-level entry points.  The full proof (sb_prove) lives in prover.cu.
#include <dlfcn.h>
#include <string.h>

#include <thread>

#include "common.cuh"
#include "prover.cuh"

static thread_local std::string g_last_error;

int sb_fail(sb_ctx* ctx, const SbError& e) {
  if (ctx) ctx->err = e.msg;
  g_last_error = e.msg;
  return e.code;
}

#define SB_TRY(ctx) try {
#define SB_CATCH(ctx)                                                            \
  }                                                                              \
  catch (const SbError& e) { return sb_fail((ctx), e); }                         \
  catch (const std::exception& e) { return sb_fail((ctx), SbError{SB_EINVAL, e.what()}); } \
  catch (...) { return sb_fail((ctx), SbError{SB_EINVAL, "unknown failure"}); }

extern "C" {

const char* sb_last_error(sb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int sb_params_standard(uint32_t stark_id, uint32_t log_n, sb_params* out) {
  if (!out) return SB_EINVAL;
  memset(out, 0, sizeof(*out));
  // StarkConfig::standard_fast_config(): security 100, 2 challenges, rate 1, cap 4, pow 16, ConstantArityBits(4,5), 84 queries;
  // rate_bits = 2 for PairingPrecomp / FinalExp / ECCAgg (aggregate_proof.rs:33,156,187)
  out->stark_id = stark_id; out->log_n = log_n; out->cap_height = 4; out->num_challenges = 2; out->pow_bits = 16;
  out->num_query_rounds = 84; out->fri_arity_bits = 4; out->fri_final_poly_bits = 5; out->rate_bits = 1;
  switch (stark_id) {
    case SB_STARK_FP12_MUL: out->n_cols = 60285; out->n_public_inputs = 432; out->constraint_degree = 3; break;
    case SB_STARK_PAIRING_PRECOMP: out->n_cols = 29376; out->n_public_inputs = 4968; out->constraint_degree = 4; out->rate_bits = 2; break;
    case SB_STARK_MILLER_LOOP: out->n_cols = 97330; out->n_public_inputs = 5064; out->constraint_degree = 3; break;
    case SB_STARK_FINAL_EXP: out->n_cols = 73527; out->n_public_inputs = 288; out->constraint_degree = 5; out->rate_bits = 2; break;
    case SB_STARK_ECC_AGG: out->n_cols = 3339; out->n_public_inputs = 12824; out->constraint_degree = 4; out->rate_bits = 2; break;
    default: break;  // custom AIR: caller fills n_cols / n_public_inputs / constraint_degree / rate_bits
  }
  return SB_OK;
}

int sb_proof_layout_for(const sb_params* p, sb_proof_layout* out) {
  if (!p || !out) return SB_EINVAL;
  try { *out = proof_layout(*p); } catch (const SbError& e) { return sb_fail(nullptr, e); }
  return SB_OK;
}
uint32_t sb_fri_step_path_len(const sb_proof_layout* l, uint32_t round) { return fri_step_path_len(*l, round); }
uint64_t sb_fri_step_offset(const sb_proof_layout* l, uint32_t round) { return fri_step_offset(*l, round); }

int sb_init(const int* devices, int n_devices, sb_ctx** out) {
  if (!out) return SB_EINVAL;
  sb_ctx* ctx = new sb_ctx();
  SB_TRY(ctx)
  int dev = 0;
  if (devices && n_devices > 0) dev = devices[0];
  else CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaSetDevice(dev));
  ctx->device = dev;
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  ctx->sm_count = prop.multiProcessorCount;
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  // several devices: one rank context per device behind this ctx; sb_prove shards the trace over them (group.cu)
  if (devices && n_devices > 1) multi_init(ctx, devices, n_devices);
  *out = ctx;
  return SB_OK;
  }
  catch (const SbError& e) { int rc = sb_fail(nullptr, e); if (ctx->stream) cudaStreamDestroy(ctx->stream); delete ctx; return rc; }
}

void sb_destroy(sb_ctx* ctx) {
  if (!ctx) return;
  multi_destroy(ctx);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->tw) { kv.second.fwd.release(); kv.second.inv.release(); }
  for (auto& kv : ctx->coset_scale) kv.second.release();
  for (auto& kv : ctx->timers) { if (kv.second.a) cudaEventDestroy(kv.second.a); if (kv.second.b) cudaEventDestroy(kv.second.b); }
  DevBuf* bufs[] = {&ctx->trace, &ctx->staging, &ctx->coeffs, &ctx->lde, &ctx->tree, &ctx->qvals, &ctx->qcoeffs, &ctx->qlde,
                    &ctx->qtree, &ctx->pis, &ctx->weights, &ctx->scratch0, &ctx->scratch1, &ctx->scratch2, &ctx->scratch3, &ctx->peer_tab, &ctx->sponge};
  for (DevBuf* b : bufs) b->release();
  air_release_all(ctx);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    for (auto& e : ctx->slab_ev) cudaEventDestroy(e);
    cudaEventDestroy(ctx->fork_ev);
    cudaStreamDestroy(ctx->copy_stream);
  }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

uint64_t sb_kernel_launches(sb_ctx* ctx) { return ctx ? ctx->launches : 0; }

float sb_stage_ms(sb_ctx* ctx, const char* stage) {
  if (!ctx || !stage) return -1.f;
  auto it = ctx->stage_ms.find(stage);
  return it == ctx->stage_ms.end() ? -1.f : it->second;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// trace ingest: any accepted host layout -> ctx->trace, column-major [C][n] u64 on the device
// ---------------------------------------------------------------------------------------------------------
static void* pinned_staging(sb_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_cap) return ctx->pinned;
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr; ctx->pinned_cap = 0;
  cudaError_t e = cudaMallocHost(&ctx->pinned, bytes);
  if (e != cudaSuccess) SB_THROW(SB_ENOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
  ctx->pinned_cap = bytes;
  return ctx->pinned;
}

const u64* ingest_trace(sb_ctx* ctx, const sb_params* p, const void* trace, int layout) {
  const size_t n = size_t(1) << p->log_n, C = p->n_cols;
  if (layout == SB_TRACE_DEVICE_COLMAJOR_U64) {
    if (trace) return (const u64*)trace;
    if (!ctx->have_trace) SB_THROW(SB_EINVAL, "no resident trace: call sb_trace_upload first");
    return ctx->trace.as<u64>();
  }
  if (!trace) SB_THROW(SB_EINVAL, "trace is NULL");
  ctx->trace.ensure(8 * n * C);
  switch (layout) {
    case SB_TRACE_COLMAJOR_U64:
      CUDA_CHECK(cudaMemcpyAsync(ctx->trace.p, trace, 8 * n * C, cudaMemcpyHostToDevice, ctx->stream));
      break;
    case SB_TRACE_COLS_U64_PTRS: {
      // Vec<PolynomialValues<F>>: C separately allocated columns.  Gather into pinned staging in slabs so that the
      // host gather of slab k+1 overlaps the H2D copy of slab k.
      const u64* const* cols = (const u64* const*)trace;
      const size_t slab_cols = std::max<size_t>(1, (64u << 20) / (8 * n));
      u64* stage = (u64*)pinned_staging(ctx, 2 * slab_cols * 8 * n);
      cudaEvent_t done[2]; cudaEventCreate(&done[0]); cudaEventCreate(&done[1]);
      int which = 0;
      for (size_t c0 = 0; c0 < C; c0 += slab_cols, which ^= 1) {
        size_t cnt = std::min(slab_cols, C - c0);
        u64* dst = stage + which * slab_cols * n;
        if (c0 >= 2 * slab_cols) cudaEventSynchronize(done[which]);
        for (size_t c = 0; c < cnt; c++) memcpy(dst + c * n, cols[c0 + c], 8 * n);
        CUDA_CHECK(cudaMemcpyAsync(ctx->trace.as<u64>() + c0 * n, dst, 8 * n * cnt, cudaMemcpyHostToDevice, ctx->stream));
        cudaEventRecord(done[which], ctx->stream);
      }
      cudaEventSynchronize(done[0]); cudaEventSynchronize(done[1]);
      cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
      break;
    }
    case SB_TRACE_ROWMAJOR_U64:
    case SB_TRACE_ROWMAJOR_U32: {
      size_t esz = layout == SB_TRACE_ROWMAJOR_U32 ? 4 : 8;
      ctx->staging.ensure(esz * n * C);
      CUDA_CHECK(cudaMemcpyAsync(ctx->staging.p, trace, esz * n * C, cudaMemcpyHostToDevice, ctx->stream));
      sb_transpose_rows_to_cols(ctx, ctx->staging.p, ctx->trace.as<u64>(), (uint32_t)n, (uint32_t)C, esz == 4);
      break;
    }
    default: SB_THROW(SB_EINVAL, "unknown trace layout %d", layout);
  }
  ctx->have_trace = true;
  return ctx->trace.as<u64>();
}

// The one parameter check of every entry point (sb_prove, sb_prove_sharded, the stage functions, the layout helper).
void check_params(const sb_params* p) {
  if (!p) SB_THROW(SB_EINVAL, "params is NULL");
  if (p->n_cols == 0 || p->n_cols >= (1u << 18)) SB_THROW(SB_EINVAL, "n_cols %u out of range [1, 2^18)", p->n_cols);
  if (p->log_n < 1 || p->log_n > 13) SB_THROW(SB_EINVAL, "log_n %u out of range [1,13]", p->log_n);
  if (p->rate_bits < 1 || p->rate_bits > 4) SB_THROW(SB_EINVAL, "rate_bits %u out of range [1,4]", p->rate_bits);
  if (p->cap_height > p->log_n + p->rate_bits) SB_THROW(SB_EINVAL, "cap_height %u larger than the tree", p->cap_height);
  if (p->num_challenges < 1 || p->num_challenges > 4) SB_THROW(SB_EINVAL, "num_challenges %u out of range [1,4]", p->num_challenges);
  if (p->constraint_degree < 1 || p->constraint_degree > 17) SB_THROW(SB_EINVAL, "constraint_degree %u out of range [1,17]", p->constraint_degree);
  if (p->pow_bits > 32) SB_THROW(SB_EINVAL, "pow_bits %u out of range [0,32]", p->pow_bits);
  if (p->num_query_rounds < 1 || p->num_query_rounds > 1024) SB_THROW(SB_EINVAL, "num_query_rounds %u out of range [1,1024]", p->num_query_rounds);
  if (p->fri_arity_bits < 1 || p->fri_arity_bits > 9) SB_THROW(SB_EINVAL, "fri_arity_bits %u out of range [1,9]", p->fri_arity_bits);
  if (p->fri_final_poly_bits > 13) SB_THROW(SB_EINVAL, "fri_final_poly_bits %u out of range [0,13]", p->fri_final_poly_bits);
  (void)fri_arities(*p);   // throws SB_EINVAL where plonky2's reduction strategy would assert
}

// from_values on the device: ctx->trace -> ctx->coeffs, ctx->lde, ctx->tree.
void commit_trace(sb_ctx* ctx, const sb_params* p, const u64* d_values) {
  const size_t n = size_t(1) << p->log_n, N = n << p->rate_bits, C = p->n_cols;
  ctx->coeffs.ensure(8 * n * C);
  ctx->lde.ensure(8 * N * C);
  ctx->tree.ensure(32 * 2 * N);
  stage_begin(ctx, "lde");
  sb_lde_trace(ctx, d_values, ctx->coeffs.as<u64>(), ctx->lde.as<u64>(), (uint32_t)C, p->log_n, p->rate_bits);
  stage_end(ctx, "lde");
  stage_begin(ctx, "leaf_hash");
  sb_hash_leaves_device(ctx, ctx->lde.as<u64>(), (uint32_t)C, (uint32_t)N, p->log_n, ctx->tree.as<u64>());
  stage_end(ctx, "leaf_hash");
  stage_begin(ctx, "merkle");
  sb_merkle_levels(ctx, ctx->tree.as<u64>(), (uint32_t)N, p->cap_height);
  stage_end(ctx, "merkle");
  ctx->cur = *p;
  ctx->have_lde = true;
}

// Vec<PolynomialValues<F>> -> one contiguous (pinned) slab.  One host core copies ~10 GB/s, PCIe takes ~55: four helper
// threads per slab keep the gather ahead of the H2D copy (FinalExp, 4.8 GB in 73 527 pageable columns: 1115 -> ~760 ms).
static void gather_columns(u64* dst, const u64* const* cols, size_t cnt, size_t n) {
  const size_t bytes = 8 * n * cnt;
  const unsigned workers = bytes >= (4u << 20) ? 4 : 1;
  if (workers == 1) { for (size_t c = 0; c < cnt; c++) memcpy(dst + c * n, cols[c], 8 * n); return; }
  std::vector<std::thread> th;
  for (unsigned w = 0; w < workers; w++)
    th.emplace_back([=] {
      const size_t a = cnt * w / workers, b = cnt * (w + 1) / workers;
      for (size_t c = a; c < b; c++) memcpy(dst + c * n, cols[c], 8 * n);
    });
  for (auto& t : th) t.join();
}

// Ingest + from_values in one pipeline for the host column layouts: the trace crosses PCIe in column slabs on a second
// stream while K1 already extends the slabs that have arrived, so the H2D copy (the larger of the two for every stark:
// 8 C n bytes at ~55 GB/s) hides the whole LDE.  `h2d_done` is recorded on the copy stream after the last slab.
void ingest_and_commit_trace(sb_ctx* ctx, const sb_params* p, const void* trace, int layout, cudaEvent_t h2d_done) {
  const size_t n = size_t(1) << p->log_n, N = n << p->rate_bits, C = p->n_cols;
  const bool pipelined = trace && (layout == SB_TRACE_COLMAJOR_U64 || layout == SB_TRACE_COLS_U64_PTRS);
  if (!pipelined) {
    const u64* d_values = ingest_trace(ctx, p, trace, layout);
    if (h2d_done) CUDA_CHECK(cudaEventRecord(h2d_done, ctx->stream));
    // The other layouts (row-major rows, device-resident columns) inside sb_prove_batch (SB_YIELD_SLABS, set by the
    // scheduler): K1 and K2 alternate over column groups as in the pipelined path below -- the same work in more launches,
    // but a throughput-bound proof then yields the SMs at every launch boundary instead of holding them with one
    // 350 ms leaf-sponge launch, and the latency-bound proofs next to it keep running.
    const size_t group = 2048;
    if (ctx->yield_slabs && C > 2 * group && sb_hash_leaves_streamable((uint32_t)C) && !getenv("SB_NO_STREAM_HASH")) {
      ctx->coeffs.ensure(8 * n * C);
      ctx->lde.ensure(8 * N * C);
      ctx->tree.ensure(32 * 2 * N);
      ctx->sponge.ensure(8ull * 12 * N);
      stage_begin(ctx, "lde");
      stage_begin(ctx, "leaf_hash");                                   // (the two stage timers overlap in this path)
      for (size_t c0 = 0; c0 < C; c0 += group) {
        const size_t cnt = std::min(group, C - c0);
        const bool last = c0 + cnt == C;
        sb_lde_trace(ctx, d_values + c0 * n, ctx->coeffs.as<u64>() + c0 * n, ctx->lde.as<u64>() + c0 * N, (uint32_t)cnt, p->log_n, p->rate_bits);
        if (last) stage_end(ctx, "lde");
        sb_hash_leaves_stream(ctx, ctx->lde.as<u64>() + c0 * N, (uint32_t)cnt, (uint32_t)N, p->log_n, ctx->sponge.as<u64>(), c0 == 0, last,
                              ctx->tree.as<u64>());
      }
      stage_end(ctx, "leaf_hash");
      stage_begin(ctx, "merkle");
      sb_merkle_levels(ctx, ctx->tree.as<u64>(), (uint32_t)N, p->cap_height);
      stage_end(ctx, "merkle");
      ctx->cur = *p;
      ctx->have_lde = true;
      return;
    }
    commit_trace(ctx, p, d_values);
    return;
  }
  if (!ctx->copy_stream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->slab_ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  }
  ctx->trace.ensure(8 * n * C);
  ctx->coeffs.ensure(8 * n * C);
  ctx->lde.ensure(8 * N * C);
  ctx->tree.ensure(32 * 2 * N);
  // SB_SLAB_BYTES: tests shrink the slabs so that small traces take the multi-slab path too
  size_t slab_bytes = 32u << 20;
  if (const char* e = getenv("SB_SLAB_BYTES")) slab_bytes = std::max<size_t>(64 * n, strtoull(e, nullptr, 10));
  const size_t slab_cols = std::max<size_t>(8, slab_bytes / (8 * n) / 8 * 8);
  // leaf hashing follows the slabs: the sponge absorbs columns in order, so a group of slabs is hashed as soon as it is
  // extended, while the next ones cross PCIe -- the copy is hidden behind K2 as well as K1.  hash_group slabs per launch
  // (>= 2048 columns, SB_HASH_GROUP_COLS: the state save / restore and the launch are noise).  SB_NO_STREAM_HASH=1: hash
  // after the last slab.
  const bool stream_hash = C > 2 * slab_cols && sb_hash_leaves_streamable((uint32_t)C) && !getenv("SB_NO_STREAM_HASH");
  size_t group_cols = 2048;
  if (const char* e = getenv("SB_HASH_GROUP_COLS")) group_cols = std::max<size_t>(8, strtoull(e, nullptr, 10));
  const size_t hash_group = std::max<size_t>(1, group_cols / slab_cols);
  if (stream_hash) ctx->sponge.ensure(8ull * 12 * N);
  u64* stage = nullptr;
  if (layout == SB_TRACE_COLS_U64_PTRS) stage = (u64*)pinned_staging(ctx, 2 * slab_cols * 8 * n);
  // the copy stream starts where the main stream is now (earlier work may still be using ctx->trace)
  CUDA_CHECK(cudaEventRecord(ctx->fork_ev, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->fork_ev, 0));
  stage_begin(ctx, "lde");
  size_t i = 0, hashed = 0;          // columns [0, hashed) are in the sponge state
  const size_t n_slabs = (C + slab_cols - 1) / slab_cols;
  for (size_t c0 = 0; c0 < C; c0 += slab_cols, i++) {
    const size_t cnt = std::min(slab_cols, C - c0);
    cudaEvent_t ev = ctx->slab_ev[i % 4];
    u64* d_slab = ctx->trace.as<u64>() + c0 * n;
    if (layout == SB_TRACE_COLMAJOR_U64) {
      CUDA_CHECK(cudaMemcpyAsync(d_slab, (const u64*)trace + c0 * n, 8 * n * cnt, cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {
      // Vec<PolynomialValues<F>>: gather the separately allocated columns into one of two pinned staging slabs
      const u64* const* cols = (const u64* const*)trace;
      u64* dst = stage + (i & 1) * slab_cols * n;
      if (i >= 2) CUDA_CHECK(cudaEventSynchronize(ctx->slab_ev[(i - 2) % 4]));   // that staging slab has left the host
      gather_columns(dst, cols + c0, cnt, n);
      CUDA_CHECK(cudaMemcpyAsync(d_slab, dst, 8 * n * cnt, cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    // (re-recording an event does not disturb a cudaStreamWaitEvent already enqueued on its previous record)
    CUDA_CHECK(cudaEventRecord(ev, ctx->copy_stream));
    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ev, 0));
    sb_lde_trace(ctx, d_slab, ctx->coeffs.as<u64>() + c0 * n, ctx->lde.as<u64>() + c0 * N, (uint32_t)cnt, p->log_n, p->rate_bits);
    if (i + 1 == n_slabs) stage_end(ctx, "lde");
    if (stream_hash && ((i + 1) % hash_group == 0 || i + 1 == n_slabs)) {
      if (hashed == 0) stage_begin(ctx, "leaf_hash");          // (in this path the two stage timers overlap)
      sb_hash_leaves_stream(ctx, ctx->lde.as<u64>() + hashed * N, (uint32_t)(c0 + cnt - hashed), (uint32_t)N, p->log_n,
                            ctx->sponge.as<u64>(), hashed == 0, i + 1 == n_slabs, ctx->tree.as<u64>());
      hashed = c0 + cnt;
    }
  }
  if (h2d_done) CUDA_CHECK(cudaEventRecord(h2d_done, ctx->copy_stream));
  ctx->have_trace = true;
  if (!stream_hash) {
    stage_begin(ctx, "leaf_hash");
    sb_hash_leaves_device(ctx, ctx->lde.as<u64>(), (uint32_t)C, (uint32_t)N, p->log_n, ctx->tree.as<u64>());
  }
  stage_end(ctx, "leaf_hash");
  stage_begin(ctx, "merkle");
  sb_merkle_levels(ctx, ctx->tree.as<u64>(), (uint32_t)N, p->cap_height);
  stage_end(ctx, "merkle");
  ctx->cur = *p;
  ctx->have_lde = true;
}

const u64* tree_cap_ptr(const u64* d_tree, size_t n_leaves, unsigned cap_height) {
  // level l starts at digest offset 2N - (2N >> l); the cap is level log2(N) - cap_height
  size_t off = 2 * n_leaves - (size_t(2) << cap_height);
  return d_tree + 4 * off;
}

extern "C" {

int sb_trace_upload(sb_ctx* ctx, const sb_params* p, const void* trace, int layout) {
  if (!ctx) return SB_EINVAL;
  SB_TRY(ctx)
  check_params(p);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  const u64* d = ingest_trace(ctx, p, trace, layout);
  if (d != ctx->trace.as<u64>()) {  // device pointer supplied: keep a resident copy
    size_t bytes = (8ull << p->log_n) * p->n_cols;
    ctx->trace.ensure(bytes);
    CUDA_CHECK(cudaMemcpyAsync(ctx->trace.p, d, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->have_trace = true;
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
  SB_CATCH(ctx)
}

int sb_lde_commit(sb_ctx* ctx, const sb_params* p, const void* trace, int layout, uint64_t* lde_out,
                  uint64_t* digests_out, uint64_t* cap_out) {
  if (!ctx) return SB_EINVAL;
  SB_TRY(ctx)
  check_params(p);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  const size_t n = size_t(1) << p->log_n, N = n << p->rate_bits, C = p->n_cols;
  ingest_and_commit_trace(ctx, p, trace, layout, nullptr);
  if (lde_out) CUDA_CHECK(cudaMemcpyAsync(lde_out, ctx->lde.p, 8 * N * C, cudaMemcpyDeviceToHost, ctx->stream));
  if (digests_out) CUDA_CHECK(cudaMemcpyAsync(digests_out, ctx->tree.p, 32 * N, cudaMemcpyDeviceToHost, ctx->stream));
  if (cap_out)
    CUDA_CHECK(cudaMemcpyAsync(cap_out, tree_cap_ptr(ctx->tree.as<u64>(), N, p->cap_height), 32ull << p->cap_height,
                               cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  stage_collect(ctx);
  return SB_OK;
  SB_CATCH(ctx)
}

int sb_coeffs_download(sb_ctx* ctx, uint64_t* coeffs_out) {
  if (!ctx || !coeffs_out) return SB_EINVAL;
  SB_TRY(ctx)
  if (!ctx->have_lde) SB_THROW(SB_EINVAL, "no committed trace on this ctx");
  size_t bytes = (8ull << ctx->cur.log_n) * ctx->cur.n_cols;
  CUDA_CHECK(cudaMemcpyAsync(coeffs_out, ctx->coeffs.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
  SB_CATCH(ctx)
}

// ---- multi-GPU stage entry points: device pointers in and out; the host does the exchange with NCCL ----
int sb_synchronize(sb_ctx* ctx) {
  if (!ctx) return SB_EINVAL;
  SB_TRY(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  stage_collect(ctx);
  return SB_OK;
  SB_CATCH(ctx)
}

int sb_lde_cols_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_trace, uint32_t n_cols_local, uint32_t n_row_blocks,
                       uint64_t* d_coeffs_out, uint64_t* d_lde_out) {
  if (!ctx || !d_trace || !d_lde_out) return SB_EINVAL;
  SB_TRY(ctx)
  check_params(p);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  unsigned log_blocks = ilog2(n_row_blocks);
  if (n_row_blocks == 0 || (1u << log_blocks) != n_row_blocks || p->log_n + p->rate_bits < log_blocks + 5)
    SB_THROW(SB_EINVAL, "n_row_blocks %u must be a power of two leaving >= 32 positions per block", n_row_blocks);
  if (n_cols_local == 0) return SB_OK;
  stage_begin(ctx, "lde");
  sb_lde_trace(ctx, d_trace, d_coeffs_out, d_lde_out, n_cols_local, p->log_n, p->rate_bits, log_blocks);
  stage_end(ctx, "lde");
  return SB_OK;
  SB_CATCH(ctx)
}

// K1 fused with the exchange: the LDE of this rank's column slice is stored straight into the row buffers of the ranks
// that own the row blocks (peer memory over NVLink for the other ranks' blocks), peer_rows[b] = device pointer of rank b's
// [n_cols_total][N / n_row_blocks] buffer.  No all-to-all pass follows; the caller barriers before reading its rows.
int sb_lde_cols_peer_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_trace, uint32_t n_cols_local, uint32_t n_row_blocks,
                            uint32_t first_col, uint64_t* d_coeffs_out, const uint64_t* peer_rows) {
  if (!ctx || !d_trace || !peer_rows) return SB_EINVAL;
  SB_TRY(ctx)
  check_params(p);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  unsigned log_blocks = ilog2(n_row_blocks);
  if (n_row_blocks == 0 || n_row_blocks > 64 || (1u << log_blocks) != n_row_blocks || p->log_n + p->rate_bits < log_blocks + 5)
    SB_THROW(SB_EINVAL, "n_row_blocks %u must be a power of two <= 64 leaving >= 32 positions per block", n_row_blocks);
  if ((uint64_t)first_col + n_cols_local > p->n_cols) SB_THROW(SB_EINVAL, "column slice [%u, +%u) outside %u columns", first_col, n_cols_local, p->n_cols);
  if (n_cols_local == 0) return SB_OK;
  ctx->peer_tab.ensure(8ull * 64);
  CUDA_CHECK(cudaMemcpyAsync(ctx->peer_tab.p, peer_rows, 8ull * n_row_blocks, cudaMemcpyHostToDevice, ctx->stream));
  stage_begin(ctx, "lde");
  sb_lde_trace(ctx, d_trace, d_coeffs_out, nullptr, n_cols_local, p->log_n, p->rate_bits, log_blocks, (u64* const*)ctx->peer_tab.p, first_col);
  stage_end(ctx, "lde");
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  stage_collect(ctx);
  return SB_OK;
  SB_CATCH(ctx)
}

int sb_hash_rows_device(sb_ctx* ctx, const uint64_t* d_cols, uint32_t leaf_len, uint32_t n_leaves, uint64_t* d_digests) {
  if (!ctx || !d_cols || !d_digests || !leaf_len || !n_leaves) return SB_EINVAL;
  SB_TRY(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  stage_begin(ctx, "leaf_hash");
  sb_hash_leaves_device(ctx, d_cols, leaf_len, n_leaves, 0, d_digests);   // log_block 0: digests stay in position order
  stage_end(ctx, "leaf_hash");
  return SB_OK;
  SB_CATCH(ctx)
}

int sb_merkle_from_position_digests(sb_ctx* ctx, const sb_params* p, const uint64_t* d_digests_pos, uint64_t* cap_out) {
  if (!ctx || !d_digests_pos || !cap_out) return SB_EINVAL;
  SB_TRY(ctx)
  check_params(p);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  const size_t N = size_t(1) << (p->log_n + p->rate_bits);
  ctx->tree.ensure(32 * 2 * N);
  stage_begin(ctx, "merkle");
  sb_digests_to_leaf_order(ctx, d_digests_pos, ctx->tree.as<u64>(), (uint32_t)N, p->log_n);
  sb_merkle_levels(ctx, ctx->tree.as<u64>(), (uint32_t)N, p->cap_height);
  stage_end(ctx, "merkle");
  CUDA_CHECK(cudaMemcpyAsync(cap_out, tree_cap_ptr(ctx->tree.as<u64>(), N, p->cap_height), 32ull << p->cap_height,
                             cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  stage_collect(ctx);
  return SB_OK;
  SB_CATCH(ctx)
}

int sb_ntt_batch(sb_ctx* ctx, uint64_t* data, uint32_t log_n, uint32_t count, int inverse) {
  if (!ctx || !data) return SB_EINVAL;
  SB_TRY(ctx)
  if (log_n > 24) SB_THROW(SB_EINVAL, "log_n too large");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  size_t bytes = (8ull << log_n) * count;
  ctx->scratch0.ensure(bytes);
  ctx->scratch1.ensure(bytes);
  u64* d = ctx->scratch0.as<u64>();
  CUDA_CHECK(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  // natural -> (DIF) bit-reversed -> explicit permutation back to natural order
  sb_ntt_device(ctx, d, log_n, count, inverse != 0, true);
  sb_bitrev_permute_device(ctx, d, ctx->scratch1.as<u64>(), log_n, count);
  CUDA_CHECK(cudaMemcpyAsync(data, ctx->scratch1.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
  SB_CATCH(ctx)
}

int sb_poseidon_permute_batch(sb_ctx* ctx, uint64_t* states, uint32_t count) {
  if (!ctx || !states) return SB_EINVAL;
  SB_TRY(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  size_t bytes = 96ull * count;
  ctx->scratch0.ensure(bytes);
  CUDA_CHECK(cudaMemcpyAsync(ctx->scratch0.p, states, bytes, cudaMemcpyHostToDevice, ctx->stream));
  sb_poseidon_permute_device(ctx, ctx->scratch0.as<u64>(), count);
  CUDA_CHECK(cudaMemcpyAsync(states, ctx->scratch0.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
  SB_CATCH(ctx)
}

int sb_hash_leaves(sb_ctx* ctx, const uint64_t* cols, uint32_t leaf_len, uint32_t count, uint64_t* digests_out) {
  if (!ctx || !cols || !digests_out) return SB_EINVAL;
  SB_TRY(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  size_t bytes = 8ull * leaf_len * count;
  ctx->scratch0.ensure(bytes);
  ctx->scratch1.ensure(32ull * count);
  CUDA_CHECK(cudaMemcpyAsync(ctx->scratch0.p, cols, bytes, cudaMemcpyHostToDevice, ctx->stream));
  sb_hash_leaves_device(ctx, ctx->scratch0.as<u64>(), leaf_len, count, 0, ctx->scratch1.as<u64>());
  CUDA_CHECK(cudaMemcpyAsync(digests_out, ctx->scratch1.p, 32ull * count, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
  SB_CATCH(ctx)
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// integer-pipe roofline denominator: dependent-free u32 multiply-add throughput
// ---------------------------------------------------------------------------------------------------------
template <bool WIDE>
__global__ void imad_peak_kernel(uint32_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t x[8];
  u64 y[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (WIDE) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[i]) : "r"(a), "r"((uint32_t)(b + k)));
        else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= x[i] ^ (uint32_t)y[i] ^ (uint32_t)(y[i] >> 32);
  if (acc == 0xdeadbeef) out[0] = acc;
}

extern "C" int sb_measure_imad_peak(sb_ctx* ctx, double* gops_out) {
  if (!ctx || !gops_out) return SB_EINVAL;
  SB_TRY(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->scratch0.ensure(64);
  const int iters = 4096, block = 256, grid = ctx->sm_count * 8;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int wide = 0; wide < 2; wide++) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(a, ctx->stream);
      if (wide) { LAUNCH(ctx, imad_peak_kernel<true>, grid, block, 0, ctx->scratch0.as<uint32_t>(), 3u, 5u, iters); }
      else { LAUNCH(ctx, imad_peak_kernel<false>, grid, block, 0, ctx->scratch0.as<uint32_t>(), 3u, 5u, iters); }
      cudaEventRecord(b, ctx->stream);
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && ms < best) best = ms;
    }
    double ops = (double)grid * block * iters * 64.0;
    gops_out[wide] = ops / (best * 1e-3) / 1e9;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return SB_OK;
  SB_CATCH(ctx)
}
