// Multi-GPU proofs behind the C ABI (include/starky_b200.h "multi-GPU groups").
//
// One proof of starky::prover::prove (/root/reference/src/aggregate_proof.rs:59,105,138,169,212) with the trace sharded
// over the GPUs of one box:
//   phase 1  column-sharded   rank g owns columns [c0_g, c0_g + C_g): K1 (iNTT + coset LDE) on its slice; every LDE value
//                             is stored straight into the row buffer of the rank that owns its row block (peer memory over
//                             NVLink: CUDA IPC mappings between processes, direct peer access inside one process), so the
//                             column->row exchange overlaps the transform and no all-to-all pass follows
//   phase 2  row-sharded      rank g holds all C columns of its N / G positions: K2 leaf sponge and K4 quotient, row-local
//   small collectives         leaf digests (all-gather, N x 32 B), halo rows (C x 8 B per rank), quotient values (2 x N x 8 B),
//                             openings (2 C extension values), FRI combine partials (n extension values per rank, added
//                             mod p on the device), the 84 query rows from their owners
//   everything else           (quotient commitment, transcript, FRI rounds, proof of work, Merkle paths) is small and runs
//                             redundantly on every rank: prover.cu's prove_impl is the ONE orchestration, the five
//                             distributed steps are its hooks, implemented here.
// Two transports implement the same `Comm` interface:
//   NcclComm   one process per GPU (torchrun, MPI, a Rust host with one process per device): NCCL for the collectives,
//              loaded with dlopen("libnccl.so.2") so that a single-GPU host needs no NCCL at all
//   LocalComm  one process driving several GPUs from one thread each (sb_init(devices, n > 1), the Rust shim's
//              GpuProver::new_multi): peer copies + host barriers, no NCCL.  Ranks may share a device (tests).
#include <dlfcn.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "prover.cuh"

// stage functions of the other translation units
struct AirProgram;
void sb_quotient_rows(sb_ctx* ctx, const sb_params* p, const u64* d_rows, size_t stride, uint32_t n_local, uint32_t pos0,
                      const u64* d_halo, const u64* d_pis, const u64* alphas, u64* d_out);
void sb_openings_device(sb_ctx* ctx, const u64* d_coeffs, unsigned log_n, uint32_t n_polys, e2_t za, const e2_t* zb,
                        e2_t* d_tab_a, e2_t* d_tab_b, e2_t* d_out_a, e2_t* d_out_b);
void sb_combine_device(sb_ctx* ctx, const u64* d_coeffs, unsigned log_n, uint32_t n_polys, e2_t alpha, uint32_t j0,
                       e2_t* d_apow, e2_t* d_partial, size_t partial_capacity_elems, e2_t* d_out);


// ---------------------------------------------------------------------------------------------------------
// NCCL through dlopen: only the handful of entry points used here
// ---------------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclInt32 = 2, ncclUint64 = 5 };
enum { ncclSum = 0 };
struct NcclApi {
  void* so = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi* nccl_api() {
  static std::mutex mu;
  static NcclApi api;
  std::lock_guard<std::mutex> lk(mu);
  if (api.so) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* so = nullptr;
  for (const char* nm : names) if ((so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!so) SB_THROW(SB_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                        \
  *(void**)(&api.field) = dlsym(so, name);                                      \
  if (!api.field) SB_THROW(SB_ENCCL, "libnccl has no symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather") SYM(AllReduce, "ncclAllReduce") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  api.so = so;
  return &api;
}
#define NCCL_CHECK(x)                                                                                       \
  do {                                                                                                      \
    int _r = (x);                                                                                           \
    if (_r != ncclSuccess) SB_THROW(SB_ENCCL, "%s:%d %s: %s", __FILE__, __LINE__, #x, nccl_api()->GetErrorString(_r)); \
  } while (0)
}  // namespace

// ---------------------------------------------------------------------------------------------------------
// shard plan (the same arithmetic as starky_bls12_381_b200/sharded.py: shard_plan)
// ---------------------------------------------------------------------------------------------------------
struct ShardPlan {
  uint32_t n_cols = 0, world = 1, rows = 0;   // rows = LDE positions per rank
  std::vector<uint32_t> col0, cols;
};
static ShardPlan make_plan(const sb_params* p, uint32_t world) {
  const uint32_t N = 1u << (p->log_n + p->rate_bits);
  if (world == 0 || (world & (world - 1))) SB_THROW(SB_EINVAL, "group size %u: the row blocks are power-of-two slices of the LDE domain", world);
  if (N / world < 32) SB_THROW(SB_EINVAL, "fewer than 32 LDE positions per rank (%u positions, %u ranks)", N, world);
  ShardPlan s;
  s.n_cols = p->n_cols; s.world = world; s.rows = N / world;
  const uint32_t base = p->n_cols / world, extra = p->n_cols % world;
  uint32_t c = 0;
  for (uint32_t g = 0; g < world; g++) { s.col0.push_back(c); s.cols.push_back(base + (g < extra ? 1 : 0)); c += s.cols.back(); }
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// transports
// ---------------------------------------------------------------------------------------------------------
struct Comm {
  int rank = 0, world = 1;
  bool peer_ok = true;          // false: no peer mapping between the ranks' GPUs -> the all-to-all path is used
  virtual ~Comm() {}
  // recv = [world][bytes]; send may not alias recv
  virtual void all_gather(sb_ctx* ctx, const void* d_send, void* d_recv, size_t bytes) = 0;
  // on return every rank has passed this point and all earlier work on every rank's stream is complete
  virtual void barrier(sb_ctx* ctx) = 0;
  // slab b of every rank -> rank b.  d_send = [world][cols[rank]][rows], d_recv = [n_cols][rows]
  virtual void all_to_all_slabs(sb_ctx* ctx, const u64* d_send, u64* d_recv, const ShardPlan& plan) = 0;
  // d_rows = [count][n_cols]: row q is valid on rank owner[q] only; on return it is valid everywhere
  virtual void share_rows(sb_ctx* ctx, u64* d_rows, const uint32_t* owner, uint32_t count, uint32_t n_cols) = 0;
  // peer-usable device pointers of every rank's `mine` buffer (same size everywhere); collective
  virtual void exchange_ptrs(sb_ctx* ctx, void* mine, size_t bytes, std::vector<void*>& all) = 0;
  virtual void release_ptrs() {}
  virtual void abort() {}       // this rank failed: wake peers that wait for it (where the transport can)
};

struct NcclComm : Comm {
  NcclApi* api = nullptr;
  ncclComm_t comm = nullptr;
  DevBuf bounce;
  std::vector<void*> opened;
  ~NcclComm() override {
    release_ptrs();
    if (comm) api->CommDestroy(comm);
    bounce.release();
  }
  void all_gather(sb_ctx* ctx, const void* d_send, void* d_recv, size_t bytes) override {
    NCCL_CHECK(api->AllGather(d_send, d_recv, bytes, ncclUint8, comm, ctx->stream));
  }
  void barrier(sb_ctx* ctx) override {
    bounce.ensure(4096);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    NCCL_CHECK(api->AllReduce(bounce.p, bounce.p, 1, ncclInt32, ncclSum, comm, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  void all_to_all_slabs(sb_ctx* ctx, const u64* d_send, u64* d_recv, const ShardPlan& plan) override {
    NCCL_CHECK(api->GroupStart());
    for (int r = 0; r < world; r++) {
      const size_t send_elems = (size_t)plan.cols[rank] * plan.rows, recv_elems = (size_t)plan.cols[r] * plan.rows;
      if (send_elems) NCCL_CHECK(api->Send(d_send + (size_t)r * send_elems, send_elems, ncclUint64, r, comm, ctx->stream));
      if (recv_elems) NCCL_CHECK(api->Recv(d_recv + (size_t)plan.col0[r] * plan.rows, recv_elems, ncclUint64, r, comm, ctx->stream));
    }
    NCCL_CHECK(api->GroupEnd());
  }
  void share_rows(sb_ctx* ctx, u64* d_rows, const uint32_t*, uint32_t count, uint32_t n_cols) override {
    // rows this rank does not own were written as zeros: a plain integer sum moves every row from its owner to everyone
    NCCL_CHECK(api->AllReduce(d_rows, d_rows, (size_t)count * n_cols, ncclUint64, ncclSum, comm, ctx->stream));
  }
  void exchange_ptrs(sb_ctx* ctx, void* mine, size_t, std::vector<void*>& all) override {
    release_ptrs();
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, mine));
    bounce.ensure(4096 + sizeof(h) * (size_t)(world + 1));
    char* d_mine = (char*)bounce.p + 4096;
    char* d_all = d_mine + sizeof(h);
    CUDA_CHECK(cudaMemcpyAsync(d_mine, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    all_gather(ctx, d_mine, d_all, sizeof(h));
    std::vector<cudaIpcMemHandle_t> hs(world);
    CUDA_CHECK(cudaMemcpyAsync(hs.data(), d_all, sizeof(h) * world, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    all.assign(world, nullptr);
    int ok = 1;
    for (int r = 0; r < world; r++) {
      if (r == rank) { all[r] = mine; continue; }
      void* q = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&q, hs[r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      opened.push_back(q);
      all[r] = q;
    }
    // every rank must take the same path: agree on the minimum
    int* d_flag = (int*)bounce.p;
    CUDA_CHECK(cudaMemcpyAsync(d_flag + 16, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(api->AllReduce(d_flag + 16, d_flag + 32, 1, ncclInt32, ncclSum, comm, ctx->stream));
    int sum = 0;
    CUDA_CHECK(cudaMemcpyAsync(&sum, d_flag + 32, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    peer_ok = sum == world;
    if (!peer_ok) { release_ptrs(); all.clear(); }
  }
  void release_ptrs() override {
    for (void* q : opened) cudaIpcCloseMemHandle(q);
    opened.clear();
  }
};

// state shared by the ranks of one process
struct LocalShared {
  int world;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  uint64_t generation = 0;
  bool failed = false;          // a rank gave up inside a collective: everyone else must not wait for it
  std::vector<const void*> slot;
  std::vector<int> device;
  explicit LocalShared(int w) : world(w), slot(w, nullptr), device(w, 0) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    if (failed) SB_THROW(SB_ENCCL, "a peer rank of this group failed");
    const uint64_t gen = generation;
    if (++waiting == world) { waiting = 0; generation++; cv.notify_all(); return; }
    const bool ok = cv.wait_for(lk, std::chrono::seconds(120), [&] { return generation != gen || failed; });
    if (!ok || failed) { failed = true; cv.notify_all(); SB_THROW(SB_ENCCL, "a peer rank of this group failed or did not arrive within 120 s"); }
  }
  void fail() {
    std::lock_guard<std::mutex> lk(mu);
    failed = true;
    cv.notify_all();
  }
};

struct LocalComm : Comm {
  std::shared_ptr<LocalShared> sh;
  void publish(sb_ctx* ctx, const void* p) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));     // what I publish is complete
    sh->slot[rank] = p;
    sh->device[rank] = ctx->device;
    sh->wait();
  }
  void copy_from(sb_ctx* ctx, void* dst, int r, size_t src_off, size_t bytes) {
    if (!bytes) return;
    const char* src = (const char*)sh->slot[r] + src_off;
    if (sh->device[r] == ctx->device) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    else CUDA_CHECK(cudaMemcpyPeerAsync(dst, ctx->device, src, sh->device[r], bytes, ctx->stream));
  }
  void done(sb_ctx* ctx) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));     // I no longer read the peers' buffers
    sh->wait();
  }
  void all_gather(sb_ctx* ctx, const void* d_send, void* d_recv, size_t bytes) override {
    publish(ctx, d_send);
    for (int r = 0; r < world; r++) copy_from(ctx, (char*)d_recv + (size_t)r * bytes, r, 0, bytes);
    done(ctx);
  }
  void barrier(sb_ctx* ctx) override {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    sh->wait();
  }
  void all_to_all_slabs(sb_ctx* ctx, const u64* d_send, u64* d_recv, const ShardPlan& plan) override {
    publish(ctx, d_send);
    for (int r = 0; r < world; r++) {
      const size_t elems = (size_t)plan.cols[r] * plan.rows;       // rank r's slab for me: its slab number `rank`
      copy_from(ctx, d_recv + (size_t)plan.col0[r] * plan.rows, r, 8 * (size_t)rank * elems, 8 * elems);
    }
    done(ctx);
  }
  void share_rows(sb_ctx* ctx, u64* d_rows, const uint32_t* owner, uint32_t count, uint32_t n_cols) override {
    publish(ctx, d_rows);
    for (uint32_t q = 0; q < count; q++)
      if ((int)owner[q] != rank) copy_from(ctx, d_rows + (size_t)q * n_cols, (int)owner[q], 8 * (size_t)q * n_cols, 8 * (size_t)n_cols);
    done(ctx);
  }
  void abort() override { sh->fail(); }
  void exchange_ptrs(sb_ctx* ctx, void* mine, size_t, std::vector<void*>& all) override {
    publish(ctx, mine);
    all.assign(world, nullptr);
    int ok = 1;
    for (int r = 0; r < world; r++) {
      all[r] = (void*)sh->slot[r];
      if (sh->device[r] != ctx->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctx->device, sh->device[r]);
        if (can) {
          cudaError_t e = cudaDeviceEnablePeerAccess(sh->device[r], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
          cudaGetLastError();
        }
        if (!can) ok = 0;
      }
    }
    // agree: slot doubles as the vote board (any rank without access clears everyone's fused path)
    sh->wait();
    { std::lock_guard<std::mutex> lk(sh->mu); if (!ok) sh->slot[0] = nullptr; }
    sh->wait();
    peer_ok = sh->slot[0] != nullptr;
    sh->wait();
    if (!peer_ok) all.clear();
  }
};

// ---------------------------------------------------------------------------------------------------------
// the group object
// ---------------------------------------------------------------------------------------------------------
struct sb_group {
  sb_ctx* ctx = nullptr;
  std::unique_ptr<Comm> comm;
  // persistent device buffers of this rank
  DevBuf rows;            // [n_cols][N / world]: this rank's row block of the LDE (peer ranks store into it)
  DevBuf slabs;           // all-to-all path only: [world][cols][rows]
  DevBuf local_trace;     // this rank's column slice when it arrives from the host
  DevBuf coeffs;          // [cols][n] coefficients of this rank's columns (openings, FRI combine)
  DevBuf work;            // digests, halo rows, gathered partials, ...
  std::vector<void*> peer_rows;
  size_t peer_rows_bytes = 0;
  std::map<std::string, float> phase_ms;
  std::string err;
};

namespace {

__global__ void first_rows_kernel(const u64* __restrict__ rows, uint32_t R, uint32_t C, u64* __restrict__ out) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = rows[(size_t)c * R];
}
// out[i] = sum over r of parts[r][i] (mod p), i < count  (NCCL has no modular reduction)
__global__ void addmod_parts_kernel(const u64* __restrict__ parts, uint32_t world, size_t count, u64* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 acc = parts[i];
  for (uint32_t r = 1; r < world; r++) acc = gl_add(acc, parts[(size_t)r * count + i]);
  out[i] = acc;
}
// dst[q][c] = rows[c][pos[q] - rank * R] if this rank owns position pos[q], else 0
__global__ void owned_rows_kernel(u64* __restrict__ dst, const u64* __restrict__ rows, uint32_t R, uint32_t C,
                                  const uint32_t* __restrict__ pos, uint32_t rank) {
  const uint32_t q = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const uint32_t p = pos[q];
  dst[(size_t)q * C + c] = (p / R == rank) ? rows[(size_t)c * R + (p - rank * R)] : 0;
}

struct Call {            // one sb_group_prove in flight
  sb_group* g;
  const sb_params* p;
  ShardPlan plan;
  const void* local_trace;
  bool on_device, fused;
  const uint64_t* pis;
};

struct Stopwatch {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  float ms() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

int guard(sb_group* g, const char* phase, const std::function<void()>& fn) {
  try {
    Stopwatch sw;
    fn();
    g->phase_ms[phase] = sw.ms();
    return SB_OK;
  } catch (const SbError& e) {
    g->err = e.msg;
    g->comm->abort();
    sb_fail(g->ctx, e);
    return e.code ? e.code : SB_EINVAL;
  } catch (const std::exception& e) {
    g->err = e.what();
    g->comm->abort();
    sb_fail(g->ctx, SbError{SB_EINVAL, e.what()});
    return SB_EINVAL;
  }
}

// ---- hook 1: the sharded trace commitment ----
int h_commit(void* user, uint64_t* cap_out) {
  Call* c = (Call*)user;
  sb_group* g = c->g;
  return guard(g, "commit", [&] {
    sb_ctx* ctx = g->ctx;
    Comm& cm = *g->comm;
    const sb_params* p = c->p;
    const ShardPlan& pl = c->plan;
    const uint32_t n = 1u << p->log_n, N = n << p->rate_bits, R = pl.rows, C = p->n_cols, cg = pl.cols[cm.rank];
    const unsigned log_blocks = ilog2(pl.world);
    // this rank's column slice on the device
    const u64* d_trace = (const u64*)c->local_trace;
    if (!c->on_device && cg) {
      g->local_trace.ensure(8ull * cg * n);
      CUDA_CHECK(cudaMemcpyAsync(g->local_trace.p, c->local_trace, 8ull * cg * n, cudaMemcpyHostToDevice, ctx->stream));
      d_trace = g->local_trace.as<u64>();
    }
    g->coeffs.ensure(8ull * std::max<uint32_t>(cg, 1) * n);
    const size_t rows_bytes = 8ull * C * R;
    bool fused = c->fused && pl.world > 1;
    if (fused && cm.peer_ok && (g->rows.cap < rows_bytes || g->peer_rows.empty() || g->peer_rows_bytes != g->rows.cap)) {
      // (re)allocate the row buffer and map every rank's: collective, and only when the shape grows
      cm.barrier(ctx);                          // nobody still stores into the old buffers
      cm.release_ptrs();
      g->rows.ensure(rows_bytes);
      cm.exchange_ptrs(ctx, g->rows.p, g->rows.cap, g->peer_rows);
      g->peer_rows_bytes = g->rows.cap;
    }
    if (fused && !cm.peer_ok) fused = false;
    g->rows.ensure(rows_bytes);
    stage_begin(ctx, "lde");
    if (fused) {
      ctx->peer_tab.ensure(8ull * 64);
      std::vector<u64*> tab(pl.world);
      for (uint32_t r = 0; r < pl.world; r++) tab[r] = (u64*)g->peer_rows[r];
      cm.barrier(ctx);                          // every rank is done reading its rows of the previous proof
      CUDA_CHECK(cudaMemcpyAsync(ctx->peer_tab.p, tab.data(), 8ull * pl.world, cudaMemcpyHostToDevice, ctx->stream));
      if (cg) sb_lde_trace(ctx, d_trace, g->coeffs.as<u64>(), nullptr, cg, p->log_n, p->rate_bits, log_blocks, (u64* const*)ctx->peer_tab.p, pl.col0[cm.rank]);
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // tab is a stack vector; and my stores are out
      cm.barrier(ctx);                          // every rank has written its columns into my rows
    } else if (pl.world > 1) {
      g->slabs.ensure(8ull * std::max<uint32_t>(cg, 1) * N);
      if (cg) sb_lde_trace(ctx, d_trace, g->coeffs.as<u64>(), g->slabs.as<u64>(), cg, p->log_n, p->rate_bits, log_blocks);
      cm.all_to_all_slabs(ctx, g->slabs.as<u64>(), g->rows.as<u64>(), pl);
    } else {
      sb_lde_trace(ctx, d_trace, g->coeffs.as<u64>(), g->rows.as<u64>(), cg, p->log_n, p->rate_bits, 0);
    }
    stage_end(ctx, "lde");
    g->phase_ms["fused"] = fused ? 1.f : 0.f;
    // row-local leaf hashing, digests of all ranks in position order, tree to the cap on every rank
    g->work.ensure(32ull * R + 32ull * N + 4096);
    u64* d_dig_local = g->work.as<u64>();
    u64* d_dig_all = d_dig_local + 4ull * R + 64;
    stage_begin(ctx, "leaf_hash");
    sb_hash_leaves_device(ctx, g->rows.as<u64>(), C, R, 0, d_dig_local);
    stage_end(ctx, "leaf_hash");
    if (pl.world > 1) cm.all_gather(ctx, d_dig_local, d_dig_all, 32ull * R);
    else d_dig_all = d_dig_local;
    ctx->tree.ensure(32ull * 2 * N);
    stage_begin(ctx, "merkle");
    sb_digests_to_leaf_order(ctx, d_dig_all, ctx->tree.as<u64>(), N, p->log_n);
    sb_merkle_levels(ctx, ctx->tree.as<u64>(), N, p->cap_height);
    stage_end(ctx, "merkle");
    CUDA_CHECK(cudaMemcpyAsync(cap_out, tree_cap_ptr(ctx->tree.as<u64>(), N, p->cap_height), 32ull << p->cap_height,
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->cur = *p;
  });
}

// ---- hook 2: row-sharded quotient values, gathered on every rank ----
int h_quotient(void* user, const uint64_t* alphas, uint64_t* d_q_out) {
  Call* c = (Call*)user;
  sb_group* g = c->g;
  return guard(g, "quotient", [&] {
    sb_ctx* ctx = g->ctx;
    Comm& cm = *g->comm;
    const sb_params* p = c->p;
    const ShardPlan& pl = c->plan;
    const uint32_t n = 1u << p->log_n, N = n << p->rate_bits, R = pl.rows, C = p->n_cols;
    g->work.ensure(8ull * C * (pl.world + 1) + 16ull * R + 4096);
    u64* d_first = g->work.as<u64>();
    u64* d_first_all = d_first + C + 32;
    u64* d_q_local = d_first_all + (size_t)C * pl.world + 32;
    const u64* d_halo = nullptr;
    // the `next` row of a block's last position lives on another rank when a block is shorter than one coset
    const uint32_t blocks_per_coset = n / R;
    if (blocks_per_coset > 1) {
      LAUNCH(ctx, first_rows_kernel, (C + 255) / 256, 256, 0, g->rows.as<u64>(), R, C, d_first);
      cm.all_gather(ctx, d_first, d_first_all, 8ull * C);
      uint32_t nxt = cm.rank + 1;
      if (nxt % blocks_per_coset == 0) nxt -= blocks_per_coset;       // wrap to the coset's first block
      d_halo = d_first_all + (size_t)nxt * C;
    }
    stage_begin(ctx, "quotient");
    sb_quotient_rows(ctx, p, g->rows.as<u64>(), R, R, cm.rank * R, d_halo, ctx->pis.as<u64>(), alphas, d_q_local);
    stage_end(ctx, "quotient");
    if (pl.world > 1) {
      // rank r's block of q_j is positions r R .. : gathering the j-th halves separately lands them in [2][N] order
      cm.all_gather(ctx, d_q_local, d_q_out, 8ull * R);
      cm.all_gather(ctx, d_q_local + R, d_q_out + N, 8ull * R);
    } else {
      CUDA_CHECK(cudaMemcpyAsync(d_q_out, d_q_local, 16ull * N, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- hook 3: openings from the column-sharded coefficient slices ----
int h_openings(void* user, const uint64_t* zeta, const uint64_t* zeta_next, uint64_t* local_out, uint64_t* next_out) {
  Call* c = (Call*)user;
  sb_group* g = c->g;
  return guard(g, "openings", [&] {
    sb_ctx* ctx = g->ctx;
    Comm& cm = *g->comm;
    const sb_params* p = c->p;
    const ShardPlan& pl = c->plan;
    const uint32_t n = 1u << p->log_n, cg = pl.cols[cm.rank];
    uint32_t width = 0;
    for (uint32_t x : pl.cols) width = std::max(width, x);
    // per rank: [2][width] extension values (ragged slices padded to the widest), gathered, then cut on the host
    g->work.ensure(32ull * n + 32ull * width * (pl.world + 1) + 4096);
    e2_t* d_tab_a = (e2_t*)g->work.p;
    e2_t* d_tab_b = d_tab_a + n;
    e2_t* d_mine = d_tab_b + n;
    e2_t* d_all = d_mine + 2ull * width;
    CUDA_CHECK(cudaMemsetAsync(d_mine, 0, 32ull * width, ctx->stream));
    const e2_t za = e2_make(zeta[0], zeta[1]), zb = e2_make(zeta_next[0], zeta_next[1]);
    if (cg) sb_openings_device(ctx, g->coeffs.as<u64>(), p->log_n, cg, za, &zb, d_tab_a, d_tab_b, d_mine, d_mine + width);
    std::vector<e2_t> host(2ull * width * pl.world);
    if (pl.world > 1) cm.all_gather(ctx, d_mine, d_all, 32ull * width);
    else d_all = d_mine;
    CUDA_CHECK(cudaMemcpyAsync(host.data(), d_all, 32ull * width * pl.world, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t r = 0; r < pl.world; r++) {
      const e2_t* part = host.data() + 2ull * width * r;
      memcpy(local_out + 2ull * pl.col0[r], part, 16ull * pl.cols[r]);
      memcpy(next_out + 2ull * pl.col0[r], part + width, 16ull * pl.cols[r]);
    }
  });
}

// ---- hook 4: FRI batch combine: per-rank partial sums with the rank's alpha-power offset, added mod p ----
int h_combine(void* user, const uint64_t* alpha, uint64_t* d_out) {
  Call* c = (Call*)user;
  sb_group* g = c->g;
  return guard(g, "combine", [&] {
    sb_ctx* ctx = g->ctx;
    Comm& cm = *g->comm;
    const sb_params* p = c->p;
    const ShardPlan& pl = c->plan;
    const uint32_t n = 1u << p->log_n, cg = pl.cols[cm.rank];
    const size_t partial_elems = (size_t)(ctx->sm_count * 16 + 8) * 128 + 2 * (size_t)n;
    g->work.ensure(16ull * (cg + 8) + 16ull * partial_elems + 16ull * n * (pl.world + 1) + 4096);
    e2_t* d_apow = (e2_t*)g->work.p;
    e2_t* d_partial = d_apow + cg + 8;
    e2_t* d_mine = d_partial + partial_elems;
    e2_t* d_all = d_mine + n;
    if (cg) sb_combine_device(ctx, g->coeffs.as<u64>(), p->log_n, cg, e2_make(alpha[0], alpha[1]), pl.col0[cm.rank], d_apow, d_partial, partial_elems, d_mine);
    else CUDA_CHECK(cudaMemsetAsync(d_mine, 0, 16ull * n, ctx->stream));
    if (pl.world > 1) {
      cm.all_gather(ctx, d_mine, d_all, 16ull * n);
      LAUNCH(ctx, addmod_parts_kernel, (2 * n + 255) / 256, 256, 0, (const u64*)d_all, pl.world, (size_t)2 * n, (u64*)d_out);
    } else {
      CUDA_CHECK(cudaMemcpyAsync(d_out, d_mine, 16ull * n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- hook 5: the full trace rows at the query positions, from the ranks that own them ----
int h_query_rows(void* user, const uint32_t* positions, uint32_t count, uint64_t* d_rows_out) {
  Call* c = (Call*)user;
  sb_group* g = c->g;
  return guard(g, "query_rows", [&] {
    sb_ctx* ctx = g->ctx;
    Comm& cm = *g->comm;
    const ShardPlan& pl = c->plan;
    const uint32_t R = pl.rows, C = c->p->n_cols;
    g->work.ensure(4ull * count + 4096);
    uint32_t* d_pos = (uint32_t*)g->work.p;
    std::vector<uint32_t> owner(count);
    for (uint32_t q = 0; q < count; q++) owner[q] = positions[q] / R;
    CUDA_CHECK(cudaMemcpyAsync(d_pos, positions, 4ull * count, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((C + 255) / 256, count);
    LAUNCH(ctx, owned_rows_kernel, grid, 256, 0, (u64*)d_rows_out, g->rows.as<u64>(), R, C, d_pos, (uint32_t)cm.rank);
    if (pl.world > 1) cm.share_rows(ctx, (u64*)d_rows_out, owner.data(), count, C);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  });
}

sb_group* new_group(sb_ctx* ctx, std::unique_ptr<Comm> comm) {
  sb_group* g = new sb_group();
  g->ctx = ctx;
  g->comm = std::move(comm);
  return g;
}

}  // namespace

extern "C" {

int sb_group_unique_id(uint8_t id[128]) {
  if (!id) return SB_EINVAL;
  try {
    ncclUniqueId u;
    NCCL_CHECK(nccl_api()->GetUniqueId(&u));
    memcpy(id, &u, 128);
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(nullptr, e); }
}

int sb_group_init_rank(sb_ctx* ctx, int rank, int world, const uint8_t id[128], sb_group** out) {
  if (!ctx || !out || !id || world < 1 || rank < 0 || rank >= world) return SB_EINVAL;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    std::unique_ptr<NcclComm> cm(new NcclComm());
    cm->api = nccl_api();
    cm->rank = rank; cm->world = world;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    NCCL_CHECK(cm->api->CommInitRank(&cm->comm, world, u, rank));
    *out = new_group(ctx, std::move(cm));
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

int sb_group_init_local(sb_ctx* const* ctxs, int world, sb_group** out) {
  if (!ctxs || !out || world < 1) return SB_EINVAL;
  try {
    auto sh = std::make_shared<LocalShared>(world);
    for (int r = 0; r < world; r++) {
      if (!ctxs[r]) SB_THROW(SB_EINVAL, "ctxs[%d] is NULL", r);
      std::unique_ptr<LocalComm> cm(new LocalComm());
      cm->rank = r; cm->world = world; cm->sh = sh;
      out[r] = new_group(ctxs[r], std::move(cm));
    }
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(nullptr, e); }
}

void sb_group_destroy(sb_group* g) {
  if (!g) return;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  g->comm.reset();
  DevBuf* bufs[] = {&g->rows, &g->slabs, &g->local_trace, &g->coeffs, &g->work};
  for (DevBuf* b : bufs) b->release();
  delete g;
}

int sb_shard_columns(const sb_params* p, uint32_t world, uint32_t rank, uint32_t* first_col, uint32_t* n_cols_local,
                     uint32_t* rows_per_rank) {
  if (!p || rank >= world) return SB_EINVAL;
  try {
    check_params(p);
    ShardPlan pl = make_plan(p, world);
    if (first_col) *first_col = pl.col0[rank];
    if (n_cols_local) *n_cols_local = pl.cols[rank];
    if (rows_per_rank) *rows_per_rank = pl.rows;
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(nullptr, e); }
}

int sb_group_rank(const sb_group* g) { return g ? g->comm->rank : -1; }
int sb_group_size(const sb_group* g) { return g ? g->comm->world : -1; }

int sb_group_column_slice(const sb_group* g, const sb_params* p, uint32_t* first_col, uint32_t* n_cols_local) {
  if (!g || !p || !first_col || !n_cols_local) return SB_EINVAL;
  try {
    check_params(p);
    ShardPlan pl = make_plan(p, (uint32_t)g->comm->world);
    *first_col = pl.col0[g->comm->rank];
    *n_cols_local = pl.cols[g->comm->rank];
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(g->ctx, e); }
}

int sb_group_prove(sb_group* g, const sb_params* p, const void* local_trace, int on_device, const uint64_t* public_inputs,
                   uint32_t flags, sb_proof** out) {
  if (!g || !p || !out) return SB_EINVAL;
  try {
    check_params(p);
    CUDA_CHECK(cudaSetDevice(g->ctx->device));
    Call call;
    call.g = g; call.p = p; call.plan = make_plan(p, (uint32_t)g->comm->world);
    call.local_trace = local_trace; call.on_device = on_device != 0; call.fused = !(flags & SB_GROUP_NO_FUSED);
    call.pis = public_inputs;
    if (call.plan.cols[g->comm->rank] && !local_trace) SB_THROW(SB_EINVAL, "local_trace is NULL");
    sb_shard_hooks hooks = {&call, h_commit, h_quotient, h_openings, h_combine, h_query_rows};
    g->phase_ms.clear();
    return sb_prove_sharded(g->ctx, p, &hooks, public_inputs, out);
  } catch (const SbError& e) { return sb_fail(g->ctx, e); }
}

float sb_group_phase_ms(const sb_group* g, const char* phase) {
  if (!g || !phase) return -1.f;
  auto it = g->phase_ms.find(phase);
  return it == g->phase_ms.end() ? -1.f : it->second;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// One process, several GPUs: sb_init(devices, n > 1) builds one rank context per device, a LocalComm group over them and
// proves with one host thread per device.  The caller sees one sb_ctx.
// ---------------------------------------------------------------------------------------------------------
struct sb_multi {
  std::vector<sb_ctx*> ranks;
  std::vector<sb_group*> groups;
};

void multi_destroy(sb_ctx* ctx) {
  sb_multi* m = ctx->multi;
  if (!m) return;
  for (sb_group* g : m->groups) sb_group_destroy(g);
  for (sb_ctx* c : m->ranks) sb_destroy(c);
  delete m;
  ctx->multi = nullptr;
}

void multi_init(sb_ctx* ctx, const int* devices, int n) {
  sb_multi* m = new sb_multi();
  ctx->multi = m;
  for (int i = 0; i < n; i++) {
    sb_ctx* c = nullptr;
    int rc = sb_init(devices + i, 1, &c);
    if (rc) { multi_destroy(ctx); SB_THROW(rc, "sb_init(device %d): %s", devices[i], sb_last_error(nullptr)); }
    m->ranks.push_back(c);
  }
  m->groups.resize(n);
  int rc = sb_group_init_local(m->ranks.data(), n, m->groups.data());
  if (rc) { m->groups.clear(); multi_destroy(ctx); SB_THROW(rc, "sb_group_init_local: %s", sb_last_error(nullptr)); }
}

// sb_prove on a multi-device ctx: the trace (host, column-major or column pointers) is cut into the ranks' column slices
int multi_prove(sb_ctx* ctx, const sb_params* p, const void* trace, int layout, const uint64_t* public_inputs, sb_proof** out) {
  sb_multi* m = ctx->multi;
  const int world = (int)m->ranks.size();
  if (layout != SB_TRACE_COLMAJOR_U64 && layout != SB_TRACE_COLS_U64_PTRS)
    SB_THROW(SB_EINVAL, "a multi-GPU context takes the trace as SB_TRACE_COLMAJOR_U64 or SB_TRACE_COLS_U64_PTRS (layout %d)", layout);
  if (!trace) SB_THROW(SB_EINVAL, "trace is NULL");
  check_params(p);
  const ShardPlan pl = make_plan(p, (uint32_t)world);
  const size_t n = size_t(1) << p->log_n;
  std::vector<int> rcs(world, 0);
  std::vector<sb_proof*> proofs(world, nullptr);
  std::vector<std::vector<u64>> gathered(world);
  std::vector<std::thread> th;
  for (int r = 0; r < world; r++) {
    th.emplace_back([&, r] {
      const void* slice = nullptr;
      if (layout == SB_TRACE_COLMAJOR_U64) slice = (const u64*)trace + (size_t)pl.col0[r] * n;
      else {
        const u64* const* cols = (const u64* const*)trace;
        gathered[r].resize((size_t)pl.cols[r] * n);
        for (uint32_t c = 0; c < pl.cols[r]; c++) memcpy(gathered[r].data() + (size_t)c * n, cols[pl.col0[r] + c], 8 * n);
        slice = gathered[r].data();
      }
      rcs[r] = sb_group_prove(m->groups[r], p, slice, 0, public_inputs, 0, &proofs[r]);
    });
  }
  for (auto& t : th) t.join();
  int rc = SB_OK;
  for (int r = 0; r < world; r++) if (rcs[r] && !rc) { rc = rcs[r]; ctx->err = m->ranks[r]->err; }
  for (int r = 1; r < world; r++) sb_proof_free(proofs[r]);
  if (rc) { sb_proof_free(proofs[0]); return rc; }
  ctx->launches = 0;
  for (sb_ctx* c : m->ranks) ctx->launches += c->launches;
  *out = proofs[0];
  return SB_OK;
}
