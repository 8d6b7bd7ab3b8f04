// sb_prove_batch: the independent proofs of one job list (the reference proves the seven starky proofs of one BLS
// signature verification one after the other on one thread, /root/reference/src/aggregate_proof.rs:279-370; they do not
// depend on each other) scheduled over a set of contexts, one internal host thread per context.
//
// What the schedule knows (measured on B200):
//   * few-leaf proofs (MillerLoop 2048, PairingPrecomp 4096, FP12Mul 32 leaves) are latency-bound: their leaf sponge is one
//     32-leaf group per SM walking a chain of ceil(C/8) permutations, and a large share of the proof is the strictly
//     sequential host transcript.  Several of them in flight on one GPU fill each other's gaps (2 in flight: 212 -> 118 ms
//     per MillerLoop proof, 4: 78 ms).
//   * many-leaf proofs (FinalExp, ECCAgg: 32768 leaves) fill the GPU on their own.  With a device-resident trace their leaf
//     sponge is ONE launch whose blocks hold every SM for the whole chain, and a latency-bound chain next to it starves
//     (MillerLoop 225 -> 937 ms).  Inside a batch the commitment therefore goes column group by column group for every
//     trace layout (capi.cu: one K1 + one K2 launch per >= 2048 columns, ~12 ms; for the host column layouts the groups
//     follow the PCIe slabs anyway): the small jobs' blocks get in at every launch boundary and fill the issue slots the
//     leaf sponge leaves (< 50 % issue) -- the full BLS set on one GPU 918 -> 806 ms in round 2, 700 ms with this round's
//     leaf sponge.  SB_SCHED_MIX=0 restores "a throughput-bound job runs alone on its device".
// Rule per device: a throughput-bound job always runs on the device's first context, so that its tens of GB of buffers
// exist once, and one at a time; latency-bound jobs share a device up to the number of contexts it has, also next to a
// throughput-bound job.  Jobs are taken in decreasing estimated cost.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "prover.cuh"

namespace {

struct DeviceState { int running_few = 0; bool running_big = false, big_is_streamed = false; };

double job_cost(const sb_params& p) {
  // permutations of the trace commitment dominate every proof; a long per-leaf chain costs latency on top
  const double N = (double)(1u << (p.log_n + p.rate_bits)), chain = (p.n_cols + 7) / 8;
  return chain * N + 2.0e4 * chain;
}

}  // namespace

extern "C" int sb_prove_batch(sb_ctx* const* ctxs, int n_ctx, sb_job* jobs, int n_jobs) {
  if (!ctxs || n_ctx < 1 || (!jobs && n_jobs) || n_jobs < 0) return SB_EINVAL;
  for (int c = 0; c < n_ctx; c++) if (!ctxs[c]) return SB_EINVAL;
  std::vector<int> order(n_jobs);
  for (int i = 0; i < n_jobs; i++) { order[i] = i; jobs[i].proof = nullptr; jobs[i].rc = SB_OK; jobs[i].ms = 0.f; }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return job_cost(jobs[a].params) > job_cost(jobs[b].params); });
  std::vector<char> taken(n_jobs, 0);
  std::map<int, DeviceState> dev;
  std::mutex mu;
  std::condition_variable cv;
  int remaining = n_jobs;
  auto dev_of = [&](int c) { return ctxs[c]->multi ? -1 - c : ctxs[c]->device; };   // a multi-device ctx is its own "device"
  std::map<int, int> first_ctx;
  for (int c = n_ctx - 1; c >= 0; c--) first_ctx[dev_of(c)] = c;

  auto is_big = [&](const sb_job& j, const sb_ctx* ctx) {
    return (uint64_t(1) << (j.params.log_n + j.params.rate_bits)) > 64ull * (uint64_t)ctx->sm_count;
  };
  const bool mix = !(getenv("SB_SCHED_MIX") && atoi(getenv("SB_SCHED_MIX")) == 0);   // SB_SCHED_MIX=0: never share a device with a big job
  // SB_SCHED_TRACE=1: one stderr line per job (context, start and end in ms since the call)
  const bool trace = getenv("SB_SCHED_TRACE") && atoi(getenv("SB_SCHED_TRACE")) != 0;
  const auto t_call = std::chrono::steady_clock::now();
  auto since = [&] { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
  int first_turn = 0;      // the first picks go in context order, so that the same job list lands on the same contexts
                           // from call to call (resident buffers and bound constraint programs are per context)
  auto worker = [&](int c) {
    sb_ctx* ctx = ctxs[c];
    const int d = dev_of(c);
    const bool takes_big = first_ctx[d] == c;
    bool first = true;
    for (;;) {
      int pick = -1;
      bool big = false;
      {
        std::unique_lock<std::mutex> lk(mu);
        bool my_turn = first;
        if (first) {
          cv.wait(lk, [&] { return first_turn == c; });
          first = false;
        }
        auto release_turn = [&] { if (my_turn) { my_turn = false; first_turn++; cv.notify_all(); } };
        for (;;) {
          if (remaining == 0) { release_turn(); return; }
          DeviceState& st = dev[d];
          bool any_left = false;
          for (int k : order) {
            if (taken[k]) continue;
            any_left = true;
            const bool b = is_big(jobs[k], ctx);
            const bool streamed = mix;      // every layout commits its trace group by group inside a batch (ctx->yield_slabs)
            if (st.running_big && (b || !st.big_is_streamed)) break;   // the device belongs to a throughput-bound job
            if (b && (!takes_big || (st.running_few > 0 && !streamed))) continue;  // needs its device's first context (and an idle device unless it streams); a smaller job may still fit
            pick = k; big = b;
            break;
          }
          if (pick >= 0 || !any_left) break;
          release_turn();
          cv.wait(lk);
        }
        release_turn();
        if (pick < 0) return;                                       // nothing left to take (others are finishing)
        taken[pick] = 1;
        DeviceState& st = dev[d];
        if (big) {
          st.running_big = true;
          st.big_is_streamed = mix;
        } else st.running_few++;
      }
      sb_job& j = jobs[pick];
      const auto t0 = std::chrono::steady_clock::now();
      const float ts = since();
      ctx->yield_slabs = mix;
      j.rc = sb_prove(ctx, &j.params, j.trace, j.layout, j.public_inputs, &j.proof);
      ctx->yield_slabs = false;
      j.ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (trace) fprintf(stderr, "sb_prove_batch: job %d (stark %u) on context %d: %.1f -> %.1f ms\n", pick, j.params.stark_id, c, ts, since());
      {
        std::lock_guard<std::mutex> lk(mu);
        DeviceState& st = dev[d];
        if (big) st.running_big = false; else st.running_few--;
        remaining--;
      }
      cv.notify_all();
    }
  };
  std::vector<std::thread> th;
  for (int c = 0; c < n_ctx; c++) th.emplace_back(worker, c);
  for (auto& t : th) t.join();
  int rc = SB_OK;
  for (int i = 0; i < n_jobs; i++) if (jobs[i].rc && !rc) rc = jobs[i].rc;
  return rc;
}
