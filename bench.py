#!/usr/bin/env python3
"""bench.py -- headline benchmark: starky prove ms/STARK (MillerLoop, FinalExp) on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--stark miller_loop] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full proof (sb_prove: trace LDE -> Poseidon Merkle -> quotient -> FRI -> proof in host memory) of one
synthetic trace of the workload stark (default FinalExponentiateStark 73527 x 8192, BASELINE configs[3]: the heaviest of
the reference's proofs -- 92 s of its ~2 min on 32 vCPU -- and the shape whose leaf hashing and quotient shard by rows).
  N = 1  value : ms per proof, trace already resident in HBM when the timed region starts (DEVICE_COLMAJOR_U64)
         e2e   : ms per proof through the plugin call with the trace in pinned HOST memory (H2D of the trace and D2H of
                 the proof inside); `e2e_pageable_cols` is the same from 73 527 separately allocated pageable columns,
                 the Vec<PolynomialValues<F>> the reference hands to prove() (aggregate_proof.rs:57-59)
         also  : MillerLoopStark 97330 x 1024 (BASELINE configs[2]) with its own value / e2e / roofline / full CPU proof
  N > 1  value : ms of ONE proof of the same stark with the trace SHARDED over the N GPUs (column-sharded LDE -> NVLink
                 exchange -> row-sharded leaf hashing + quotient -> small collectives): strong scaling.
                 `replicas` (one independent proof per GPU, weak) and the sharded MillerLoop proof are extras.
  --impl reference : the CPU restatement of the reference's prover (oracle/, all host threads) on the same stark; each
                 step is a bounded sample (the same AIR and all columns on 1/32 of the rows, time x 32).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

U32_MACS_PER_PERM = 6400          # algorithmic u32 multiply-adds per Poseidon-12 permutation
U32_MACS_PER_CONSTRAINT = 10      # algorithmic u32 multiply-adds per constraint evaluation
WORKLOADS = {
    "fp12_mul": "FP12MulStark 60285 cols x 16 rows, rate_bits 1 (BASELINE configs[0])",
    "pairing_precomp": "PairingPrecompStark 29376 cols x 1024 rows, rate_bits 2 (BASELINE configs[1])",
    "miller_loop": "MillerLoopStark 97330 cols x 1024 rows, rate_bits 1 (BASELINE configs[2])",
    "final_exp": "FinalExponentiateStark 73527 cols x 8192 rows, rate_bits 2 (BASELINE configs[3])",
    "ecc_agg": "ECCAggStark 3339 cols x 8192 rows, rate_bits 2 (BASELINE configs[4] member)",
}
K_CONSTRAINTS = {"fp12_mul": 82560, "pairing_precomp": 113634, "miller_loop": 145574, "final_exp": 360800, "ecc_agg": 20013}
# rows of the CPU sample: the same AIR and all columns on 2^-shift of the rows, time x 2^shift (leaf hashing, quotient
# and openings are linear in the rows; the NTTs lose a log factor, so the scaled figure slightly favours the CPU)
CPU_SAMPLE_SHIFT = {"fp12_mul": 0, "pairing_precomp": 2, "miller_loop": 3, "final_exp": 5, "ecc_agg": 2}


def config_for(stark, world):
    """The `config` object of the JSON line: a function of (stark, world) only, so both arms print the same dict."""
    return {"workload": WORKLOADS[stark], "trace": "synthetic: uniform u32 cells and public inputs, seeded PCG64",
            "parallelism": "one proof on one GPU" if world == 1 else
                           "one proof, trace sharded over %d GPUs (columns for the LDE, rows for leaf hashing and quotient)" % world,
            "l2": "inputs larger than L2 (the trace alone is 0.8 GB for MillerLoop, 4.8 GB for FinalExp)"}


def synthetic(info, seed):
    """SURVEY 8d distribution A: every trace cell and public input uniform in [0, 2^32) (real traces are u32 limbs/bits)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    trace = rng.integers(0, 1 << 32, (info.columns, info.num_rows), dtype=np.uint64)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    return trace, pis


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", os.environ.get("SB_BENCH_SMI_MS", "250"),
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [c for c in sm if mx and c > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def probe_plan(world):
    """Smallest shard plan that shard_plan accepts at every power-of-two world size up to 64 (used by callers that only
    need *a* valid plan, e.g. to probe peer access)."""
    from starky_bls12_381_b200.sharded import shard_plan
    return shard_plan(8, 10, 1, world)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(stark, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `stark`, from the committed `ncu --set full`
    capture of this round (profiles/r2_dram_traffic.json, written by tools/perf/summarize_ncu.py); None if not captured."""
    path = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    try:
        return json.load(open(path)).get(stark, {}).get(kernel)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of starky::prover::prove, all host threads
# ---------------------------------------------------------------------------------------------------------------------
def oracle_all_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 for nproc > 1: the CPU arm must still use every host core."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    import oracle_lib as O
    O.build()
    O.lib().orc_set_num_threads(n)
    return O, int(O.lib().orc_num_threads())


def cpu_sample(O, sb, stark, seed, shift=None, check_ctx=None):
    """One proof by the CPU port of the same AIR and all columns on num_rows >> shift rows.  Returns (ms scaled to the
    full height, description)."""
    from starky_bls12_381_b200 import airfiles
    info = sb.STARKS[stark]
    shift = CPU_SAMPLE_SHIFT[stark] if shift is None else shift
    rows = info.num_rows >> shift
    flat = airfiles.air_path(stark, "air")
    rng = np.random.Generator(np.random.PCG64(seed))
    trace = rng.integers(0, 1 << 32, (info.columns, rows), dtype=np.uint64)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    p = O.make_params(stark_id=info.stark_id, log_n=rows.bit_length() - 1, n_cols=info.columns, n_pis=info.public_inputs,
                      degree=info.constraint_degree, rate_bits=info.rate_bits, flags=1)
    t0 = time.perf_counter()
    rc, words = O.prove(flat, p, trace, pis)
    dt = time.perf_counter() - t0
    assert rc == 0, O.err()
    what = ("one full proof" if shift == 0 else
            "one proof of the same AIR and all %d columns on %d of the %d rows, time x %d" % (info.columns, rows, info.num_rows, 1 << shift))
    if check_ctx is not None:          # the GPU path on the very same sample must give the very same proof
        gp = sb.standard_params(info.stark_id, rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
        same = bool(np.array_equal(check_ctx.prove(gp, trace, pis).words, words))
        return 1e3 * dt * (1 << shift), what, same
    return 1e3 * dt * (1 << shift), what


def cpu_baseline_for(ctx, sb, stark, p, trace, pis, gpu_proof):
    """cpu_baseline block of one stark at N = 1: the CPU port with all host threads.  Starks whose whole proof fits the time
    budget (everything but FinalExp: <= ~20 s) are proved in full on the SAME trace and compared word for word with the
    GPU's proof; FinalExp (minutes on CPU) is sampled on 1/32 of the rows, and the GPU proves that same sample."""
    O, cores = oracle_all_threads()
    from starky_bls12_381_b200 import airfiles
    if stark != "final_exp":
        op = O.Params.from_buffer_copy(bytes(p))
        t0 = time.perf_counter()
        rc, words = O.prove(airfiles.air_path(stark, "air"), op, trace, pis)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        return {"value": cpu_ms, "unit": "ms", "cores": cores, "kind": "port",
                "sample": "one full proof of the same trace (whole workload, no scaling)",
                "proof_bit_identical_to_gpu": bool(rc == 0 and np.array_equal(words, gpu_proof.words))}
    ms, what, same = cpu_sample(O, sb, stark, 0xB2500000 + sb.STARKS[stark].stark_id, check_ctx=ctx)
    return {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": what, "proof_bit_identical_to_gpu_on_the_sample": same}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import starky_bls12_381_b200 as sb
    O, cores = oracle_all_threads()
    times, what = [], ""
    for it in range(args.warmup + args.steps):
        ms, what = cpu_sample(O, sb, args.stark, 0xB2000000 + it)
        if it >= args.warmup:
            times.append(ms)
    ms = float(np.mean(times))
    line = {"impl": "reference", "metric": "starky_prove_ms_per_stark", "value": ms, "unit": "ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "u64 (Goldilocks)", "data": "synthetic",
            "config": config_for(args.stark, world),
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": what,
                             "note": "CPU restatement of the reference algorithm (oracle/), not the Rust binary: no cargo in the image"},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# one stark on one GPU: resident / end-to-end timings, stage breakdown, roofline blocks
# ---------------------------------------------------------------------------------------------------------------------
def roofline_for(stark, info, kern, ms_step, imad, hbm_peak, peak_src):
    C, n, N = info.columns, info.num_rows, info.num_rows << info.rate_bits
    perms = -(-C // 8) * N + (N - 16)
    t_hash = (kern["leaf_hash"] + kern["merkle"]) * 1e-3
    t_lde, t_q = kern["lde"] * 1e-3, kern["quotient"] * 1e-3
    lde_bytes = 8.0 * C * (n + n + N)           # trace read + coefficients kept + LDE written
    k2 = "leaf_sponge_mm_kernel" if N <= 64 * 148 else "leaf_sponge_mm_het_kernel"
    return {
        # dominant kernel of the step: the Poseidon leaf sponge (integer-pipe bound)
        "kernel": k2 + "+merkle_level_kernel", "bound": "imad",
        "achieved": perms * U32_MACS_PER_PERM / t_hash / 1e9, "peak": imad["mad_lo_u32_gops"], "unit": "Gop/s (u32 multiply-add)",
        "frac": perms * U32_MACS_PER_PERM / t_hash / 1e9 / imad["mad_lo_u32_gops"],
        "traffic": measured_traffic(stark, k2), "algorithmic_bytes": 8 * C * N,
        "peak_source": "sb_measure_imad_peak: dependent-free mad.lo.u32, measured in this run",
        "share_of_step": (kern["leaf_hash"] + kern["merkle"]) / ms_step,
        "stages": {
            "lde": {"bound": "hbm", "achieved": lde_bytes / t_lde / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": lde_bytes / t_lde / 1e9 / hbm_peak, "ms": kern["lde"], "peak_source": peak_src,
                    "traffic": measured_traffic(stark, "lde8_kernel"), "algorithmic_bytes": lde_bytes},
            "merkle_hbm_view": {"bound": "hbm", "achieved": 8.0 * C * N / t_hash / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": 8.0 * C * N / t_hash / 1e9 / hbm_peak, "ms": kern["leaf_hash"] + kern["merkle"]},
            "quotient": {"bound": "imad", "achieved": K_CONSTRAINTS[stark] * N * U32_MACS_PER_CONSTRAINT / t_q / 1e9,
                         "peak": imad["mad_lo_u32_gops"], "unit": "Gop/s (u32 multiply-add)",
                         "frac": K_CONSTRAINTS[stark] * N * U32_MACS_PER_CONSTRAINT / t_q / 1e9 / imad["mad_lo_u32_gops"],
                         "ms": kern["quotient"], "traffic": measured_traffic(stark, "quotient_run_kernel"), "algorithmic_bytes": 8 * C * N},
            "lde_quotient_merkle": {"ms": kern["lde"] + kern["quotient"] + kern["leaf_hash"] + kern["merkle"],
                                    "lde_merkle_gbs": 8.0 * C * N / ((kern["lde"] + kern["leaf_hash"] + kern["merkle"]) * 1e-3) / 1e9},
        },
    }


def pageable_columns(trace):
    """The reference's Vec<PolynomialValues<F>>: one separately allocated (pageable) buffer per column + the pointer array."""
    cols = [np.array(trace[c], dtype=np.uint64, copy=True) for c in range(trace.shape[0])]
    ptrs = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    return cols, ptrs


def measure_stark(ctx, sb, stark, steps, warmup, timed, seed, pageable_leg=True):
    """value / e2e / stage breakdown of `stark` on this rank's GPU.  Returns (dict, trace, pis, last proof)."""
    import torch
    info = sb.STARKS[stark]
    p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
    trace, pis = synthetic(info, seed)
    host = torch.from_numpy(trace.view(np.int64)).pin_memory()
    host_ptr = host.data_ptr()
    ctx.trace_upload(p, host_ptr)
    resident = lambda: ctx.prove(p, None, pis, sb.TraceLayout.DEVICE_COLMAJOR_U64)
    e2e_fn = lambda: ctx.prove(p, host_ptr, pis, sb.TraceLayout.COLMAJOR_U64)
    # warm-up through the same loop as the timed steps: there the previous proof is still alive while the next one is made,
    # so the library's pool of pinned proof buffers needs two of them (a first cudaMallocHost of 52 MB inside the timed
    # region cost 40-80 ms of one step)
    timed(resident, warmup)
    l0 = ctx.kernel_launches()
    dt, proofs = timed(resident, steps)
    step_ms = list(getattr(timed, "step_ms", []))
    launches = ctx.kernel_launches() - l0
    stage_of = [pr if isinstance(pr, dict) else pr.timings for pr in proofs]
    stage = {k: float(np.mean([t[k] for t in stage_of])) for k in stage_of[0]}
    kern = {k: ctx.stage_ms(k) for k in ("lde", "leaf_hash", "merkle", "quotient")}
    timed(e2e_fn, 2)
    dt_e2e, _ = timed(e2e_fn, steps)
    out = {"ms": 1e3 * dt / steps, "ms_e2e": 1e3 * dt_e2e / steps, "step_ms": step_ms, "stage_ms": stage, "kernel_ms": kern, "launches": int(launches),
           "h2d_bytes": 8 * info.columns * info.num_rows + 8 * info.public_inputs,
           "d2h_bytes": int(proofs[-1].layout.total_words) * 8,
           "leaf_hash_mperm_s": (-(-info.columns // 8) * (info.num_rows << info.rate_bits)) / (kern["leaf_hash"] * 1e-3) / 1e6}
    if pageable_leg:
        cols, ptrs = pageable_columns(trace)
        pg = lambda: ctx.prove(p, ctypes.addressof(ptrs), pis, sb.TraceLayout.COLS_U64_PTRS)
        timed(pg, 2)
        k = max(1, min(steps, 3))
        dt_pg, _ = timed(pg, k)
        out["ms_e2e_pageable_cols"] = 1e3 * dt_pg / k
        del cols, ptrs
    del host
    return out, trace, pis, proofs[-1], p


FULL_SET = ["final_exp", "miller_loop", "miller_loop", "pairing_precomp", "pairing_precomp", "ecc_agg", "fp12_mul"]
FULL_SET_COST = {"final_exp": 650.0, "miller_loop": 225.0, "pairing_precomp": 75.0, "ecc_agg": 35.0, "fp12_mul": 6.0}


def full_set_assignment(world):
    """BASELINE configs[4]: the seven starky proofs of one BLS signature verification (aggregate_proof.rs:279-370 proves
    them one after the other; they are independent once the native inputs are known).  Longest-processing-time-first
    onto `world` GPUs; returns one list of stark names per rank."""
    return full_set_assignment_of(FULL_SET, world)


def full_set_plan(world, fe_over="all"):
    """How the seven proofs are laid over `world` GPUs.  From four GPUs on, the critical path (FinalExp, ~600 ms on one GPU)
    is sharded (SURVEY 8e "proof-level parallelism on top"); below that whole proofs are assigned longest-first.
      fe_over = "all":  FinalExp over every GPU, the other six proofs longest-first over the same GPUs on their other
                        contexts, at the same time (a FinalExp shard of <= 9472 leaves per GPU is latency-bound like them);
      fe_over = "half": FinalExp over half of the GPUs, the other six proofs share the rest (the round-1 plan).
    Returns (ranks of the sharded FinalExp proof or [], per-rank lists of whole proofs)."""
    if world < 4:
        return [], full_set_assignment(world)
    rest_names = [x for x in FULL_SET if x != "final_exp"]
    if fe_over == "all":
        return list(range(world)), full_set_assignment_of(rest_names, world)
    k = world // 2
    return list(range(k)), [[] for _ in range(k)] + full_set_assignment_of(rest_names, world - k)


def full_set_assignment_of(names, world):
    loads, out = [0.0] * world, [[] for _ in range(world)]
    for name in sorted(names, key=lambda k: -FULL_SET_COST[k]):
        g = min(range(world), key=lambda r: loads[r])
        out[g].append(name)
        loads[g] += FULL_SET_COST[name]
    return out


def run_full_set(sb, contexts, names, rank, timed, sharded_job=None):
    """Proves `names` on this rank's GPU, end to end from pinned host memory, through sb_prove_batch: the library's scheduler
    (internal host threads, one per context) runs the latency-bound proofs up to len(contexts) in flight and gives the
    throughput-bound ones the GPU to themselves.  Returns (seconds max over ranks, per-proof ms)."""
    import torch
    from starky_bls12_381_b200.binding import prove_batch
    jobs, keep = [], []
    for i, name in enumerate(names):
        info = sb.STARKS[name]
        trace, pis = synthetic(info, 0xB2300000 + 16 * rank + i)
        host = torch.from_numpy(trace.view(np.int64)).pin_memory()
        del trace
        p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
        keep.append(host)
        jobs.append((p, host.data_ptr(), sb.TraceLayout.COLMAJOR_U64, pis))
    per = []

    def batch():
        for name, (res, ms) in zip(names, prove_batch(contexts, jobs)):
            if isinstance(res, Exception):
                raise res
            per.append((name, ms))

    def go():
        th, err = None, []
        if jobs and sharded_job is not None:      # this rank's whole proofs run next to its FinalExp shard, on other contexts
            def guarded():
                try:
                    batch()
                except Exception as e:            # noqa: BLE001
                    err.append(e)
            th = threading.Thread(target=guarded)
            th.start()
        if sharded_job is not None:
            t0 = time.perf_counter()
            sharded_job()
            per.append(("final_exp(sharded)", 1e3 * (time.perf_counter() - t0)))
        if th is not None:
            th.join()
            if err:
                raise err[0]
        elif jobs:
            batch()
    go()                 # warm-up: buffers of every shape allocated, constraint programs bound
    go()                 # (the scheduler may give a context another shape the second time: settle the buffer sizes)
    per.clear()
    dt, _ = timed(go, 1)
    return dt, per


def optional(leg, world):
    """Runs a secondary leg of the bench.  On one GPU a failure there (e.g. out of memory on a smaller part) must not cost the
    headline line; with several ranks an exception is re-raised (a rank that skips a collective would hang the others)."""
    try:
        return leg()
    except Exception as e:          # noqa: BLE001
        if world > 1:
            raise
        print("bench: secondary leg %s failed: %r" % (leg.__name__, e), file=sys.stderr, flush=True)
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--stark", default="final_exp", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="sharded legs: NCCL all-to-all after K1 instead of K1 storing into peer memory")
    ap.add_argument("--no-full-set", action="store_true", help="skip the 7-proof BLS set (BASELINE configs[4])")
    ap.add_argument("--fe-over", default="all", choices=["all", "half"],
                    help="7-proof set from 4 GPUs on: FinalExp sharded over all GPUs next to the other proofs, or over half of them")
    ap.add_argument("--bundled", action="store_true",
                    help="N=1: also prove the seven proofs of the reference's bundled light_client_update_period_1052/1053 inputs "
                         "(valid traces from the witness generators, ~1 min of host-side trace generation outside the timed region)")
    ap.add_argument("--no-extras", action="store_true", help="headline line only (no 'also', sharded FinalExp, in-flight, full set)")
    ap.add_argument("--also", default="miller_loop",
                    help="further starks measured after the headline workload (N=1: whole proofs; N>1: sharded proofs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)            # timing rules: at least three warm-up steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import starky_bls12_381_b200 as sb
    info = sb.STARKS[args.stark]
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = sb.Context(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        out = []
        timed.step_ms = []
        for _ in range(steps):
            ts = time.perf_counter()
            res = fn()
            timed.step_ms.append(round(1e3 * (time.perf_counter() - ts), 2))     # (every fn here returns after its proof is complete)
            if out and hasattr(out[-1], "timings"):
                out[-1] = out[-1].timings        # keep the stage timings, release the proof: its pinned buffer goes back to
            out.append(res)                       # the library's pool, as it does in a caller that consumes each proof
            del res
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    extras = not args.no_extras
    also_names = [x for x in args.also.split(",") if x and x != args.stark] if extras else []
    C, n, N = info.columns, info.num_rows, info.num_rows << info.rate_bits
    line = {}

    if world == 1:
        # ---------------- N = 1: whole proofs on one GPU ----------------
        m, trace, pis, last, p = measure_stark(ctx, sb, args.stark, args.steps, args.warmup, timed, 0xB2000000 + info.stark_id)
        clocks = sampler.stop()
        hbm_peak, peak_src = peaks()
        imad = ctx.measure_imad_peak()
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_for(ctx, sb, args.stark, p, trace, pis, last)
        del trace
        line = {
            "metric": "starky_prove_ms_per_stark", "value": m["ms"], "unit": "ms", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms"], "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (Goldilocks)", "data": "synthetic", "config": config_for(args.stark, 1),
            "e2e": {"value": m["ms_e2e"], "unit": "ms", "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"]},
            "e2e_pageable_cols": {"value": m.get("ms_e2e_pageable_cols"), "unit": "ms",
                                  "note": "trace handed over as %d separately allocated pageable columns (SB_TRACE_COLS_U64_PTRS, the "
                                          "reference's Vec<PolynomialValues<F>>): gathered into pinned staging slabs inside the call" % C},
            "gpu_launches": m["launches"], "clocks": clocks,
            "roofline": roofline_for(args.stark, info, m["kernel_ms"], m["ms"], imad, hbm_peak, peak_src),
            "cpu_baseline": cpu, "step_ms": m["step_ms"], "stage_ms": m["stage_ms"], "kernel_ms": m["kernel_ms"], "leaf_hash_mperm_s": m["leaf_hash_mperm_s"],
        }
        also = {}
        for name in also_names:
            def leg(name=name):
                ai = sb.STARKS[name]
                am, atrace, apis, alast, ap_ = measure_stark(ctx, sb, name, 3, 3, timed, 0xB2000000 + ai.stark_id, pageable_leg=False)
                out = {"workload": WORKLOADS[name], "value": am["ms"], "unit": "ms",
                       "e2e": {"value": am["ms_e2e"], "unit": "ms", "h2d_bytes_per_step": am["h2d_bytes"], "d2h_bytes_per_step": am["d2h_bytes"]},
                       "stage_ms": am["stage_ms"], "kernel_ms": am["kernel_ms"], "leaf_hash_mperm_s": am["leaf_hash_mperm_s"],
                       "roofline": roofline_for(name, ai, am["kernel_ms"], am["ms"], imad, hbm_peak, peak_src)}
                if not args.no_cpu_baseline:
                    out["cpu_baseline"] = cpu_baseline_for(ctx, sb, name, ap_, atrace, apis, alast)
                del atrace
                return out
            leg.__name__ = "also_" + name
            also[name] = optional(leg, world)
        line["also"] = also
    else:
        # ---------------- N > 1: ONE proof sharded over all ranks (strong scaling) ----------------
        from starky_bls12_381_b200 import multi
        group = multi.Group.from_torch(ctx, rank, world, local_rank)       # NCCL communicator + peer row buffers inside the library
        p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
        # one trace whose column slices are drawn where they live (seed + rank): no 4.8 GB trace is replicated per rank
        c0, cg = group.column_slice(p)
        lrng = np.random.Generator(np.random.PCG64(0xB2000000 + info.stark_id + 1000 * rank))
        local_host = torch.from_numpy(lrng.integers(0, 1 << 32, (cg, info.num_rows), dtype=np.uint64).view(np.int64)).pin_memory()
        local_dev = local_host.cuda()
        pis = np.random.Generator(np.random.PCG64(0xB2000000 + info.stark_id)).integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
        fused = not args.no_fused
        res_fn = lambda: group.prove(p, local_dev.data_ptr(), pis, on_device=True, fused=fused)
        e2e_fn = lambda: group.prove(p, local_host.data_ptr(), pis, on_device=False, fused=fused)
        timed(res_fn, args.warmup)          # warm-up in the pattern of the timed loop (two proof buffers alive, see measure_stark)
        l0 = ctx.kernel_launches()
        dt, proofs = timed(res_fn, args.steps)
        launches = ctx.kernel_launches() - l0
        kern = {k: ctx.stage_ms(k) for k in ("lde", "leaf_hash", "merkle", "quotient")}
        timed(e2e_fn, 2)
        dt_e2e, _ = timed(e2e_fn, args.steps)
        same = multi.same_on_every_rank(proofs[-1].words)
        clocks = sampler.stop() if rank == 0 else None
        ms, ms_e2e = 1e3 * dt / args.steps, 1e3 * dt_e2e / args.steps
        hbm_peak, peak_src = peaks()
        imad = ctx.measure_imad_peak()
        line = {
            "metric": "starky_prove_ms_per_stark", "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64 (Goldilocks)", "data": "synthetic", "config": config_for(args.stark, world),
            "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": 8 * cg * n + 8 * info.public_inputs,
                    "d2h_bytes_per_step": int(proofs[-1].layout.total_words) * 8,
                    "note": "every rank copies its column slice from pinned host memory inside the call; bytes are rank 0's"},
            "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": None,
            "roofline": {"kernel": "leaf_sponge (this rank's %d of %d leaves)" % (N // world, N), "bound": "imad",
                         "achieved": (-(-C // 8) * (N // world)) * U32_MACS_PER_PERM / (kern["leaf_hash"] * 1e-3) / 1e9,
                         "peak": imad["mad_lo_u32_gops"], "unit": "Gop/s (u32 multiply-add)",
                         "frac": (-(-C // 8) * (N // world)) * U32_MACS_PER_PERM / (kern["leaf_hash"] * 1e-3) / 1e9 / imad["mad_lo_u32_gops"],
                         "traffic": None, "share_of_step": kern["leaf_hash"] / ms,
                         "peak_source": "sb_measure_imad_peak: dependent-free mad.lo.u32, measured in this run"},
            "sharded_proof": {"ranks": world, "same_proof_on_every_rank": same, "k1_stores_into_peer_memory": bool(group.fused_ok and fused),
                              "phase_ms_rank0": proofs[-1].phase_ms, "library_stage_ms_rank0": {k: round(float(v), 2) for k, v in proofs[-1].timings.items()},
                              "kernel_ms_rank0": kern,
                              "note": "sb_prove on a multi-GPU group: column-sharded LDE storing into the owners' row buffers over NVLink, "
                                      "row-sharded leaf hashing and quotient, digests / halo rows / quotient values / openings / FRI "
                                      "combine partials / query rows exchanged with NCCL inside libstarkyb200; quotient commitment, "
                                      "transcript, FRI rounds and proof of work redundantly on every rank"},
        }
        del local_host, local_dev
        torch.cuda.empty_cache()
        if extras:
            # replicas: one independent proof per GPU (the reference's seven proofs are independent), weak scaling
            def replicas():
                rm, rtrace, _, _, _ = measure_stark(ctx, sb, args.stark, max(1, min(args.steps, 3)), 1, timed, 0xB2000000 + info.stark_id + 1000 * rank,
                                                    pageable_leg=False)
                del rtrace
                return {"ms_per_proof_per_gpu": rm["ms"], "ms_per_proof_whole_box": rm["ms"] / world, "e2e_ms_per_proof_whole_box": rm["ms_e2e"] / world,
                        "scaling": "weak", "note": "every rank proves its own trace, no data-path collective"}
            line["replicas"] = optional(replicas, world)
            also = {}
            for name in also_names:
                def leg(name=name):
                    ai = sb.STARKS[name]
                    ap_ = sb.standard_params(ai.stark_id, ai.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
                    a0, ag = group.column_slice(ap_)
                    # every rank draws its own column slice (seed + rank): no 4.8 GB trace is replicated on the host
                    arng = np.random.Generator(np.random.PCG64(0xB2100000 + ai.stark_id + 1000 * rank))
                    ahost = torch.from_numpy(arng.integers(0, 1 << 32, (ag, ai.num_rows), dtype=np.uint64).view(np.int64)).pin_memory()
                    adev = ahost.cuda()
                    apis = np.random.Generator(np.random.PCG64(0xB2100000 + ai.stark_id)).integers(0, 1 << 32, ai.public_inputs, dtype=np.uint64)
                    fn = lambda: group.prove(ap_, adev.data_ptr(), apis, on_device=True, fused=fused)
                    timed(fn, 2)
                    k = max(1, min(args.steps, 3))
                    adt, aproofs = timed(fn, k)
                    akern = {kk: ctx.stage_ms(kk) for kk in ("lde", "leaf_hash", "merkle", "quotient")}
                    fe2e = lambda: group.prove(ap_, ahost.data_ptr(), apis, on_device=False, fused=fused)
                    timed(fe2e, 2)
                    adt2, _ = timed(fe2e, k)
                    return {"workload": WORKLOADS[name], "value": 1e3 * adt / k, "unit": "ms", "ranks": world,
                            "e2e": {"value": 1e3 * adt2 / k, "unit": "ms", "h2d_bytes_per_step": 8 * ag * ai.num_rows, "d2h_bytes_per_step": int(aproofs[-1].layout.total_words) * 8},
                            "same_proof_on_every_rank": multi.same_on_every_rank(aproofs[-1].words),
                            "phase_ms_rank0": aproofs[-1].phase_ms, "kernel_ms_rank0": akern}
                leg.__name__ = "also_sharded_" + name
                also[name] = optional(leg, world)
            line["also"] = also
        group.close()

    if extras:
        # ---- several proofs in flight on one GPU (one context and one host thread per proof): the leaf sponge of these shapes
        # is latency-bound and the host transcript is a strictly sequential sponge, so concurrent proofs fill each other's gaps
        more = [sb.Context(local_rank) for _ in range(3)]

        def in_flight_leg():
            hinfo = sb.STARKS["miller_loop"]            # the latency-bound shape: FinalExp fills the GPU on its own
            hp = sb.standard_params(hinfo.stark_id, hinfo.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
            htrace, hpis = synthetic(hinfo, 0xB2000000 + hinfo.stark_id + 1000 * rank)
            hhost = torch.from_numpy(htrace.view(np.int64)).pin_memory()
            del htrace
            hptr = hhost.data_ptr()
            k = max(1, min(args.steps, 4))
            for c in [ctx] + more:
                c.prove(hp, hptr, hpis, sb.TraceLayout.COLMAJOR_U64)

            def run_with(contexts):
                def run():
                    def worker(c):
                        for _ in range(k):
                            c.prove(hp, hptr, hpis, sb.TraceLayout.COLMAJOR_U64)
                    ts = [threading.Thread(target=worker, args=(c,)) for c in contexts]
                    for t in ts:
                        t.start()
                    for t in ts:
                        t.join()
                return run
            d2, _ = timed(run_with([ctx, more[0]]), 1)
            d4, _ = timed(run_with([ctx] + more), 1)
            return {"workload": WORKLOADS["miller_loop"], "2": {"ms_per_proof": 1e3 * d2 / (2 * k) / world}, "4": {"ms_per_proof": 1e3 * d4 / (4 * k) / world},
                    "note": "k contexts per GPU, one host thread each, end to end from pinned host memory, every GPU busy: the "
                            "latency-bound leaf sponge and the sequential host transcript of one proof overlap the kernels of the others"}
        inflight = optional(in_flight_leg, world)

        # ---- BASELINE configs[4]: the seven proofs of one BLS signature verification over all ranks ----
        def full_leg():
            if args.no_full_set:
                return None
            fe_ranks, per_rank = full_set_plan(world, args.fe_over)
            mine = per_rank[rank]
            sharded_job, sub = None, None
            if fe_ranks:
                from starky_bls12_381_b200 import multi
                sub = multi.Group.from_torch(ctx, rank, world, local_rank, ranks=fe_ranks)      # collective over all ranks
                if rank in fe_ranks:
                    fi = sb.STARKS["final_exp"]
                    sp = sb.standard_params(fi.stark_id, fi.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
                    s0, sg = sub.column_slice(sp)
                    srng = np.random.Generator(np.random.PCG64(0xB2400000 + fe_ranks.index(rank)))
                    slocal = torch.from_numpy(srng.integers(0, 1 << 32, (sg, fi.num_rows), dtype=np.uint64).view(np.int64)).pin_memory()
                    spis = np.random.Generator(np.random.PCG64(0xB2400099)).integers(0, 1 << 32, fi.public_inputs, dtype=np.uint64)
                    sharded_job = lambda: sub.prove(sp, slocal.data_ptr(), spis, on_device=False, fused=not args.no_fused)
            # the contexts of the whole proofs: not the one the FinalExp shard of this rank is running on
            dt_full, per = run_full_set(sb, more if sharded_job is not None else [ctx] + more, mine, rank, timed, sharded_job)
            out = {"workload": "2 x PairingPrecomp + 2 x MillerLoop + FP12Mul + FinalExp + ECCAgg (BASELINE configs[4]), synthetic traces, "
                               "end to end from pinned host memory", "gpus": world, "ms": 1e3 * dt_full,
                   "assignment": per_rank, "final_exp_sharded_over_ranks": fe_ranks,
                   "final_exp_k1_stores_into_peer_memory": bool(sub is not None and sub.member and sub.fused_ok and not args.no_fused) if fe_ranks else None,
                   "rank0_proof_ms": {("%s#%d" % (k, i)): round(v, 2) for i, (k, v) in enumerate(per)},
                   "plan": args.fe_over if fe_ranks else "whole proofs",
                   "note": "from 4 GPUs on FinalExp is sharded over the GPUs named above and the other six proofs are assigned longest-first "
                           "(plan all: to the same GPUs, running next to the FinalExp shards on other contexts; plan half: to the other "
                           "GPUs); below 4 GPUs whole proofs longest-first; per GPU sb_prove_batch schedules the proofs; ms = makespan, "
                           "max over ranks"}
            if sub is not None:
                sub.close()
            return out
        full = optional(full_leg, world)

        def bundled_leg():
            """BASELINE configs[4] on the reference's own inputs (main.rs:8-55): tests/golden/bundled_inputs.json through
            bls.prepare and the witness generators, VALID traces, flags = 0 (a non-divisible quotient would be an error)."""
            if not args.bundled or world != 1:
                return None
            from starky_bls12_381_b200 import airfiles, bundled
            from starky_bls12_381_b200.binding import prove_batch
            inp = bundled.load_inputs()
            t0 = time.perf_counter()
            js = bundled.jobs_native(inp)            # the library's C++ witness generators: row-major u32 traces
            gen_s = time.perf_counter() - t0
            keep, batch = [], []
            for name, trace, pis in js:
                host = torch.from_numpy(trace.view(np.int32)).pin_memory()
                keep.append(host)
                batch.append((sb.standard_params(sb.STARKS[name].stark_id, trace.shape[0].bit_length() - 1), host.data_ptr(),
                              sb.TraceLayout.ROWMAJOR_U32, pis))
            del js
            ctxs = [ctx] + more
            prove_batch(ctxs, batch)
            prove_batch(ctxs, batch)
            dtb, outs = timed(lambda: prove_batch(ctxs, batch), 1)
            res = outs[0]
            accepted = None
            if not args.no_cpu_baseline:
                O, _ = oracle_all_threads()
                accepted = all(O.verify(airfiles.air_path(nm, "air"), O.Params.from_buffer_copy(bytes(b[0])), r[0].words) == 0
                               for nm, b, r in zip(bundled.ORDER, batch, res))
            return {"workload": "the seven proofs of the reference's bundled run (light_client_update_period_1052 public keys, "
                                "_1053 sync aggregate and attested header; main.rs:8-55), valid traces, end to end from pinned host memory",
                    "data": "bundled", "ms": 1e3 * dtb, "proof_ms": {("%s#%d" % (nm, i)): round(r[1], 2) for i, (nm, r) in enumerate(zip(bundled.ORDER, res))},
                    "final_exp_output_is_one": bool([int(v) for v in batch[-1][3][-144:]] == [1] + [0] * 143),
                    "every_proof_accepted_by_the_oracle_verifier": accepted, "host_trace_generation_s": round(gen_s, 2),
                    "trace_generation": "csrc/witness.cpp (sb_witness_*), row-major u32, outside the timed region"}
        bundled_out = optional(bundled_leg, world)
        for c in more:
            c.close()
        line["in_flight"] = inflight
        line["full_bls_set"] = full
        if bundled_out is not None:
            line["bundled_bls_set"] = bundled_out

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
