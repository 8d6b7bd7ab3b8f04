#!/usr/bin/env python3
"""bench.py -- headline benchmark: starky prove ms/STARK on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--stark pairing_precomp] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full proof (sb_prove: trace LDE -> Poseidon Merkle -> quotient -> FRI -> proof in host memory) of one
synthetic trace of the workload stark.  At N GPUs every rank proves its own trace (the reference's seven proofs are
independent, SURVEY 8e "proof-level parallelism"), so scaling is weak and `value` = wall ms per step / N.
  value : trace already resident in HBM when the timed region starts (layout DEVICE_COLMAJOR_U64)
  e2e   : through the plugin call with the trace in pinned HOST memory (H2D of the trace and D2H of the proof inside)
  --impl reference : the CPU restatement of the reference's prover (oracle/, all host threads) on the same workload.
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
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
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
    """Smallest shard plan that shard_plan accepts at every power-of-two world size up to 64: the symmetric-memory probe only
    needs *a* valid plan, and a plan the planner itself rejects (fewer than 32 LDE positions per rank) would read as
    'no symmetric memory' and silently demote the fused K1 to the all-to-all."""
    from starky_bls12_381_b200.sharded import shard_plan
    return shard_plan(8, 10, 1, world)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, info, rank, world):
    """CPU arm: the oracle restatement of starky::prover::prove, all host threads, same stark and shape."""
    if rank != 0:
        return
    import oracle_lib as O
    from starky_bls12_381_b200 import airfiles
    O.build()
    flat = airfiles.air_path(args.stark, "air")
    trace, pis = synthetic(info, 0xB2000000 + info.stark_id)
    p = O.make_params(stark_id=info.stark_id, log_n=info.num_rows.bit_length() - 1, n_cols=info.columns,
                      n_pis=info.public_inputs, degree=info.constraint_degree, rate_bits=info.rate_bits, flags=1)
    cores = O.lib().orc_num_threads()
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rc, _ = O.prove(flat, p, trace, pis)
        assert rc == 0, O.err()
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    line = {"impl": "reference", "metric": "starky_prove_ms_per_stark", "value": ms, "unit": "ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (Goldilocks)", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.stark], "trace": "uniform u32 cells, seeded PCG64",
                       "note": "CPU restatement of the reference algorithm (oracle/), not the Rust binary: no cargo in the image"},
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": "one full proof per step"},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_also(ctx, sb, name):
    """One further stark of BASELINE.json's metric (MillerLoop, FinalExp), proved after the headline workload: 1 warm-up +
    2 timed proofs with the trace resident, 1 end to end from pinned host memory."""
    import torch
    info = sb.STARKS[name]
    p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
    trace, pis = synthetic(info, 0xB2000000 + info.stark_id)
    host = torch.from_numpy(trace.view(np.int64)).pin_memory()
    del trace
    ctx.trace_upload(p, host.data_ptr())
    ctx.prove(p, None, pis, sb.TraceLayout.DEVICE_COLMAJOR_U64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    proofs = [ctx.prove(p, None, pis, sb.TraceLayout.DEVICE_COLMAJOR_U64) for _ in range(2)]
    ms = 1e3 * (time.perf_counter() - t0) / 2
    t0 = time.perf_counter()
    ctx.prove(p, host.data_ptr(), pis, sb.TraceLayout.COLMAJOR_U64)
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    C, n, N = info.columns, info.num_rows, info.num_rows << info.rate_bits
    kern = {k: ctx.stage_ms(k) for k in ("lde", "leaf_hash", "merkle", "quotient")}
    return {"workload": WORKLOADS[name], "ms": ms, "ms_e2e": ms_e2e, "stage_ms": {k: float(v) for k, v in proofs[-1].timings.items()},
            "kernel_ms": kern, "lde_merkle_gbs": 8.0 * C * N / ((kern["lde"] + kern["leaf_hash"] + kern["merkle"]) * 1e-3) / 1e9,
            "leaf_hash_mperm_s": (-(-C // 8) * N) / (kern["leaf_hash"] * 1e-3) / 1e6, "h2d_bytes": 8 * C * n}


FULL_SET = ["final_exp", "miller_loop", "miller_loop", "pairing_precomp", "pairing_precomp", "ecc_agg", "fp12_mul"]
FULL_SET_COST = {"final_exp": 650.0, "miller_loop": 225.0, "pairing_precomp": 75.0, "ecc_agg": 35.0, "fp12_mul": 6.0}


def full_set_assignment(world):
    """BASELINE configs[4]: the seven starky proofs of one BLS signature verification (aggregate_proof.rs:279-370 proves
    them one after the other; they are independent once the native inputs are known).  Longest-processing-time-first
    onto `world` GPUs; returns one list of stark names per rank."""
    loads, out = [0.0] * world, [[] for _ in range(world)]
    for name in sorted(FULL_SET, key=lambda k: -FULL_SET_COST[k]):
        g = min(range(world), key=lambda r: loads[r])
        out[g].append(name)
        loads[g] += FULL_SET_COST[name]
    return out


def full_set_plan(world):
    """How the seven proofs are laid over `world` GPUs.  From four GPUs on, the critical path (FinalExp, ~630 ms on one GPU)
    is sharded over half of them (SURVEY 8e "proof-level parallelism on top") and the other six proofs share the rest;
    below that whole proofs are assigned longest-first.  Returns (ranks of the sharded FinalExp proof or [], per-rank lists)."""
    if world < 4:
        return [], full_set_assignment(world)
    k = world // 2
    rest = full_set_assignment_of([x for x in FULL_SET if x != "final_exp"], world - k)
    return list(range(k)), [[] for _ in range(k)] + rest


def full_set_assignment_of(names, world):
    loads, out = [0.0] * world, [[] for _ in range(world)]
    for name in sorted(names, key=lambda k: -FULL_SET_COST[k]):
        g = min(range(world), key=lambda r: loads[r])
        out[g].append(name)
        loads[g] += FULL_SET_COST[name]
    return out


def run_full_set(sb, contexts, names, rank, timed, sharded_job=None):
    """Proves `names` on this rank's GPU, end to end from pinned host memory, with len(contexts) proofs in flight (one host
    thread per context, work queue in cost order).  Returns (seconds max over ranks, per-proof ms)."""
    import torch
    jobs = []
    for i, name in enumerate(names):
        info = sb.STARKS[name]
        trace, pis = synthetic(info, 0xB2300000 + 16 * rank + i)
        host = torch.from_numpy(trace.view(np.int64)).pin_memory()
        del trace
        p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
        jobs.append((name, p, host, pis))
    lock, per = threading.Lock(), []
    # Phase A: the latency-bound proofs (few leaves: MillerLoop, PairingPrecomp, FP12Mul -- the leaf sponge is a long
    # sequential chain, the host transcript is a large share) up to four in flight, so that the transcript of one overlaps the
    # kernels of the other; phase B: the throughput-bound ones (FinalExp, ECCAgg: 32768 leaves fill the GPU) one at a
    # time.  Measured alternatives: FinalExp next to MillerLoop slows the latter to 937 ms (from 225); five latency-bound
    # proofs in flight serialise on the one-block-per-SM leaf sponge (893 ms for the phase instead of ~500).
    few = lambda j: (sb.STARKS[j[0]].num_rows << sb.STARKS[j[0]].rate_bits) <= 64 * 148
    phases = [([j for j in jobs if few(j)], contexts), ([j for j in jobs if not few(j)], contexts[:1])]

    def go():
        if sharded_job is not None:
            t0 = time.perf_counter()
            sharded_job()
            per.append(("final_exp(sharded)", 1e3 * (time.perf_counter() - t0)))
        for phase_jobs, phase_ctx in phases:
            # job i of a phase always runs on context i % len(contexts): the warm-up pass then sizes exactly the device
            # buffers (and loads the constraint programs) the timed pass needs -- with a shared work queue a context can
            # meet its largest shape for the first time inside the timed pass and pay a multi-GB cudaMalloc there
            lanes = full_set_assignment_of([j[0] for j in phase_jobs], len(phase_ctx))     # longest-first over the contexts
            pool = list(phase_jobs)
            mine_jobs = []
            for lane in lanes:
                got = []
                for nm in lane:
                    i = next(i for i, j in enumerate(pool) if j[0] == nm)
                    got.append(pool.pop(i))
                mine_jobs.append(got)

            def worker(k, c):
                for name, p, host, pis in mine_jobs[k]:
                    t0 = time.perf_counter()
                    c.prove(p, host.data_ptr(), pis, sb.TraceLayout.COLMAJOR_U64)
                    with lock:
                        per.append((name, 1e3 * (time.perf_counter() - t0)))
            ts = [threading.Thread(target=worker, args=(k, c)) for k, c in enumerate(phase_ctx)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
    go()                 # warm-up: buffers of every shape allocated, constraint programs loaded
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
    ap.add_argument("--stark", default="pairing_precomp", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharded-stark", default="final_exp",
                    help="shape of the second sharded trace commitment (column slices drawn per rank); '' to skip")
    ap.add_argument("--no-fused", action="store_true", help="sharded legs: NCCL all-to-all after K1 instead of K1 storing into peer memory")
    ap.add_argument("--no-full-set", action="store_true", help="skip the 7-proof BLS set (BASELINE configs[4])")
    ap.add_argument("--also", default="miller_loop,final_exp",
                    help="N=1 only: further starks proved once each after the headline workload (reported under 'also')")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import starky_bls12_381_b200 as sb
    info = sb.STARKS[args.stark]
    if args.impl == "reference":
        return run_reference(args, info, rank, world)

    import torch
    import torch.distributed as dist
    from starky_bls12_381_b200 import airfiles
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    airfiles.air_path(args.stark, "airbin")
    ctx = sb.Context(local_rank)
    p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
    trace, pis = synthetic(info, 0xB2000000 + info.stark_id + 1000 * rank)
    # pinned host copy for the end-to-end leg; resident device copy for the kernel leg
    host = torch.from_numpy(trace.view(np.int64)).pin_memory()
    host_ptr = host.data_ptr()
    ctx.trace_upload(p, host_ptr)
    C, n, N = info.columns, info.num_rows, info.num_rows << info.rate_bits

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        out = [fn() for _ in range(steps)]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, out

    resident = lambda: ctx.prove(p, None, pis, sb.TraceLayout.DEVICE_COLMAJOR_U64)
    e2e_fn = lambda: ctx.prove(p, host_ptr, pis, sb.TraceLayout.COLMAJOR_U64)
    for _ in range(args.warmup):
        resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.kernel_launches()
    dt, proofs = timed(resident, args.steps)
    launches = ctx.kernel_launches() - l0
    stage = {k: float(np.mean([pr.timings[k] for pr in proofs])) for k in proofs[0].timings}
    kern = {k: ctx.stage_ms(k) for k in ("lde", "leaf_hash", "merkle", "quotient")}
    e2e_fn()
    dt_e2e, proofs_e2e = timed(e2e_fn, args.steps)
    # ---- ONE trace commitment sharded over all ranks (column-sharded LDE -> all-to-all -> row-sharded leaf
    # hashing -> digest all-gather -> tree), the LDE+Merkle GB/s half of BASELINE.json's metric ----
    from starky_bls12_381_b200.sharded import GpuBackend, TorchGroup, commit_sharded, prove_sharded, quotient_sharded, shard_plan
    plan = shard_plan(C, info.num_rows.bit_length() - 1, info.rate_bits, world)
    base_trace = trace if rank == 0 else synthetic(info, 0xB2000000 + info.stark_id)[0]
    c0, cg = plan.col_start[rank], plan.col_count[rank]
    local = torch.from_numpy(np.ascontiguousarray(base_trace[c0:c0 + cg]).view(np.int64)).cuda()
    backend = GpuBackend(ctx, p)
    pis0 = pis if rank == 0 else np.random.Generator(np.random.PCG64(7)).integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    if world > 1:                      # every rank must evaluate with rank 0's public inputs
        t_pis = torch.from_numpy(pis0.view(np.int64).copy()).cuda()
        dist.broadcast(t_pis, 0)
        pis0 = t_pis.cpu().numpy().view(np.uint64).copy()

    def sharded_fn():
        out = commit_sharded(backend, plan, rank, local)
        out["quotient"] = quotient_sharded(backend, plan, rank, out["rows"], out["cap"], pis0)
        return out
    for _ in range(args.warmup):
        sh = sharded_fn()
    dt_sh, sh_out = timed(sharded_fn, args.steps)
    sh_cap = sh_out[-1]["cap"]
    sh_q = sh_out[-1]["quotient"]["q"].cpu().numpy().view(np.uint64).copy()
    sh_q_ms = ctx.stage_ms("quotient")
    sh_out_alphas = sh_out[-1]["quotient"]["alphas"]
    del sh_out, sh, local
    torch.cuda.empty_cache()
    # ---- the same sharded commitment on the FinalExp shape (BASELINE configs[3]; throughput-bound, the shape that scales):
    # every rank draws its own column slice (seed + rank), so no 4.8 GB trace is replicated on the host ----
    def fe_leg():
        fe = None
        if args.sharded_stark and args.sharded_stark != args.stark:
            fi = sb.STARKS[args.sharded_stark]
            fp = sb.standard_params(fi.stark_id, fi.num_rows.bit_length() - 1)
            fplan = shard_plan(fi.columns, fi.num_rows.bit_length() - 1, fi.rate_bits, world)
            fcg = fplan.col_count[rank]
            frng = np.random.Generator(np.random.PCG64(0xB2100000 + fi.stark_id + 1000 * rank))
            flocal = torch.from_numpy(frng.integers(0, 1 << 32, (fcg, fi.num_rows), dtype=np.uint64).view(np.int64)).cuda()
            fbackend = GpuBackend(ctx, fp)
            airfiles.air_path(args.sharded_stark, "airbin")
            fpis = np.random.Generator(np.random.PCG64(0xB2100000 + fi.stark_id)).integers(0, 1 << 32, fi.public_inputs, dtype=np.uint64)

            fused = world > 1 and not args.no_fused
            if fused:
                # symmetric memory needs peer access between the GPUs of the box; probe it once and let every rank agree, so
                # that a box without it falls back to the NCCL all-to-all instead of losing the bench line
                ok = 1
                try:
                    probe = TorchGroup(world, rank).symmetric_rows(probe_plan(world), torch.device("cuda", local_rank))
                    probe[2]()
                    del probe
                except Exception as e:          # noqa: BLE001
                    ok = 0
                    print("symmetric memory unavailable on rank %d: %r" % (rank, e), file=sys.stderr, flush=True)
                flag = torch.tensor([ok], device="cuda")
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                fused = bool(flag.item())

            def ffn():
                out = commit_sharded(fbackend, fplan, rank, flocal, fused=fused)
                qq = quotient_sharded(fbackend, fplan, rank, out["rows"], out["cap"], fpis)
                return out["cap"], qq["q"][:, :4]
            ffn()
            dt_fe, _ = timed(ffn, max(1, args.steps - 1))
            fN = fi.num_rows << fi.rate_bits
            ms_fe = 1e3 * dt_fe / max(1, args.steps - 1)
            fe = {"workload": WORKLOADS[args.sharded_stark], "ranks": world, "ms": ms_fe,
                  "lde_merkle_gbs": 8.0 * fi.columns * fN / (ms_fe * 1e-3) / 1e9, "a2a_bytes_out_per_rank": fplan.a2a_bytes_out(0),
                  "leaf_hash_ms_rank0": ctx.stage_ms("leaf_hash"), "lde_ms_rank0": ctx.stage_ms("lde"),
                  "quotient_ms_rank0": ctx.stage_ms("quotient"), "k1_stores_into_peer_memory": fused,
                  "note": "column-sharded LDE -> all-to-all (or, fused: K1 stores into the owners' row buffers over NVLink) -> row-sharded leaf hashing -> digest all-gather -> tree -> alphas -> "
                          "halo row exchange -> row-sharded quotient -> all-gather of the 2 x N quotient values"}
            # the WHOLE proof of that sharded trace: sb_prove_sharded on every rank, the five distributed steps (commitment,
            # quotient, openings, FRI batch combine, query rows) as NCCL collectives; every rank ends with the same proof
            fp_inv = sb.standard_params(fi.stark_id, fi.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
            pbackend = GpuBackend(ctx, fp_inv)
            comm = TorchGroup(world, rank)
            pfn = lambda: prove_sharded(pbackend, fplan, rank, flocal, fpis, comm=comm, fused=fused)
            pfn()
            dt_fp, fproofs = timed(pfn, max(1, args.steps - 1))
            caps = torch.from_numpy(fproofs[-1].words[:64].view(np.int64).copy()).cuda()
            same = True
            if world > 1:
                allc = [torch.empty_like(caps) for _ in range(world)]
                dist.all_gather(allc, caps)
                same = all(bool(torch.equal(allc[0], c)) for c in allc)
            fe["sharded_proof"] = {"ms": 1e3 * dt_fp / max(1, args.steps - 1), "ranks": world, "proof_words": int(fproofs[-1].layout.total_words),
                                   "same_proof_on_every_rank": same, "hook_ms_rank0": {k: round(v, 2) for k, v in fproofs[-1].hook_ms.items()},
                                   "library_stage_ms_rank0": {k: round(float(v), 2) for k, v in fproofs[-1].timings.items()},
                                   "note": "one FinalExp-shaped proof, trace sharded over all ranks (sb_prove_sharded): commitment + quotient as "
                                           "above, openings from column-sharded coefficient slices (all-gather), FRI batch combine (per-rank "
                                           "partial sums, all-gather + add), query rows from their owners; quotient commitment, transcript, "
                                           "FRI rounds and proof of work redundantly on every rank"}
            del flocal, fbackend, pbackend, fproofs
            torch.cuda.empty_cache()
        return fe

    fe = optional(fe_leg, world)
    # ---- several proofs in flight on one GPU (one context and one host thread per proof): the leaf sponge of these shapes
    # is latency-bound (one 32-leaf group per SM) and the host transcript is a strictly sequential sponge (~1 us per
    # permutation), so concurrent proofs fill each other's gaps -- how a scheduler for the reference's seven independent
    # proofs runs them.  End to end: every proof is taken from pinned host memory. ----
    ctx2 = sb.Context(local_rank)
    more = [sb.Context(local_rank) for _ in range(2)]
    for c in [ctx2] + more:
        c.prove(p, host_ptr, pis, sb.TraceLayout.COLMAJOR_U64)

    def in_flight(contexts):
        def run():
            def worker(c):
                for _ in range(args.steps):
                    c.prove(p, host_ptr, pis, sb.TraceLayout.COLMAJOR_U64)
            ts = [threading.Thread(target=worker, args=(c,)) for c in contexts]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return run
    dt_pipe2, _ = timed(in_flight([ctx, ctx2]), 1)
    dt_pipe4, _ = timed(in_flight([ctx, ctx2] + more), 1)
    # ---- BASELINE configs[4]: the seven proofs of one BLS signature verification over all ranks, two in flight per GPU ----
    def full_leg():
        full = None
        if not args.no_full_set:
            for nm in set(FULL_SET):
                airfiles.air_path(nm, "airbin")
            fe_ranks, per_rank = full_set_plan(world)
            mine = per_rank[rank]
            sharded_job, fe_fused = None, None
            if fe_ranks:
                grp = dist.new_group(ranks=fe_ranks)                      # collective: every rank calls it
                if rank in fe_ranks:
                    fi = sb.STARKS["final_exp"]
                    sp = sb.standard_params(fi.stark_id, fi.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
                    splan = shard_plan(fi.columns, fi.num_rows.bit_length() - 1, fi.rate_bits, len(fe_ranks))
                    sr = fe_ranks.index(rank)
                    srng = np.random.Generator(np.random.PCG64(0xB2400000 + sr))
                    slocal = torch.from_numpy(srng.integers(0, 1 << 32, (splan.col_count[sr], fi.num_rows), dtype=np.uint64).view(np.int64)).pin_memory()
                    spis = np.random.Generator(np.random.PCG64(0xB2400099)).integers(0, 1 << 32, fi.public_inputs, dtype=np.uint64)
                    sbackend, scomm = GpuBackend(ctx, sp), TorchGroup(len(fe_ranks), sr, grp)
                    fs_fused = 0 if args.no_fused else 1
                    if fs_fused:
                        # same probe as the sharded legs, agreed inside the FinalExp sub-group only
                        try:
                            probe = scomm.symmetric_rows(probe_plan(len(fe_ranks)), torch.device("cuda", local_rank))
                            probe[2]()
                            del probe
                        except Exception as e:          # noqa: BLE001
                            fs_fused = 0
                            print("symmetric memory unavailable in the FinalExp sub-group on rank %d: %r" % (rank, e), file=sys.stderr, flush=True)
                        flag = torch.tensor([fs_fused], device="cuda")
                        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=grp)
                        fs_fused = int(flag.item())
                    fe_fused = bool(fs_fused)
                    sharded_job = lambda: prove_sharded(sbackend, splan, sr, slocal, spis, comm=scomm, fused=fe_fused)
            dt_full, per = run_full_set(sb, [ctx, ctx2] + more, mine, rank, timed, sharded_job)
            full = {"workload": "2 x PairingPrecomp + 2 x MillerLoop + FP12Mul + FinalExp + ECCAgg (BASELINE configs[4]), synthetic traces, "
                                "end to end from pinned host memory", "gpus": world, "ms": 1e3 * dt_full,
                    "assignment": per_rank, "final_exp_sharded_over_ranks": fe_ranks, "final_exp_k1_stores_into_peer_memory": fe_fused, "rank0_proof_ms": {("%s#%d" % (k, i)): round(v, 2) for i, (k, v) in enumerate(per)},
                    "note": "from 4 GPUs on FinalExp is sharded over half of them (sb_prove_sharded) and the other six proofs share the rest; otherwise "
                            "longest-first assignment of whole proofs to GPUs; per GPU the latency-bound proofs (<= 9472 leaves) run up to four in "
                            "flight first, then the throughput-bound ones one at a time; ms = makespan, max over ranks"}
        return full

    full = optional(full_leg, world)
    ctx2.close()
    for c in more:
        c.close()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = 1e3 * dt / args.steps
    ms_e2e = 1e3 * dt_e2e / args.steps
    proof_bytes = int(proofs[0].layout.total_words) * 8

    if rank == 0:
        hbm_peak, peak_src = peaks()
        imad = ctx.measure_imad_peak()
        perms = -(-C // 8) * N + (N - 16)
        t_hash = (kern["leaf_hash"] + kern["merkle"]) * 1e-3
        t_lde = kern["lde"] * 1e-3
        t_q = kern["quotient"] * 1e-3
        lde_bytes = 8.0 * C * (n + n + N)           # trace read + coefficients kept + LDE written
        roof = {
            # dominant kernel of the step: the Poseidon leaf sponge (integer-pipe bound)
            "kernel": ("leaf_sponge_sp_kernel" if N <= 64 * 148 else "leaf_sponge_dp_kernel") + "+merkle_level_kernel", "bound": "imad",
            "achieved": perms * U32_MACS_PER_PERM / t_hash / 1e9, "peak": imad["mad_lo_u32_gops"], "unit": "Gop/s (u32 multiply-add)",
            "frac": perms * U32_MACS_PER_PERM / t_hash / 1e9 / imad["mad_lo_u32_gops"],
            # dram__bytes_read.sum + dram__bytes_write.sum of the leaf sponge from the committed ncu --set full capture
            # (profiles/r1_top_kernels_pairing_precomp_final.txt); algorithmic = 8 C N = 962.6 MB
            "traffic": 975198208 if args.stark == "pairing_precomp" else None,
            "peak_source": "sb_measure_imad_peak: dependent-free mad.lo.u32, measured in this run",
            "share_of_step": (kern["leaf_hash"] + kern["merkle"]) / ms_step,
            "stages": {
                "lde": {"bound": "hbm", "achieved": lde_bytes / t_lde / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": lde_bytes / t_lde / 1e9 / hbm_peak, "ms": kern["lde"], "peak_source": peak_src},
                "merkle_hbm_view": {"bound": "hbm", "achieved": 8.0 * C * N / t_hash / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": 8.0 * C * N / t_hash / 1e9 / hbm_peak, "ms": kern["leaf_hash"] + kern["merkle"]},
                "quotient": {"bound": "imad", "achieved": K_CONSTRAINTS[args.stark] * N * U32_MACS_PER_CONSTRAINT / t_q / 1e9,
                             "peak": imad["mad_lo_u32_gops"], "unit": "Gop/s (u32 multiply-add)",
                             "frac": K_CONSTRAINTS[args.stark] * N * U32_MACS_PER_CONSTRAINT / t_q / 1e9 / imad["mad_lo_u32_gops"],
                             "ms": kern["quotient"]},
            },
        }
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            import oracle_lib as O
            O.build()
            flat = airfiles.air_path(args.stark, "air")
            op = O.Params.from_buffer_copy(bytes(p))
            t0 = time.perf_counter()
            rc, words = O.prove(flat, op, trace, pis)
            cpu_ms = 1e3 * (time.perf_counter() - t0)
            same = bool(rc == 0 and np.array_equal(words, proofs[0].words))
            cpu = {"value": cpu_ms, "unit": "ms", "cores": int(O.lib().orc_num_threads()), "kind": "port",
                   "sample": "one full proof of the same trace", "proof_bit_identical_to_gpu": same}
        ms_sh = 1e3 * dt_sh / args.steps
        from helpers import pos_to_natural
        ctx.lde_commit(p, trace, want_lde=False, want_digests=False)
        q_one = ctx.quotient_values(p, pis0, sh_out_alphas)[:, pos_to_natural(p.log_n, p.rate_bits)]
        sh_q_same = bool(np.array_equal(q_one, sh_q))
        sharded = {"ms": ms_sh, "lde_merkle_gbs": 8.0 * C * N / (ms_sh * 1e-3) / 1e9, "ranks": world,
                   "a2a_bytes_out_per_rank": plan.a2a_bytes_out(0), "digest_allgather_bytes": 32 * N,
                   "cap_equals_single_gpu_path": bool(np.array_equal(sh_cap, proofs[0].words[:4 * (1 << p.cap_height)].reshape(-1, 4))),
                   "quotient_ms_rank0": sh_q_ms, "quotient_equals_single_gpu_path": sh_q_same,
                   "note": "one trace, columns sharded for K1, rows sharded for K2 and K4 (quotient), NCCL all-to-all + "
                           "all-gathers (digests, halo rows, quotient values) in the timed region"}
        also = {}
        if world == 1 and args.also:
            for name in [x for x in args.also.split(",") if x and x != args.stark]:
                try:
                    also[name] = run_also(ctx, sb, name)
                except Exception as e:     # keep the headline line alive
                    also[name] = {"error": repr(e)}
        line = {
            "metric": "starky_prove_ms_per_stark", "value": ms_step / world, "unit": "ms", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks)", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.stark], "trace": "uniform u32 cells, seeded PCG64, one trace per rank",
                       "parallelism": "proofs: one independent proof per GPU (the reference's 7 proofs are independent), no "
                                      "data-path collective; 'sharded_commit' is the column->row sharded trace commitment over all ranks",
                       "l2": "inputs larger than L2 (trace %.0f MB, LDE %.0f MB)" % (8e-6 * C * n, 8e-6 * C * N),
                       "lde_merkle_gbs": 8.0 * C * N / ((kern["lde"] + kern["leaf_hash"] + kern["merkle"]) * 1e-3) / 1e9},
            "e2e": {"value": ms_e2e / world, "unit": "ms", "h2d_bytes_per_step": 8 * C * n + 8 * info.public_inputs,
                    "d2h_bytes_per_step": proof_bytes},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "stage_ms": stage, "kernel_ms": kern, "sharded_commit": sharded, "sharded_commit_scaling_shape": fe, "also": also,
            "full_bls_set": full,
            "in_flight": {"2": {"ms_per_proof": 1e3 * dt_pipe2 / (2 * args.steps) / world}, "4": {"ms_per_proof": 1e3 * dt_pipe4 / (4 * args.steps) / world},
                          "note": "k contexts per GPU, one host thread each, end to end from pinned host memory: the latency-bound leaf "
                                  "sponge and the sequential host transcript of one proof overlap the kernels of the others"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
