"""sb_prove_batch of a list of starks on one GPU (4 contexts), synthetic traces from pinned host memory: makespan and
per-proof wall times.   python tools/perf/batch_once.py miller_loop,pairing_precomp,ecc_agg [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import starky_bls12_381_b200 as sb
from starky_bls12_381_b200.binding import prove_batch

names = sys.argv[1].split(",")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctxs = [sb.Context(0) for _ in range(4)]
jobs, keep = [], []
for i, name in enumerate(names):
    info = sb.STARKS[name]
    rng = np.random.default_rng(100 + i)
    host = torch.from_numpy(rng.integers(0, 1 << 32, (info.columns, info.num_rows), dtype=np.uint64).view(np.int64)).pin_memory()
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
    keep.append(host)
    jobs.append((p, host.data_ptr(), sb.TraceLayout.COLMAJOR_U64, pis))
for r in range(reps + 2):
    t0 = time.perf_counter()
    res = prove_batch(ctxs, jobs)
    dt = 1e3 * (time.perf_counter() - t0)
    if r >= 2:
        print("env MIX=%s NO_STREAM=%s: makespan %.1f ms  " % (os.environ.get("SB_SCHED_MIX"), os.environ.get("SB_NO_STREAM_HASH"), dt) +
              "  ".join("%s %.0f" % (n, ms) for n, (_, ms) in zip(names, res)))
    del res
