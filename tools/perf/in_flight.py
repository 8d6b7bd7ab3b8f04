"""Throughput of k proofs in flight on one GPU (k contexts, k host threads): ms per proof for a stark shape."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import starky_bls12_381_b200 as sb
from starky_bls12_381_b200 import airfiles

name = sys.argv[1] if len(sys.argv) > 1 else "pairing_precomp"
info = sb.STARKS[name]
airfiles.air_path(name, "airbin")
p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
rng = np.random.default_rng(1)
trace = rng.integers(0, 1 << 32, (info.columns, info.num_rows), dtype=np.uint64)
pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
host = torch.from_numpy(trace.view(np.int64)).pin_memory()
for k in (1, 2, 3, 4, 6):
    ctxs = [sb.Context(0) for _ in range(k)]
    for c in ctxs:
        c.trace_upload(p, host.data_ptr())
        c.prove(p, None, pis, sb.TraceLayout.DEVICE_COLMAJOR_U64)
    reps = 4

    def worker(c):
        for _ in range(reps):
            c.prove(p, None, pis, sb.TraceLayout.DEVICE_COLMAJOR_U64)
    for attempt in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ts = [threading.Thread(target=worker, args=(c,)) for c in ctxs]
        for t in ts: t.start()
        for t in ts: t.join()
        dt = time.perf_counter() - t0
    print(name, "in flight", k, "ms per proof %.2f" % (1e3 * dt / (k * reps)), flush=True)
    for c in ctxs:
        c.close()
