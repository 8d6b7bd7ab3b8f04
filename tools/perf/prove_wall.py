"""Where the wall time of one resident proof goes beyond the library's own event timeline (ms_total): the C call, the
Python wrapper, releasing the proof."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import starky_bls12_381_b200 as sb
from starky_bls12_381_b200 import binding as B

name = sys.argv[1] if len(sys.argv) > 1 else "miller_loop"
info = sb.STARKS[name]
p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1, flags=sb.Flags.ALLOW_INVALID_TRACE)
rng = np.random.default_rng(1)
trace = rng.integers(0, 1 << 32, (info.columns, info.num_rows), dtype=np.uint64)
pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
ctx = sb.Context(0)
ctx.trace_upload(p, trace)
L = B.lib()
for i in range(6):
    out = C.POINTER(B._CProof)()
    t0 = time.perf_counter()
    rc = L.sb_prove(ctx._h, C.byref(p), None, sb.TraceLayout.DEVICE_COLMAJOR_U64, B._ptr(pis), C.byref(out))
    t1 = time.perf_counter()
    assert rc == 0
    pr = B.Proof(out)
    t2 = time.perf_counter()
    ms_total = pr.timings["ms_total"]
    del pr
    t3 = time.perf_counter()
    print("%s: sb_prove %.2f ms (event timeline %.2f), wrap %.3f ms, release %.3f ms" % (name, 1e3 * (t1 - t0), ms_total, 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
ctx.close()
