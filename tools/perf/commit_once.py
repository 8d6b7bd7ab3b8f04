"""One trace commitment of a named stark on synthetic data (profiling target for ncu)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import starky_bls12_381_b200 as sb
name = sys.argv[1] if len(sys.argv) > 1 else "pairing_precomp"
info = sb.STARKS[name]
p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1)
trace = np.random.default_rng(1).integers(0, 1 << 32, (info.columns, info.num_rows), dtype=np.uint64)
ctx = sb.Context(0)
ctx.trace_upload(p, trace)
for _ in range(2):
    ctx.lde_commit(p, None, sb.TraceLayout.DEVICE_COLMAJOR_U64, want_lde=False, want_digests=False)
print({k: round(ctx.stage_ms(k), 3) for k in ("lde", "leaf_hash", "merkle")})
ctx.close()
