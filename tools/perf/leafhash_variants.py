"""Times the trace commitment stages (lde / leaf_hash / merkle) per stark shape for each leaf-hash variant."""
import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import starky_bls12_381_b200 as sb

def run(name, lanes):
    os.environ["SB_LEAF_LANES"] = str(lanes) if lanes else ""
    if not lanes: os.environ.pop("SB_LEAF_LANES")
    info = sb.STARKS[name]
    p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1)
    rng = np.random.default_rng(1)
    trace = rng.integers(0, 1 << 32, (info.columns, info.num_rows), dtype=np.uint64)
    ctx = sb.Context(0)
    ctx.trace_upload(p, trace)
    del trace
    best = None
    for _ in range(2):
        ctx.lde_commit(p, None, sb.TraceLayout.DEVICE_COLMAJOR_U64, want_lde=False, want_digests=False)
        t = {k: ctx.stage_ms(k) for k in ("lde", "leaf_hash", "merkle")}
        best = t if best is None or t["leaf_hash"] < best["leaf_hash"] else best
    ctx.close()
    return best

if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["pairing_precomp", "miller_loop", "ecc_agg", "fp12_mul", "final_exp"]
    for name in names:
        for lanes in (1, 4, 12):
            if name == "final_exp" and lanes == 12: continue
            print(name, "lanes", lanes, {k: round(v, 2) for k, v in run(name, lanes).items()}, flush=True)
