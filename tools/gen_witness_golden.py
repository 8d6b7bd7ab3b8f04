#!/usr/bin/env python3
"""tests/golden/witness_final_exp.json: SHA-256 of the FinalExponentiateStark trace of a seeded input, computed from the
PYTHON restatement of the reference's generate_trace (starky_bls12_381_b200/witness, ~35 s), in the layout the C++
generator writes (row-major uint32 [8192][73527]).  tests/test_witness_cpp.py checks sb_witness_final_exp against it
without regenerating the 4.8 GB Python trace on every run (SB_FULL_WITNESS_COMPARE=1 runs the cell-for-cell comparison).

    python tools/gen_witness_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from starky_bls12_381_b200 import witness as W  # noqa: E402

SEED = 0xB2007600


def digest_rows_u32(trace_colmajor_u64):
    h = hashlib.sha256()
    rows = trace_colmajor_u64.shape[1]
    for r0 in range(0, rows, 256):                       # in row blocks: no second multi-GB copy
        h.update(np.ascontiguousarray(trace_colmajor_u64[:, r0:r0 + 256].T).astype(np.uint32).tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    x = W.random_fp12(np.random.default_rng(SEED))
    trace, pis = W.final_exp_trace(x)
    assert int(trace.max()) < (1 << 32)
    out = {"seed": SEED, "generator": "tools/gen_witness_golden.py (Python restatement of final_exponentiate.rs:137-281)",
           "rows": int(trace.shape[1]), "columns": int(trace.shape[0]),
           "sha256_rowmajor_u32": digest_rows_u32(trace), "sha256_public_inputs_u64": hashlib.sha256(pis.tobytes()).hexdigest()}
    path = os.path.join(ROOT, "tests", "golden", "witness_final_exp.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, out["sha256_rowmajor_u32"][:16])
