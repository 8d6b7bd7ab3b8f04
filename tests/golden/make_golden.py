"""Regenerates tests/golden/vectors.json.

Two kinds of vectors:
  * EXTERNAL anchors (not produced by this repository): plonky2's two published Poseidon-12 permutation KATs.
  * REGRESSION vectors produced by the CPU oracle (oracle/) on seeded inputs: sponge hashes of ragged lengths, a small
    trace commitment (coefficients / LDE / digests / cap), quotient values and a complete proof of a toy AIR with a
    valid witness.  They freeze today's restatement of SURVEY Appendix A so that the oracle and the CUDA path are both
    held to it; they are NOT outputs of the Rust reference (no cargo in the build image: "parity unpinned").

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
import toy_air  # noqa: E402

P = O.P


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def full_width(rng, shape):
    return (rng.integers(0, 1 << 63, shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, shape, dtype=np.uint64)) % np.uint64(P)


def make():
    rng = np.random.default_rng(0xB2000001)
    out = {"_comment": "see make_golden.py; external = plonky2 KATs, the rest = oracle regression vectors"}
    out["poseidon_external"] = [
        {"in": [0] * 12, "out": [int(x) for x in O.permute([0] * 12)]},
        {"in": list(range(12)), "out": [int(x) for x in O.permute(list(range(12)))]},
    ]
    states = full_width(rng, (6, 12))
    states[0] = P - 1
    out["poseidon"] = [{"in": [int(x) for x in s], "out": [int(x) for x in O.permute(s)]} for s in states]
    out["hash_or_noop"] = []
    for n in (1, 3, 4, 5, 8, 9, 16, 23, 135):
        x = full_width(rng, n)
        out["hash_or_noop"].append({"in": [int(v) for v in x], "out": [int(v) for v in O.hash_or_noop(x)]})
    # small trace commitment: 5 columns x 8 rows, rate_bits 1, and 11 columns x 64 rows, rate_bits 2
    out["lde_commit"] = []
    for log_n, n_cols, r, seed in ((3, 5, 1, 11), (6, 11, 2, 12)):
        trace = full_width(np.random.default_rng(seed), (n_cols, 1 << log_n))
        p = O.make_params(log_n=log_n, n_cols=n_cols, rate_bits=r)
        got = O.lde_commit(p, trace, want_coeffs=True)
        out["lde_commit"].append({"log_n": log_n, "n_cols": n_cols, "rate_bits": r, "seed": seed,
                                  "cap": [[int(v) for v in h] for h in got["cap"]],
                                  "sha_leaves": sha(got["leaves"]), "sha_digests": sha(got["digests"]),
                                  "sha_coeffs": sha(got["coeffs"])})
    # toy AIR with a valid witness: quotient values for fixed alphas and the whole proof
    d = tempfile.mkdtemp()
    air = toy_air.limbs(d, 4)
    log_n = 6
    trace, pis = air["witness"](log_n)
    p = O.make_params(stark_id=201, log_n=log_n, n_cols=air["n_cols"], n_pis=air["n_pis"], degree=air["degree"],
                      rate_bits=air["rate_bits"])
    alphas = np.array([0x1234567890ABCDEF % P, 0x0FEDCBA987654321 % P], np.uint64)
    q = O.quotient_values(air["flat"], p, trace, pis, alphas)
    rc, words = O.prove(air["flat"], p, trace, pis)
    assert rc == 0 and O.verify(air["flat"], p, words) == 0
    lay = O.layout(p)
    out["toy_limbs4"] = {"log_n": log_n, "sha_trace": sha(trace), "sha_quotient_values": sha(q),
                         "trace_cap0": [int(v) for v in words[lay.off_trace_cap:lay.off_trace_cap + 4]],
                         "pow_witness": int(words[lay.off_pow]), "total_words": int(words.size), "sha_proof": sha(words)}
    # a REAL stark at the size the reference instantiates it with: FP12MulStark 60285 x 16 on a VALID trace from the witness
    # generator (seeded Fp12 operands): trace, public inputs and the whole proof
    from starky_bls12_381_b200 import airfiles, witness as W
    wr = np.random.default_rng(0xB2005000)
    x, y = W.random_fp12(wr), W.random_fp12(wr)
    trace, pis = W.fp12_mul_trace(x, y)
    flat = airfiles.air_path("fp12_mul", "air")
    p = O.make_params(stark_id=0, log_n=4, n_cols=trace.shape[0], n_pis=pis.size, degree=3, rate_bits=1)
    rc, words = O.prove(flat, p, trace, pis)
    assert rc == 0 and O.verify(flat, p, words) == 0
    lay = O.layout(p)
    out["fp12_mul_valid"] = {"seed": 0xB2005000, "sha_trace": sha(trace), "sha_public_inputs": sha(pis),
                             "trace_cap0": [int(v) for v in words[lay.off_trace_cap:lay.off_trace_cap + 4]],
                             "pow_witness": int(words[lay.off_pow]), "total_words": int(words.size), "sha_proof": sha(words)}
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "vectors.json")
    json.dump(make(), open(path, "w"), indent=1)
    print("wrote", path)
