"""The reference's bundled run (main.rs:8-55 -> aggregate_proof.rs:228-370) as seven prover jobs: the inputs of
tests/golden/bundled_inputs.json (sync committee 1052's public keys, update 1053's sync aggregate and attested header)
go through bls.prepare and the witness generators exactly as generate_aggregate_proof feeds its *_main functions:

    ecc_aggregate     (512 points, participation bits)                  aggregate_proof.rs:186-227, :276
    pairing_precomp 1 (Q1 = hash_to_curve_g2(signing_root))             :24-69, :304
    miller_loop 1     (apk, Q1)                                          :71-121, :311
    pairing_precomp 2 (Q2 = signature)                                   :337
    miller_loop 2     (-g1, Q2)                                          :345
    fp12_mul          (ml1, ml2)                                         :122-151, :355
    final_exp         (ml1 * ml2)   -- result must be ONE                :153-184, :364
"""
import json
import os

from . import bls
from .api import STARKS
from .witness import native as N

FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bundled_inputs.json")
ORDER = ["ecc_agg", "pairing_precomp", "miller_loop", "pairing_precomp", "miller_loop", "fp12_mul", "final_exp"]


def load_inputs(path=FIXTURE, check=True):
    d = json.load(open(path))
    root = bls.signing_root(d["attested_header"], bytes.fromhex(d["domain"][2:]))
    inp = bls.prepare(d["pubkeys"], d["sync_committee_bits"], d["sync_committee_signature"], root)
    if check:
        der = d["derived"]
        assert "0x" + root.hex() == der["signing_root"], "signing root differs from the fixture"
        assert [str(v) for v in inp["apk"]] == der["apk"], "aggregated public key differs from the fixture"
        assert [[str(v) for v in c] for c in inp["q1"]] == der["q1"], "hashed message differs from the fixture"
        assert [[str(v) for v in c] for c in inp["q2"]] == der["q2"], "signature point differs from the fixture"
    return inp


def jobs(inp, only=None):
    """-> [(stark name, column-major uint64 trace, public inputs)] in the reference's proving order; `only`: a set of
    names to build (the large traces take a while in Python)."""
    from . import witness as W
    one = (1, 0)
    q1, q2 = inp["q1"], inp["q2"]
    ml1 = N.miller_loop(inp["apk"][0], inp["apk"][1], q1[0], q1[1], one)
    ml2 = N.miller_loop(bls.NEG_G1[0], bls.NEG_G1[1], q2[0], q2[1], one)
    makers = [
        ("ecc_agg", lambda: W.ecc_aggregate_trace(inp["points"], inp["bits"], STARKS["ecc_agg"].num_rows)[:2]),
        ("pairing_precomp", lambda: W.pairing_precomp_trace(q1[0], q1[1], one, 1024)),
        ("miller_loop", lambda: W.miller_loop_trace(inp["apk"][0], inp["apk"][1], (q1[0], q1[1], one), 1024)),
        ("pairing_precomp", lambda: W.pairing_precomp_trace(q2[0], q2[1], one, 1024)),
        ("miller_loop", lambda: W.miller_loop_trace(bls.NEG_G1[0], bls.NEG_G1[1], (q2[0], q2[1], one), 1024)),
        ("fp12_mul", lambda: W.fp12_mul_trace(ml1, ml2, 16)),
        ("final_exp", lambda: W.final_exp_trace(N.fp12_mul(ml1, ml2), 8192)),
    ]
    out = []
    for name, make in makers:
        if only is None or name in only:
            trace, pis = make()
            out.append((name, trace, pis))
    return out
