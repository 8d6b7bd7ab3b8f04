"""The constraint programs (starky_bls12_381_b200/air): counts and fingerprints against SURVEY.md Appendix B, and the
GPU bytecode's semantics (Python emulator of the instruction set) against the oracle's plain evaluation of the same
constraints in the reference's emission order -- at one random point, for all five starks."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from starky_bls12_381_b200 import airfiles

sys.path.insert(0, O.ROOT)
from tools.airgen.compile import emulate, read_airbin  # noqa: E402

P = O.P
MANIFEST = json.load(open(os.path.join(airfiles.AIR_DIR, "manifest.json")))
# SURVEY.md Appendix B
EXPECT = {
    "fp12_mul": (60285, 432, 82560, 26916, 55644, 0, 0, 8064, 377135969582744497, 5601190230),
    "pairing_precomp": (29376, 4968, 113634, 28944, 84474, 216, 0, 13176, 282723424429217662, 11576248902),
    "miller_loop": (97330, 5064, 145574, 48012, 97562, 0, 0, 14832, 337172101024110717, 17649424608),
    "final_exp": (73527, 288, 360800, 119598, 224818, 8192, 8192, 45256, 2027473019779257126, 110204239101),
    "ecc_agg": (3339, 12824, 20013, 1251, 18188, 574, 0, 14638, 1110302787620028249, 379379123),
}


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_manifest_matches_survey_appendix_b(name):
    m = MANIFEST[name]
    got = (m["n_cols"], m["n_public_inputs"], m["K"], m["plain"], m["transition"], m["first"], m["last"],
           m["uses_next"], m["cs_refs"], m["cs_class"])
    assert got == EXPECT[name]


def flat_class_counts(path):
    hdr = np.fromfile(path, dtype=np.uint32, count=10)
    n_consts, n_nodes, n_cons = int(hdr[5]), int(hdr[6]), int(hdr[7])
    off = 40 + 8 * n_consts + 12 * n_nodes
    cons = np.fromfile(path, dtype=np.uint32, offset=off, count=2 * n_cons).reshape(-1, 2)
    return cons[:, 0].astype(np.int64)


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_bytecode_equals_plain_horner_fold(name):
    flat = airfiles.air_path(name, "air")
    prog = read_airbin(airfiles.air_path(name, "airbin"))
    info = O.air_info(flat)
    C, NPI, K = info["n_cols"], info["n_pis"], info["n_constraints"]
    assert (C, NPI, K) == EXPECT[name][:3] == (prog.n_cols, prog.n_pis, prog.K)
    rng = np.random.default_rng(abs(hash(name)) % (1 << 32))
    rnd = lambda n: (rng.integers(0, 1 << 63, n, dtype=np.uint64) % np.uint64(P))
    local, nxt, pis = rnd(C), rnd(C), rnd(NPI)
    c = [int(x) for x in O.eval_constraints_row(flat, local, nxt, pis)]
    cls = flat_class_counts(flat)
    assert (np.bincount(cls, minlength=5)[1:] == np.array(EXPECT[name][3:7])).all()
    alphas = [int(x) for x in rnd(2)]
    factors = {1: 1, 2: int(rnd(1)[0]), 3: int(rnd(1)[0]), 4: int(rnd(1)[0])}
    want, weights = [], []
    for a in alphas:
        acc = 0
        for k in range(K):                       # ConstraintConsumer: acc = acc*alpha + f_k*c_k
            acc = (acc * a + factors[int(cls[k])] * c[k]) % P
        want.append(acc)
        w = [1] * K
        for k in range(K - 2, -1, -1):
            w[k] = w[k + 1] * a % P
        weights.append(w)
    values = [int(x) for x in local] + [int(x) for x in nxt] + [int(x) for x in pis]
    assert emulate(prog, values, factors, weights) == want
