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


# ---------------------------------------------------------------------------------------------------------------------
# The RUN form the library builds at load time (csrc/quotient.cu: translate_runs -- groups reordered by column locality,
# bodies sorted and folded into OP_RUN records, weight slots permuted), emulated here in plain Python integers against
# the same Horner fold.  This is the CPU check of the loader's rewrite; the CUDA evaluator of that form is checked
# against the oracle at every LDE point by tests/test_gpu_prove.py.
# ---------------------------------------------------------------------------------------------------------------------
def run_form(image):
    import ctypes as C
    import starky_bls12_381_b200 as sb
    L = sb.lib()
    L.sb_air_run_form.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    info = np.zeros(8, np.uint32)
    assert L.sb_air_run_form(image, len(image), None, 0, None, None, None, None, info.ctypes.data) == 0, L.sb_last_error(None)
    n_code2, n_slots, K, n_groups = (int(x) for x in info[:4])
    code2 = np.zeros(n_code2, np.uint64)
    slot_off2, slot_ks2 = np.zeros(n_slots + 1, np.uint32), np.zeros(max(K, 1), np.uint32)
    gpc2, gslot2 = np.zeros(n_groups + 1, np.uint32), np.zeros(n_groups + 1, np.uint32)
    assert L.sb_air_run_form(image, len(image), code2.ctypes.data, n_code2, slot_off2.ctypes.data, slot_ks2.ctypes.data,
                             gpc2.ctypes.data, gslot2.ctypes.data, info.ctypes.data) == 0
    return dict(code=[int(x) for x in code2], slot_off=slot_off2, slot_ks=slot_ks2[:K], gpc=gpc2, gslot=gslot2,
                n_runs=int(info[4]), n_run_bodies=int(info[5]), n_bodies=int(info[6]), n_cols=int(info[7]))


def emulate_run_form(rf, consts, n_cols, values, class_factors, weights):
    """Reference interpreter of the run form at ONE point.  values: indexed by variable (local, next, public inputs)."""
    nj = len(weights)
    so, sk = rf["slot_off"], rf["slot_ks"]
    slot_w = [[sum(w[int(k)] for k in sk[so[s]:so[s + 1]]) % P for s in range(len(so) - 1)] for w in weights]
    code, acc, G = rf["code"], [0] * nj, [0] * nj
    S, sel_left, have_group, slot, T, cls, pc = 1, 0, False, 0, 0, 1, 0
    ZERO = 0xFFFFF

    def val(v):
        return 0 if v == ZERO else values[v]

    def body(t):
        nonlocal slot
        for j in range(nj):
            G[j] = (G[j] + slot_w[j][slot] * (t % P)) % P
        slot += 1

    while pc < len(code):
        w = code[pc]
        op = w & 15
        if op == 12:                                        # OP_RUN: four words
            kind, signs, count, imm = (w >> 4) & 7, (w >> 8) & 0xFF, (w >> 16) & 0xFFF, w >> 32
            ops, dl = [], []
            for k in range(6):
                half = (code[pc + 1 + k // 2] >> (32 * (k % 2))) & 0xFFFFFFFF
                d = half >> 20
                ops.append(half & 0xFFFFF); dl.append(d - 4096 if d >= 2048 else d)
            sg = lambda t: -1 if (signs >> t) & 1 else 1
            c = -imm if signs & 0x80 else imm
            for i in range(count):
                v = [val(o + i * d) if o != ZERO else 0 for o, d in zip(ops, dl)]
                if kind == 0: t = v[0] - v[1]
                elif kind == 1: t = sg(0) * v[0] + c
                elif kind == 2: t = sg(0) * (v[0] << 32) + sum(sg(1 + q) * v[1 + q] for q in range(4)) + c
                elif kind == 3: t = sg(0) * v[0] * v[1] + sg(1) * (v[2] << 32) + sg(2) * v[3] + sg(3) * v[4] + c
                else: raise ValueError("bad run kind %d" % kind)
                assert sel_left == 0
                body(t)
            pc += 4
            continue
        pc += 1
        end, n0, n1 = (w >> 4) & 1, (w >> 5) & 1, (w >> 6) & 1
        v0, v1, v2 = (w >> 8) & 0x3FFFF, (w >> 26) & 0x3FFFF, (w >> 44) & 0x3FFFF
        imm = (w >> 26) & 0xFFFFFFFF
        s0 = -1 if n0 else 1
        if op == 11:
            if have_group:
                for j in range(nj):
                    acc[j] = (acc[j] + S * G[j]) % P
            have_group, G, cls, sel_left, S, T = True, [0] * nj, v0, v1, 1, 0
            if sel_left == 0:
                S = class_factors[cls]
            continue
        if op == 1: T += s0 * values[v0]
        elif op == 2: T += s0 * values[v0] + (-1 if n1 else 1) * values[v1]
        elif op == 3: T += s0 * values[v0] << 32
        elif op == 4: T += s0 * imm * values[v0]
        elif op == 5: T += s0 * values[v0] * values[v1]
        elif op == 6: T += s0 * consts[v2] * values[v0]
        elif op == 7: T += s0 * consts[v2] * values[v0] * values[v1]
        elif op == 8:
            T += s0 * code[pc] * values[v0] * values[v1] * values[v2]; pc += 1
        elif op == 9: T += s0 * imm
        elif op == 10: T += s0 * consts[v2]
        elif op != 0: raise ValueError("bad opcode %d" % op)
        if end:
            T %= P
            if sel_left:
                S = S * T % P
                sel_left -= 1
                if sel_left == 0:
                    S = S * class_factors[cls] % P
            else:
                body(T)
            T = 0
    if have_group:
        for j in range(nj):
            acc[j] = (acc[j] + S * G[j]) % P
    assert slot == len(so) - 1
    return acc


def _point(prog, seed):
    rng = np.random.default_rng(seed)
    rnd = lambda n: [int(x) for x in (rng.integers(0, 1 << 63, n, dtype=np.uint64) % np.uint64(P))]
    values = rnd(2 * prog.n_cols + prog.n_pis)
    alphas = rnd(2)
    factors = {1: 1, 2: rnd(1)[0], 3: rnd(1)[0], 4: rnd(1)[0]}
    weights = []
    for a in alphas:
        w = [1] * prog.K
        for k in range(prog.K - 2, -1, -1):
            w[k] = w[k + 1] * a % P
        weights.append(w)
    return values, factors, weights


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_run_form_of_the_loader_equals_the_word_form(name):
    path = airfiles.air_path(name, "airbin")
    prog = read_airbin(path)
    rf = run_form(open(path, "rb").read())
    assert rf["n_bodies"] == len(prog.slot_off) - 1 and rf["n_cols"] == prog.n_cols
    assert rf["n_run_bodies"] >= 0.99 * rf["n_bodies"]                 # the four loops cover (almost) every body
    assert rf["n_run_bodies"] / rf["n_runs"] >= 2.0                    # and the records are real runs
    assert sorted(int(k) for k in rf["slot_ks"]) == list(range(prog.K))
    values, factors, weights = _point(prog, 1 + sorted(EXPECT).index(name))
    want = emulate(prog, values, factors, weights)
    assert emulate_run_form(rf, prog.consts, prog.n_cols, values, factors, weights) == want


def test_run_form_of_toy_programs_equals_the_word_form(tmp_path):
    import toy_air
    for air in (toy_air.fibonacci(str(tmp_path)), toy_air.limbs(str(tmp_path), 4), toy_air.limbs(str(tmp_path), 3)):
        prog = read_airbin(air["airbin"])
        rf = run_form(open(air["airbin"], "rb").read())
        values, factors, weights = _point(prog, 99)
        assert emulate_run_form(rf, prog.consts, prog.n_cols, values, factors, weights) == emulate(prog, values, factors, weights)
