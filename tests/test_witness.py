"""CPU-only: the witness generators (starky_bls12_381_b200/witness, the host mirror of the reference's generate_trace)
against (1) the reference's own known-answer vectors for the BLS12-381 arithmetic they are built on, and (2) the
constraint programs extracted from the reference by tools/airgen: EVERY constraint of a stark must vanish on EVERY
checked row of a generated trace (transition constraints on all rows but the last, first/last-row constraints on their
row).  The two sides are derived independently -- the constraints by symbolic execution of eval_packed_generic, the
traces by restating fill_trace_* -- so agreement pins both.  The oracle then proves and verifies the FP12Mul trace."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from starky_bls12_381_b200 import airfiles, witness as W
from starky_bls12_381_b200.witness import native as N

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls_kats.json")))


def constraint_classes(flat):
    hdr = np.fromfile(flat, dtype=np.uint32, count=10)
    off = 40 + 8 * int(hdr[5]) + 12 * int(hdr[6])
    return np.fromfile(flat, dtype=np.uint32, offset=off, count=2 * int(hdr[7])).reshape(-1, 2)[:, 0]


def violations(name, trace, pis, rows=None):
    """Number of (row, constraint) pairs that do not vanish."""
    flat = airfiles.air_path(name, "air")
    cls = constraint_classes(flat)
    n = trace.shape[1]
    tr = np.ascontiguousarray(trace.T)
    bad = 0
    for row in (range(n) if rows is None else rows):
        v = O.eval_constraints_row(flat, tr[row], tr[(row + 1) % n], pis)
        live = (cls == 1) | ((cls == 2) & (row != n - 1)) | ((cls == 3) & (row == 0)) | ((cls == 4) & (row == n - 1))
        bad += int(np.count_nonzero((v != 0) & live))
    return bad


def sample_rows(n, rng, k):
    """First/last rows, every 12-row block boundary region of a few blocks, and k random rows."""
    s = {0, 1, 10, 11, 12, 13, n - 2, n - 1}
    s.update(int(r) for r in rng.integers(0, n, k))
    return sorted(s)


# ---- the arithmetic under the generators, against the reference's own vectors ----
def test_bls_signature_pairing_product_is_one():
    g = GOLD["bls_signature"]
    pk = [int(v) for v in g["pk"]]
    hm = [tuple(int(v) for v in c) for c in g["hm"]]
    gen = [int(v) for v in g["g"]]
    sig = [tuple(int(v) for v in c) for c in g["sig"]]
    e1 = N.miller_loop(pk[0], N.P - pk[1], *hm)            # pairing(-pk, H(m))
    e2 = N.miller_loop(gen[0], gen[1], *sig)                # pairing(g, sig)
    assert N.fp12_final_exponentiate(N.fp12_mul(e1, e2)) == N.FP12_ONE


def test_final_exponentiate_kat():
    x = tuple(int(v) for v in GOLD["final_exponentiate_to_one"])
    assert N.fp12_final_exponentiate(x) == N.FP12_ONE


def test_g1_aggregate_kat():
    g = GOLD["g1_aggregate"]
    pts = [(int(x), int(y)) for x, y in g["points"]]

    def add(p, q):
        lam = N.fp_mul(N.fp_sub(q[1], p[1]), N.fp_inv(N.fp_sub(q[0], p[0])))
        x3 = N.fp_sub(N.fp_sub(N.fp_mul(lam, lam), q[0]), p[0])
        return x3, N.fp_sub(N.fp_mul(lam, N.fp_sub(p[0], x3)), p[1])
    acc = None
    for p, b in zip(pts, g["bits"]):
        if b:
            acc = p if acc is None else add(acc, p)
    assert acc == tuple(int(v) for v in g["res"])


def test_tower_identities():
    rng = np.random.default_rng(7)
    x, y = W.random_fp12(rng), W.random_fp12(rng)
    assert N.fp12_mul(x, N.fp12_inv(x)) == N.FP12_ONE
    assert N.fp12_frobenius(N.fp12_mul(x, y), 1) == N.fp12_mul(N.fp12_frobenius(x, 1), N.fp12_frobenius(y, 1))
    f = x
    for _ in range(12):
        f = N.fp12_frobenius(f, 1)
    assert f == x
    # cyclotomic square == square on the cyclotomic subgroup (after the easy part of the final exponentiation)
    c = N.fp12_mul(N.fp12_frobenius(x, 6), N.fp12_inv(x))
    c = N.fp12_mul(N.fp12_frobenius(c, 2), c)
    assert N.fp12_cyclotomic_square(c) == N.fp12_mul(c, c)


# ---- generated traces satisfy the extracted constraint programs ----
def test_fp12_mul_trace_satisfies_every_constraint_and_proves():
    rng = np.random.default_rng(1)
    x, y = W.random_fp12(rng), W.random_fp12(rng)
    trace, pis = W.fp12_mul_trace(x, y)
    assert trace.shape == (60285, 16) and pis.size == 432
    assert violations("fp12_mul", trace, pis) == 0
    flat = airfiles.air_path("fp12_mul", "air")
    p = O.make_params(stark_id=0, log_n=4, n_cols=60285, n_pis=432, degree=3, rate_bits=1)
    rc, words = O.prove(flat, p, trace, pis)
    assert rc == 0, O.err()
    assert O.verify(flat, p, words) == 0, O.err()
    # a wrong product is caught by the constraints and by the verifier (degree 3 at rate_bits 1: the quotient domain
    # equals the LDE domain, so the prover itself cannot see the non-divisibility -- same as starky)
    bad = trace.copy()
    bad[W._FP12MUL.TOTAL_COLUMNS - 40, 3] ^= 1
    assert violations("fp12_mul", bad, pis) > 0
    rc, words_bad = O.prove(flat, p, bad, pis)
    assert rc != 0 or O.verify(flat, p, words_bad) != 0
    # a wrong claimed output (public input) is caught too
    pis2 = pis.copy()
    pis2[-1] ^= 1
    assert violations("fp12_mul", trace, pis2) > 0


def test_fp12_mul_trace_with_zero_and_one_operands():
    """Edge values: x = 1, y with zero coefficients (negation of 0 is p in the reference, native.rs:417-424)."""
    y = tuple([5, 0, 0, N.P - 1, 0, 1, 0, 0, 2, 0, 0, 0])
    trace, pis = W.fp12_mul_trace(N.FP12_ONE, y)
    assert violations("fp12_mul", trace, pis) == 0
    assert [int(v) for v in pis[288:300]] == N.limbs(5)


def test_pairing_precomp_trace_satisfies_every_constraint():
    rng = np.random.default_rng(2)
    q = [(W.random_fp(rng), W.random_fp(rng)) for _ in range(3)]
    trace, pis = W.pairing_precomp_trace(*q)
    assert trace.shape == (29376, 1024) and pis.size == 4968
    assert violations("pairing_precomp", trace, pis, sample_rows(1024, rng, 120)) == 0
    bad = trace.copy()
    bad[W._PP.RX_OFFSET + 2, 30] ^= 1
    assert violations("pairing_precomp", bad, pis, [29, 30]) > 0


def test_miller_loop_trace_satisfies_every_constraint():
    rng = np.random.default_rng(3)
    q = [(W.random_fp(rng), W.random_fp(rng)) for _ in range(3)]
    trace, pis = W.miller_loop_trace(W.random_fp(rng), W.random_fp(rng), q)
    assert trace.shape == (97330, 1024) and pis.size == 5064
    assert violations("miller_loop", trace, pis, sample_rows(1024, rng, 60) + [815, 816, 817]) == 0


def test_ecc_aggregate_trace_satisfies_every_constraint():
    rng = np.random.default_rng(4)
    pts = [(W.random_fp(rng), W.random_fp(rng)) for _ in range(512)]
    bits = [bool(b) for b in rng.integers(0, 2, 512)]
    trace, pis, res = W.ecc_aggregate_trace(pts, bits)
    assert trace.shape == (3339, 8192) and pis.size == 12824
    assert violations("ecc_agg", trace, pis, sample_rows(8192, rng, 300) + [6131, 6132, 6143, 6144]) == 0
    pis[-1] ^= 1          # wrong aggregate
    assert violations("ecc_agg", trace, pis, [6130, 6131, 6132, 6143]) > 0


@pytest.mark.slow
def test_final_exp_trace_satisfies_every_constraint():
    rng = np.random.default_rng(5)
    trace, pis = W.final_exp_trace(W.random_fp12(rng))
    assert trace.shape == (73527, 8192) and pis.size == 288
    E = W._FE
    edges = [getattr(E, "T%d_ROW" % i) for i in range(32)] + [E.TOTAL_ROW]
    rows = sorted(set(sample_rows(8192, rng, 40) + edges + [r - 1 for r in edges if r] + [r + 1 for r in edges]))
    assert violations("final_exp", trace, pis, rows) == 0
