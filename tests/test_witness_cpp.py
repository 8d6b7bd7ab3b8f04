"""The C++ witness generator (csrc/witness.cpp) against the Python restatement of the reference's
generate_trace (starky_bls12_381_b200/witness): cell for cell and public input for public input, and every one of the
82 560 extracted constraints vanishes on every row of the C++ trace (CPU).  GPU: sb_prove_fp12_mul (operands in, proof
out) equals the proof of the Python trace and the oracle's verifier accepts it."""
import numpy as np
import pytest

import oracle_lib as O
import starky_bls12_381_b200 as sb
from helpers import to_oracle_params
from starky_bls12_381_b200 import airfiles, witness as W
from starky_bls12_381_b200.binding import (witness_ecc_agg, witness_final_exp, witness_fp12_mul, witness_miller_loop,
                                           witness_pairing_precomp)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_cpp_fp12_mul_trace_equals_the_python_restatement(seed):
    rng = np.random.default_rng(0xB2007000 + seed)
    x, y = W.random_fp12(rng), W.random_fp12(rng)
    if seed == 3:                      # edge values: zero and p - 1 coefficients
        x = (0, W.N.P - 1) + tuple(x[2:])
        y = (W.N.P - 1,) * 12
    want_t, want_p = W.fp12_mul_trace(x, y, 16)            # column-major uint64 [60285][16]
    got_t, got_p = witness_fp12_mul(x, y, 16)              # row-major uint32 [16][60285]
    assert np.array_equal(got_p, want_p)
    assert np.array_equal(got_t.astype(np.uint64).T, want_t)


def test_every_constraint_vanishes_on_the_cpp_trace():
    rng = np.random.default_rng(0xB2007100)
    trace, pis = witness_fp12_mul(W.random_fp12(rng), W.random_fp12(rng), 16)
    flat = airfiles.air_path("fp12_mul", "air")
    rows = trace.astype(np.uint64)
    for r in range(16):
        c = O.eval_constraints_row(flat, rows[r], rows[(r + 1) % 16], pis)
        # transition constraints are not enforced on the last row, first / last-row classes only there: check what vanishes
        # everywhere -- the plain ones -- on all rows and all classes on the rows where they apply via the oracle prover below
        if r < 15:
            assert not c.any(), (r, int(np.count_nonzero(c)))
    p = sb.standard_params(sb.StarkId.FP12_MUL, 4)
    rc, words = O.prove(flat, to_oracle_params(p), np.ascontiguousarray(rows.T), pis)     # flags = 0: the quotient must divide
    assert rc == 0, O.err()
    assert O.verify(flat, to_oracle_params(p), words) == 0, O.err()


def test_cpp_ecc_agg_trace_equals_the_python_restatement():
    rng = np.random.default_rng(0xB2007300)
    pts = [(W.random_fp(rng), W.random_fp(rng)) for _ in range(512)]
    bits = [bool(b) for b in rng.integers(0, 2, 512)]
    bits[0], bits[1], bits[2] = False, True, False            # first operand "at infinity", a skipped point
    want_t, want_p, want_res = W.ecc_aggregate_trace(pts, bits)          # column-major uint64 [3339][8192]
    got_t, got_p, got_res = witness_ecc_agg(pts, bits)                   # row-major uint32 [8192][3339]
    assert got_res == want_res
    assert np.array_equal(got_p, want_p)
    assert np.array_equal(got_t.astype(np.uint64).T, want_t)
    with pytest.raises(sb.SbError):
        witness_ecc_agg(pts, bits, 4096)                      # 511 additions of 12 rows do not fit


def _rand_fp2(rng):
    return (W.random_fp(rng), W.random_fp(rng))


@pytest.mark.parametrize("seed", [1, 2])
def test_cpp_pairing_precomp_trace_equals_the_python_restatement(seed):
    rng = np.random.default_rng(0xB2007400 + seed)
    x, y, z = _rand_fp2(rng), _rand_fp2(rng), _rand_fp2(rng)
    if seed == 2:
        z = (1, 0)                                            # an affine point handed over as projective
    want_t, want_p = W.pairing_precomp_trace(x, y, z, 1024)  # column-major uint64 [29376][1024]
    got_t, got_p = witness_pairing_precomp(x, y, z, 1024)    # row-major uint32 [1024][29376]
    assert np.array_equal(got_p, want_p)
    assert np.array_equal(got_t.astype(np.uint64).T, want_t)


def test_cpp_miller_loop_trace_equals_the_python_restatement():
    rng = np.random.default_rng(0xB2007500)
    px, py = W.random_fp(rng), W.random_fp(rng)
    q = (_rand_fp2(rng), _rand_fp2(rng), _rand_fp2(rng))
    want_t, want_p = W.miller_loop_trace(px, py, q, 1024)    # column-major uint64 [97330][1024]
    got_t, got_p = witness_miller_loop(px, py, q, 1024)      # row-major uint32 [1024][97330]
    assert np.array_equal(got_p, want_p)
    assert np.array_equal(got_t.astype(np.uint64).T, want_t)
    with pytest.raises(sb.SbError):
        witness_miller_loop(W.N.P, py, q, 1024)               # unreduced coordinate


def test_cpp_final_exp_trace_equals_the_python_restatement():
    """The 2.4 GB FinalExponentiateStark trace against the Python restatement's: by the committed SHA-256 of the Python
    trace of the same seeded input (tests/golden/witness_final_exp.json, tools/gen_witness_golden.py -- regenerating the
    Python trace takes 35 s), public inputs against the native tower; SB_FULL_WITNESS_COMPARE=1: cell for cell."""
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "witness_final_exp.json")))
    x = W.random_fp12(np.random.default_rng(g["seed"]))
    got_t, got_p = witness_final_exp(x)                       # row-major uint32 [8192][73527] (2.4 GB)
    assert got_t.shape == (g["rows"], g["columns"]) and got_t.dtype == np.uint32
    h = hashlib.sha256()
    for r0 in range(0, got_t.shape[0], 256):
        h.update(got_t[r0:r0 + 256].tobytes())
    assert h.hexdigest() == g["sha256_rowmajor_u32"]
    assert hashlib.sha256(got_p.tobytes()).hexdigest() == g["sha256_public_inputs_u64"]
    out = W.N.fp12_final_exponentiate(x)
    assert [int(v) for v in got_p[144:]] == [l for c in out for l in W.N.limbs(c)]
    if os.environ.get("SB_FULL_WITNESS_COMPARE"):
        want_t, want_p = W.final_exp_trace(x)                 # column-major uint64 [73527][8192]
        assert np.array_equal(got_p, want_p)
        for c0 in range(0, want_t.shape[0], 4096):            # compare in column blocks: no second 4.8 GB copy
            assert np.array_equal(got_t[:, c0:c0 + 4096].astype(np.uint64).T, want_t[c0:c0 + 4096]), c0
    with pytest.raises(sb.SbError):
        witness_final_exp(x, 4096)                            # one row-selector column per row: 8192 rows only


def test_api_generate_trace_rows_is_the_row_major_form_of_generate_trace():
    """api.XStark.generate_trace_rows (C++ generators, the reference's Vec<[F; COLUMNS]> as uint32 rows) against
    generate_trace (Python restatement, trace_rows_to_poly_values order) on the two small starks."""
    rng = np.random.default_rng(0xB2007900)
    st = sb.FP12MulStark.new(16)
    x, y = W.random_fp12(rng), W.random_fp12(rng)
    (cols, pis), (rows, pis2) = st.generate_trace(x, y), st.generate_trace_rows(x, y)
    assert rows.dtype == np.uint32 and rows.shape == (16, st.info.columns)
    assert np.array_equal(rows.astype(np.uint64).T, cols) and np.array_equal(pis, pis2)
    st = sb.PairingPrecompStark.new(1024)
    q = [_rand_fp2(rng) for _ in range(3)]
    (cols, pis), (rows, pis2) = st.generate_trace(*q), st.generate_trace_rows(*q)
    assert rows.shape == (1024, st.info.columns)
    assert np.array_equal(rows.astype(np.uint64).T, cols) and np.array_equal(pis, pis2)


def test_unreduced_operands_are_rejected():
    bad = (W.N.P,) + (0,) * 11
    with pytest.raises(sb.SbError):
        witness_fp12_mul(bad, bad, 16)
    with pytest.raises(sb.SbError):
        witness_fp12_mul((1,) * 12, (1,) * 12, 12)          # not a power of two


@pytest.mark.gpu
def test_gpu_proof_from_operands_equals_the_proof_of_the_python_trace():
    rng = np.random.default_rng(0xB2007200)
    x, y = W.random_fp12(rng), W.random_fp12(rng)
    trace, pis = W.fp12_mul_trace(x, y, 16)
    p = sb.standard_params(sb.StarkId.FP12_MUL, 4)
    ctx = sb.Context(0)
    try:
        want = ctx.prove(p, trace, pis)
        got = ctx.prove_fp12_mul(p, x, y)
    finally:
        ctx.close()
    assert np.array_equal(got.words, want.words)
    assert O.verify(airfiles.air_path("fp12_mul", "air"), to_oracle_params(p), got.words) == 0, O.err()


@pytest.mark.gpu
def test_gpu_proofs_from_operands_of_the_larger_starks():
    """sb_prove_pairing_precomp / sb_prove_miller_loop (operands in, proof out; the trace is generated in C++ as row-major
    u32) give the proof of the Python restatement's trace word for word, and the oracle's verifier accepts it."""
    rng = np.random.default_rng(0xB2007700)
    x, y, z = _rand_fp2(rng), _rand_fp2(rng), _rand_fp2(rng)
    px, py = W.random_fp(rng), W.random_fp(rng)
    ctx = sb.Context(0)
    try:
        p = sb.standard_params(sb.StarkId.PAIRING_PRECOMP, 10)
        trace, pis = W.pairing_precomp_trace(x, y, z, 1024)
        want, got = ctx.prove(p, trace, pis), ctx.prove_pairing_precomp(p, x, y, z)
        assert np.array_equal(got.words, want.words)
        assert O.verify(airfiles.air_path("pairing_precomp", "air"), to_oracle_params(p), got.words) == 0, O.err()
        p = sb.standard_params(sb.StarkId.MILLER_LOOP, 10)
        trace, pis = W.miller_loop_trace(px, py, (x, y, z), 1024)
        want, got = ctx.prove(p, trace, pis), ctx.prove_miller_loop(p, px, py, (x, y, z))
        assert np.array_equal(got.words, want.words)
        assert O.verify(airfiles.air_path("miller_loop", "air"), to_oracle_params(p), got.words) == 0, O.err()
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_final_exp_proof_from_its_operand_verifies():
    rng = np.random.default_rng(0xB2007800)
    x = W.random_fp12(rng)
    p = sb.standard_params(sb.StarkId.FINAL_EXP, 13)
    ctx = sb.Context(0)
    try:
        got = ctx.prove_final_exp(p, x)
    finally:
        ctx.close()
    assert O.verify(airfiles.air_path("final_exp", "air"), to_oracle_params(p), got.words) == 0, O.err()
    want_out = W.N.fp12_final_exponentiate(x)
    pis = got.field("off_public_inputs", 288)
    assert [int(v) for v in pis[144:288]] == [l for c in want_out for l in W.N.limbs(c)]
