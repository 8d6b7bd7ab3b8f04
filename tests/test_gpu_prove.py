"""GPU parity through the C ABI: quotient values (K4) and full proofs (sb_prove) bit-exact against the CPU oracle, for
toy AIRs with VALID traces (the oracle verifier must accept the GPU's proof) and for the reference's five constraint
programs on seeded random traces (stage values and whole proofs, with SB_FLAG_ALLOW_INVALID_TRACE)."""
import numpy as np
import pytest

import oracle_lib as O
import starky_bls12_381_b200 as sb
import toy_air
from helpers import P, random_trace, to_oracle_params
from starky_bls12_381_b200 import airfiles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def airs(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("toyair"))
    return dict(fib=toy_air.fibonacci(d), limbs4=toy_air.limbs(d, 4), limbs3=toy_air.limbs(d, 3), limbs5=toy_air.limbs(d, 5))


def toy_params(air, log_n, stark_id, **kw):
    p = sb.Params(stark_id, log_n, air["n_cols"], air["n_pis"], air["degree"], air["rate_bits"], 4, 2,
                  kw.get("pow_bits", 16), kw.get("num_queries", 84), 4, 5, kw.get("flags", 0), 0, 0)
    return p


@pytest.mark.parametrize("name,log_n", [("fib", 5), ("fib", 10), ("limbs3", 6), ("limbs4", 6), ("limbs4", 10),
                                        ("limbs5", 7), ("limbs4", 13)])
def test_toy_quotient_and_proof_match_oracle(ctx, airs, name, log_n):
    air = airs[name]
    sid = 200 + sorted(airs).index(name)
    ctx.air_load(sid, air["airbin"])
    trace, pis = air["witness"](log_n)
    p = toy_params(air, log_n, sid)
    op = to_oracle_params(p)
    # stage parity: quotient values for injected alphas
    alphas = np.array([0x1234567890ABCDEF % P, 0x0FEDCBA987654321 % P], np.uint64)
    ctx.lde_commit(p, trace, want_lde=False, want_digests=False)
    got_q = ctx.quotient_values(p, pis, alphas)
    want_q = O.quotient_values(air["flat"], op, trace, pis, alphas)
    assert np.array_equal(got_q, want_q)
    # whole proof
    proof = ctx.prove(p, trace, pis)
    rc, want = O.prove(air["flat"], op, trace, pis)
    assert rc == 0, O.err()
    assert proof.layout.total_words == want.size
    if not np.array_equal(proof.words, want):
        l = proof.layout
        first = int(np.nonzero(proof.words != want)[0][0])
        raise AssertionError("proof differs from the oracle's at word %d (layout: %s)" % (
            first, {k: getattr(l, k) for k, _ in l._fields_}))
    assert O.verify(air["flat"], op, proof.words) == 0, O.err()


def test_invalid_trace_error_codes(ctx, airs):
    air = airs["limbs4"]
    ctx.air_load(210, air["airbin"])
    trace, pis = air["witness"](6)
    trace[5, 17] = (int(trace[5, 17]) + 1) % P
    with pytest.raises(sb.SbError) as e:
        ctx.prove(toy_params(air, 6, 210), trace, pis)
    assert e.value.code == -4 and "not divisible" in str(e.value)
    # benchmarking flag: truncate like the oracle does, byte-identical, and the verifier rejects it
    p = toy_params(air, 6, 210, flags=sb.Flags.ALLOW_INVALID_TRACE)
    proof = ctx.prove(p, trace, pis)
    rc, want = O.prove(air["flat"], to_oracle_params(p), trace, pis)
    assert rc == 0 and np.array_equal(proof.words, want)
    assert O.verify(air["flat"], to_oracle_params(p), proof.words) != 0
    # shape mismatch against the constraint program
    bad = toy_params(air, 6, 210)
    bad.n_cols = 15
    with pytest.raises(sb.SbError) as e:
        ctx.prove(bad, trace[:15], pis)
    assert e.value.code == -1


REAL = [("fp12_mul", 4), ("pairing_precomp", 4), ("miller_loop", 4), ("final_exp", 3), ("ecc_agg", 6)]


@pytest.mark.parametrize("name,log_n", REAL)
def test_reference_constraint_programs_quotient_parity(ctx, name, log_n):
    """The five starks' real constraint programs on seeded random u32 traces (SURVEY 8d distribution A) and one
    full-width trace (distribution B), at a small height: q_j(x) at every LDE point == oracle."""
    info = sb.STARKS[name]
    airfiles.air_path(name, "airbin")
    flat = airfiles.air_path(name, "air")
    p = sb.standard_params(info.stark_id, log_n)
    rng = np.random.default_rng(0xB2000000 + info.stark_id)
    for full in (False, True):
        trace = random_trace(rng, info.columns, log_n, full_width=full)
        pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
        alphas = rng.integers(0, 1 << 63, 2, dtype=np.uint64) % np.uint64(P)
        ctx.lde_commit(p, trace, want_lde=False, want_digests=False)
        got = ctx.quotient_values(p, pis, alphas)
        want = O.quotient_values(flat, to_oracle_params(p), trace, pis, alphas)
        assert np.array_equal(got, want), (name, full)


@pytest.mark.parametrize("name,log_n", [("pairing_precomp", 4), ("ecc_agg", 6)])
def test_word_form_interpreter_still_matches_oracle(ctx, monkeypatch, name, log_n):
    """SB_QUOTIENT_VM=1 selects round 1's word-form interpreter instead of the run-form evaluator: both must give the oracle's
    quotient values (the run form is what every other test in this file exercises)."""
    monkeypatch.setenv("SB_QUOTIENT_VM", "1")
    info = sb.STARKS[name]
    flat = airfiles.air_path(name, "air")
    p = sb.standard_params(info.stark_id, log_n)
    rng = np.random.default_rng(0xB2000100 + info.stark_id)
    trace = random_trace(rng, info.columns, log_n, full_width=True)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    alphas = rng.integers(0, 1 << 63, 2, dtype=np.uint64) % np.uint64(P)
    ctx.lde_commit(p, trace, want_lde=False, want_digests=False)
    got = ctx.quotient_values(p, pis, alphas)
    want = O.quotient_values(flat, to_oracle_params(p), trace, pis, alphas)
    assert np.array_equal(got, want)
    monkeypatch.delenv("SB_QUOTIENT_VM")
    assert np.array_equal(ctx.quotient_values(p, pis, alphas), want)


@pytest.mark.parametrize("name,log_n", [("fp12_mul", 4), ("miller_loop", 6), ("pairing_precomp", 5), ("ecc_agg", 7), ("final_exp", 5),
                                        ("final_exp", 8)])
def test_reference_starks_full_proof_parity_on_random_traces(ctx, name, log_n):
    info = sb.STARKS[name]
    flat = airfiles.air_path(name, "air")
    airfiles.air_path(name, "airbin")
    p = sb.standard_params(info.stark_id, log_n, flags=sb.Flags.ALLOW_INVALID_TRACE)
    rng = np.random.default_rng(0xB2001000 + info.stark_id)
    trace = random_trace(rng, info.columns, log_n)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    proof = ctx.prove(p, trace, pis)
    rc, want = O.prove(flat, to_oracle_params(p), trace, pis)
    assert rc == 0, O.err()
    assert np.array_equal(proof.words, want)
    assert proof.timings["ms_total"] > 0 and ctx.kernel_launches() > 0
