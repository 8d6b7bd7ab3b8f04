"""GPU: the reference's five starks at the sizes the reference instantiates them with, on VALID
traces from the witness generators: sb_prove through the C ABI WITHOUT SB_FLAG_ALLOW_INVALID_TRACE (the quotient must
divide), and the oracle's independent verifier (restatement of starky::verifier::verify_stark_proof) must accept the
GPU's proof and reject a tampered one.  All five -- including the full-size MillerLoop 97330 x 1024 (the sp leaf sponge)
and FinalExp 73527 x 8192 (the dp throughput leaf sponge) -- are additionally compared word for word with the oracle
prover's proof of the same trace (one CPU proof each: the two large ones are marked slow, ~20 s and a few minutes)."""
import numpy as np
import pytest

import oracle_lib as O
import starky_bls12_381_b200 as sb
from helpers import to_oracle_params
from starky_bls12_381_b200 import airfiles, witness as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def make(name, rng):
    fp, fp2 = (lambda: W.random_fp(rng)), (lambda: (W.random_fp(rng), W.random_fp(rng)))
    if name == "fp12_mul":
        return sb.FP12MulStark.new(16).generate_trace(W.random_fp12(rng), W.random_fp12(rng))
    if name == "pairing_precomp":
        return sb.PairingPrecompStark.new(1024).generate_trace(fp2(), fp2(), fp2())
    if name == "miller_loop":
        return sb.MillerLoopStark.new(1024).generate_trace(fp(), fp(), (fp2(), fp2(), fp2()))
    if name == "final_exp":
        return sb.FinalExponentiateStark.new(8192).generate_trace(W.random_fp12(rng))
    if name == "ecc_agg":
        return sb.ECCAggStark.new(8192).generate_trace([(fp(), fp()) for _ in range(512)],
                                                       [bool(b) for b in rng.integers(0, 2, 512)])
    raise KeyError(name)


@pytest.mark.parametrize("name,compare", [("fp12_mul", True), ("pairing_precomp", True), ("ecc_agg", True),
                                          pytest.param("miller_loop", True, marks=pytest.mark.slow),
                                          pytest.param("final_exp", True, marks=pytest.mark.slow)])
def test_valid_trace_proof_verifies(ctx, name, compare):
    info = sb.STARKS[name]
    rng = np.random.default_rng(0xB2002000 + info.stark_id)
    trace, pis = make(name, rng)
    assert trace.shape == (info.columns, info.num_rows) and pis.size == info.public_inputs
    flat = airfiles.air_path(name, "air")
    airfiles.air_path(name, "airbin")
    cfg = sb.StarkConfig.standard_fast_config()
    cfg.fri_config.rate_bits = info.rate_bits
    proof = sb.prove(ctx, name, cfg, trace, pis)                 # flags = 0: a non-divisible quotient is an error
    p = sb.api.params_for(name, cfg)
    op = to_oracle_params(p)
    assert O.verify(flat, op, proof.words) == 0, O.err()
    l = proof.layout
    for off in (int(l.off_trace_cap), int(l.off_local_values) + 7, int(l.off_quotient_polys), int(l.off_pow_witness),
                int(l.off_queries) + int(l.q_off_trace_leaf) + 5, int(l.off_public_inputs) + 1):
        bad = proof.words.copy()
        bad[off] = (int(bad[off]) + 1) % O.P
        assert O.verify(flat, op, bad) != 0, off
    if compare:
        O.lib().orc_set_num_threads(__import__("os").cpu_count() or 1)
        rc, want = O.prove(flat, op, trace, pis)
        assert rc == 0, O.err()
        assert np.array_equal(proof.words, want)


def test_invalid_real_trace_is_refused(ctx):
    """PairingPrecomp (degree 4 at rate_bits 2: quotient degree factor 3 < blow-up 4): one wrong limb makes the quotient
    non-divisible and sb_prove must refuse, like starky's trim_to_len panic."""
    rng = np.random.default_rng(0xB2002100)
    fp2 = lambda: (W.random_fp(rng), W.random_fp(rng))
    trace, pis = sb.PairingPrecompStark.new(1024).generate_trace(fp2(), fp2(), fp2())
    trace[W._PP.RX_OFFSET + 2, 30] ^= 1
    airfiles.air_path("pairing_precomp", "airbin")
    cfg = sb.StarkConfig.standard_fast_config()
    cfg.fri_config.rate_bits = 2
    with pytest.raises(sb.SbError) as e:
        sb.prove(ctx, "pairing_precomp", cfg, trace, pis)
    assert e.value.code == -4
