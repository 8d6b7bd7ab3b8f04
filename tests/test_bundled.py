"""The reference's bundled inputs (light_client_update_period_1052 / 1053, main.rs:8-55) through this repository's own
input preparation (starky_bls12_381_b200/bls.py: decompression, aggregation, SSZ signing root, hash to G2) and witness
generators.  CPU: the derived values equal the committed fixture, the pairing equation holds (the statement the seven
proofs establish), and the FP12Mul job of the set proves and verifies in the oracle.  GPU (slow): all seven proofs of the
set through sb_prove_batch on VALID traces, every one accepted by the oracle's verifier, FinalExp's output is ONE."""
import numpy as np
import pytest

import oracle_lib as O
import starky_bls12_381_b200 as sb
from helpers import to_oracle_params
from starky_bls12_381_b200 import airfiles, bls, bundled
from starky_bls12_381_b200.witness import native as N


@pytest.fixture(scope="module")
def inputs():
    return bundled.load_inputs()


def test_bundled_inputs_satisfy_the_pairing_equation(inputs):
    assert sum(inputs["bits"]) == 509 and len(inputs["points"]) == 512
    for pt in inputs["points"][:8] + [inputs["apk"]]:
        assert (pt[1] * pt[1] - pt[0] ** 3 - 4) % N.P == 0                       # on G1
    for q in (inputs["q1"], inputs["q2"]):
        lhs, rhs = bls.f2_sqr(q[1]), bls.f2_add(bls.f2_mul(bls.f2_sqr(q[0]), q[0]), (4, 4))
        assert lhs == rhs                                                         # on G2
    assert bls.pairing_product_is_one(inputs)
    # a different message breaks it
    other = dict(inputs, q1=bls.hash_to_curve_g2(b"\x01" * 32))
    assert not bls.pairing_product_is_one(other)


def test_hash_to_curve_building_blocks():
    # expand_message_xmd: RFC 9380 appendix K.1 (SHA-256, DST "QUUX-V01-CS02-with-expander-SHA256-128"), msg "" and "abc", 0x20 bytes
    dst = b"QUUX-V01-CS02-with-expander-SHA256-128"
    assert bls.expand_message_xmd(b"", dst, 32).hex() == "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"
    assert bls.expand_message_xmd(b"abc", dst, 32).hex() == "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"
    # the psi constants equal the reference's (hash_to_curve.rs:257-262, :280)
    assert bls.PSI_CX == (0, 4002409555221667392624310435006688643935503118305586438271171395842971157480381377015405980053539358417135540939437)
    assert bls.PSI_CY == (2973677408986561043442465346520108879172042883009249989176415018091420807192182638567116318576472649347015917690530,
                          1028732146235106349975324479215795277384839936929757896155643118032610843298655225875571310552543014690878354869257)
    assert bls.PSI2_CX == 4002409555221667392624310435006688643935503118305586438271171395842971157480381377015405980053539358417135540939436
    # the hashed point lies in the order-r subgroup
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    assert bls.G2.mul(bls.hash_to_curve_g2(b"starky"), R) is None


def test_bundled_fp12_mul_job_proves_and_verifies_in_the_oracle(inputs):
    (name, trace, pis), = bundled.jobs(inputs, {"fp12_mul"})
    p = sb.standard_params(sb.StarkId.FP12_MUL, 4)
    flat = airfiles.air_path("fp12_mul", "air")
    rc, words = O.prove(flat, to_oracle_params(p), trace, pis)
    assert rc == 0, O.err()
    assert O.verify(flat, to_oracle_params(p), words) == 0, O.err()


def test_bundled_native_jobs_equal_the_python_jobs_on_the_small_starks(inputs):
    """bundled.jobs_native (C++ witness generators, row-major u32) against bundled.jobs (Python restatement, column-major u64)
    on the reference's own inputs: FP12Mul and both PairingPrecomp jobs (the others are compared in test_witness_cpp.py)."""
    import numpy as np
    want = bundled.jobs(inputs, {"fp12_mul", "pairing_precomp"})
    got = bundled.jobs_native(inputs, {"fp12_mul", "pairing_precomp"})
    assert [j[0] for j in got] == [j[0] for j in want] == ["pairing_precomp", "pairing_precomp", "fp12_mul"]
    for (_, wt, wp), (_, gt, gp) in zip(want, got):
        assert np.array_equal(gp, wp)
        assert np.array_equal(gt.astype(np.uint64).T, wt)


@pytest.mark.gpu
@pytest.mark.slow
def test_gpu_proves_the_seven_bundled_proofs(inputs):
    from starky_bls12_381_b200.binding import prove_batch
    js = bundled.jobs(inputs)
    assert [j[0] for j in js] == bundled.ORDER
    batch = []
    for name, trace, pis in js:
        info = sb.STARKS[name]
        p = sb.standard_params(info.stark_id, trace.shape[1].bit_length() - 1)            # flags = 0: the quotient must divide
        batch.append((p, trace, sb.TraceLayout.COLMAJOR_U64, pis))
    ctxs = [sb.Context(0) for _ in range(3)]
    try:
        res = prove_batch(ctxs, batch)
    finally:
        for c in ctxs:
            c.close()
    for (name, trace, pis), (p, _, _, _), (proof, ms) in zip(js, batch, res):
        assert not isinstance(proof, Exception), (name, proof)
        assert O.verify(airfiles.air_path(name, "air"), to_oracle_params(p), proof.words) == 0, (name, O.err())
    fe_pis = js[-1][2]
    assert [int(v) for v in fe_pis[-144:]] == [1] + [0] * 143          # the pairing product final-exponentiates to ONE
