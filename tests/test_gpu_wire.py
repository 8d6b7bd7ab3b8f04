"""GPU: sb_prove of the golden inputs + sb_proof_serialize reproduce the committed image byte for byte -- the file the
Rust test (rust/starky_gpu/tests/verify.rs) hands to the reference's verify_stark_proof IS a GPU proof."""
import os
import sys

import numpy as np
import pytest

import starky_bls12_381_b200 as sb
from starky_bls12_381_b200.binding import WireFormat, deserialize_words, serialize_words

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_gpu_proof_serializes_to_the_committed_golden_image():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_proof import golden_inputs
    p, trace, pis = golden_inputs()
    ctx = sb.Context(0)
    try:
        proof = ctx.prove(p, trace, pis)
    finally:
        ctx.close()
    want = open(os.path.join(HERE, "golden", "ecc_agg_proof.sbproof"), "rb").read()
    got = serialize_words(p, proof.words, WireFormat.POD)
    assert got == want
    _, back = deserialize_words(serialize_words(p, proof.words, WireFormat.PLONKY2_BUFFER), WireFormat.PLONKY2_BUFFER, p)
    assert np.array_equal(back, proof.words)
