"""Proof wire formats (csrc/wire.cu), host only: POD image and plonky2-Write byte stream round-trip to the
same POD words; the serde-JSON form re-read with Python's json equals the POD field by field; malformed images are
rejected with SB_EINVAL.  The committed golden image (tests/golden/ecc_agg_proof.sbproof: a VALID ECCAggStark
3339 x 8192 proof at the reference's configuration, accepted by the oracle's verifier) is what
rust/starky_gpu/tests/verify.rs feeds to the reference's verify_stark_proof on a box with cargo."""
import json
import lzma
import os

import numpy as np
import pytest

import oracle_lib as O
import starky_bls12_381_b200 as sb
import toy_air
from helpers import to_oracle_params
from starky_bls12_381_b200 import airfiles
from starky_bls12_381_b200.binding import WireFormat, deserialize_words, serialize_words

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ecc_agg_proof.sbproof")


@pytest.fixture(scope="module")
def toy_proof(tmp_path_factory):
    air = toy_air.limbs(str(tmp_path_factory.mktemp("wire")), 4)
    trace, pis = air["witness"](10)                     # two FRI rounds: the steps are exercised
    p = sb.Params(205, 10, air["n_cols"], air["n_pis"], air["degree"], air["rate_bits"], 4, 2, 16, 84, 4, 5, 0, 0, 0)
    rc, words = O.prove(air["flat"], to_oracle_params(p), trace, pis)
    assert rc == 0, O.err()
    return p, words


def from_json(doc, layout):
    """The POD words back from the serde-JSON form (field order of sb_proof_layout)."""
    pr = doc["proof"]
    w = []
    hashes = lambda hs: [x for h in hs for x in h["elements"]]
    exts = lambda es: [x for e in es for x in e]
    w += hashes(pr["trace_cap"]) + hashes(pr["quotient_polys_cap"])
    assert pr["permutation_zs_cap"] is None
    op = pr["openings"]
    assert op["permutation_zs"] is None and op["permutation_zs_next"] is None
    w += exts(op["local_values"]) + exts(op["next_values"]) + exts(op["quotient_polys"])
    fri = pr["opening_proof"]
    for cap in fri["commit_phase_merkle_caps"]:
        w += hashes(cap)
    w += exts(fri["final_poly"]["coeffs"]) + [fri["pow_witness"]]
    for qr in fri["query_round_proofs"]:
        (tl, tp), (ql, qp) = qr["initial_trees_proof"]["evals_proofs"]
        w += tl + hashes(tp["siblings"]) + ql + hashes(qp["siblings"])
        for st in qr["steps"]:
            w += exts(st["evals"]) + hashes(st["merkle_proof"]["siblings"])
    w += doc["public_inputs"]
    return np.array(w, dtype=np.uint64)


def test_round_trips_and_json_shape(toy_proof):
    p, words = toy_proof
    pod = serialize_words(p, words, WireFormat.POD)
    assert pod[:8] == b"SBPROOF1" and len(pod) == 8 + 64 + 8 + 8 * words.size
    q, back = deserialize_words(pod, WireFormat.POD)
    assert bytes(q)[:48] == bytes(p)[:48] and np.array_equal(back, words)
    buf = serialize_words(p, words, WireFormat.PLONKY2_BUFFER)
    l = sb.ProofLayout()
    assert sb.lib().sb_proof_layout_for(p, l) == 0
    n_paths = l.n_queries * (2 + l.n_fri_rounds)
    assert len(buf) == 8 * words.size + n_paths          # one u8 length per Merkle proof, nothing else added
    _, back = deserialize_words(buf, WireFormat.PLONKY2_BUFFER, p)
    assert np.array_equal(back, words)
    # first bytes = the trace cap, little-endian
    assert np.array_equal(np.frombuffer(buf[:8 * 64], dtype="<u8"), words[:64])
    doc = json.loads(serialize_words(p, words, WireFormat.SERDE_JSON))
    assert doc["config"]["fri_config"] == {"rate_bits": p.rate_bits, "cap_height": 4, "proof_of_work_bits": 16,
                                           "reduction_strategy": {"ConstantArityBits": [4, 5]}, "num_query_rounds": 84}
    assert doc["degree_bits"] == 10 and len(doc["proof"]["opening_proof"]["query_round_proofs"]) == 84
    assert len(doc["proof"]["opening_proof"]["commit_phase_merkle_caps"]) == l.n_fri_rounds == 2
    assert np.array_equal(from_json(doc, l), words)


def test_malformed_images_are_rejected(toy_proof):
    p, words = toy_proof
    pod = bytearray(serialize_words(p, words, WireFormat.POD))
    buf = bytearray(serialize_words(p, words, WireFormat.PLONKY2_BUFFER))
    for bad in (bytes(pod[:-8]), bytes(pod) + b"\0" * 8, b"XBPROOF1" + bytes(pod[8:])):
        with pytest.raises(sb.SbError):
            deserialize_words(bad, WireFormat.POD)
    for bad in (bytes(buf[:-1]), bytes(buf) + b"\0"):
        with pytest.raises(sb.SbError):
            deserialize_words(bad, WireFormat.PLONKY2_BUFFER, p)
    noncanon = bytearray(buf)
    noncanon[0:8] = (0xFFFFFFFF00000001).to_bytes(8, "little")         # p itself: not a canonical element
    with pytest.raises(sb.SbError):
        deserialize_words(bytes(noncanon), WireFormat.PLONKY2_BUFFER, p)
    other = p.copy()
    other.n_cols += 1
    with pytest.raises(sb.SbError):
        deserialize_words(bytes(buf), WireFormat.PLONKY2_BUFFER, other)
    with pytest.raises(sb.SbError):
        deserialize_words(bytes(pod), WireFormat.POD, other)
    with pytest.raises(sb.SbError):
        serialize_words(other, words, WireFormat.POD)


def test_golden_ecc_agg_proof_is_a_valid_reference_sized_proof():
    """The image handed to the Rust test: parameters = StarkConfig::standard_fast_config() with rate_bits 2
    (aggregate_proof.rs:186-187), ECCAggStark 3339 x 8192; the oracle's verifier accepts it and rejects a tampered copy."""
    data = open(GOLDEN, "rb").read()
    q, words = deserialize_words(data, WireFormat.POD)
    info = sb.STARKS["ecc_agg"]
    want = sb.standard_params(info.stark_id, 13)
    assert bytes(q)[:48] == bytes(want)[:48] and q.flags == 0
    flat = airfiles.air_path("ecc_agg", "air")
    assert O.verify(flat, to_oracle_params(q), words) == 0, O.err()
    bad = words.copy()
    bad[70] ^= np.uint64(1)
    assert O.verify(flat, to_oracle_params(q), bad) != 0
