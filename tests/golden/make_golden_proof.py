"""Regenerates tests/golden/ecc_agg_proof.sbproof: one VALID ECCAggStark proof (3339 columns x 8192 rows,
StarkConfig::standard_fast_config() with rate_bits = 2 -- /root/reference/src/aggregate_proof.rs:186-188) as an
SB_WIRE_POD image (include/starky_b200.h).  The witness comes from the restated generate_trace
(starky_bls12_381_b200/witness, ecc_aggregate.rs:37-82) on seeded G1 points; the proof is produced by the CPU oracle,
whose proofs the GPU path equals word for word (tests/test_gpu_valid_proofs.py compares exactly this trace, and
tests/test_gpu_wire.py checks that sb_prove + sb_proof_serialize reproduce this file byte for byte).
rust/starky_gpu/tests/verify.rs feeds it to the reference's verify_stark_proof.

    python tests/golden/make_golden_proof.py
"""
import lzma
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
import starky_bls12_381_b200 as sb  # noqa: E402
from helpers import to_oracle_params  # noqa: E402
from starky_bls12_381_b200 import airfiles, witness as W  # noqa: E402
from starky_bls12_381_b200.binding import WireFormat, serialize_words  # noqa: E402

SEED = 0xB2002000 + 4          # the seed tests/test_gpu_valid_proofs.py uses for ecc_agg


def golden_inputs():
    rng = np.random.default_rng(SEED)
    fp = lambda: W.random_fp(rng)
    trace, pis = sb.ECCAggStark.new(8192).generate_trace([(fp(), fp()) for _ in range(512)],
                                                         [bool(b) for b in rng.integers(0, 2, 512)])
    cfg = sb.StarkConfig.standard_fast_config()
    cfg.fri_config.rate_bits = 2
    return sb.api.params_for("ecc_agg", cfg), trace, pis


def main():
    p, trace, pis = golden_inputs()
    flat = airfiles.air_path("ecc_agg", "air")
    rc, words = O.prove(flat, to_oracle_params(p), trace, pis)
    assert rc == 0, O.err()
    assert O.verify(flat, to_oracle_params(p), words) == 0, O.err()
    img = serialize_words(p, words, WireFormat.POD)
    out = os.path.join(HERE, "ecc_agg_proof.sbproof")
    with open(out, "wb") as f:
        f.write(img)        # uncompressed: the Rust test reads it with std::fs::read
    print("wrote %s: %d bytes (%d words)" % (out, len(img), words.size))


if __name__ == "__main__":
    main()
