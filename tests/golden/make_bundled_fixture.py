"""Regenerates (needs /root/reference):
  * starky_bls12_381_b200/witness/iso_g2.json -- the 3-isogeny coefficient table of the reference
    (/root/reference/src/hash_to_curve.rs:9-82, ISOGENY_COEFFICIENTS_G2);
  * tests/golden/bundled_inputs.json -- the inputs of the reference's bundled run (main.rs:8-55): the 512 compressed
    public keys of light_client_update_period_1052.json's next sync committee, and the sync aggregate (bits + signature)
    and attested beacon header of light_client_update_period_1053.json, with the values derived from them by
    starky_bls12_381_b200/bls.py (signing root, aggregated public key, hashed message, signature point) as regression
    anchors.  The derived values are accepted only if the pairing product they define final-exponentiates to ONE.

    python tests/golden/make_bundled_fixture.py
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"
DOMAIN = bytes.fromhex("070000006a95a1a967855d676d48be69883b712607f952d5198d0f5677564636")      # main.rs:30


def iso_table():
    src = open(os.path.join(REF, "hash_to_curve.rs")).read()
    body = src[src.index("ISOGENY_COEFFICIENTS_G2"):src.index("pub fn map_to_curve_simple_swu_9mod16")]
    nums = re.findall(r'"(\d+)"', body)
    assert len(nums) == 32
    it = iter(nums)
    return [[[next(it), next(it)] for _ in range(4)] for _ in range(4)]


def main():
    with open(os.path.join(ROOT, "starky_bls12_381_b200", "witness", "iso_g2.json"), "w") as f:
        json.dump({"ISOGENY_COEFFICIENTS_G2": iso_table(), "source": "/root/reference/src/hash_to_curve.rs:9-82"}, f, indent=1)
    from starky_bls12_381_b200 import bls
    prev = json.load(open(os.path.join(REF, "light_client_update_period_1052.json")))["data"]
    cur = json.load(open(os.path.join(REF, "light_client_update_period_1053.json")))["data"]
    header = cur["attested_header"]["beacon"]
    root = bls.signing_root(header, DOMAIN)
    agg = cur["sync_aggregate"]
    inp = bls.prepare(prev["next_sync_committee"]["pubkeys"], agg["sync_committee_bits"], agg["sync_committee_signature"], root)
    assert bls.pairing_product_is_one(inp), "the derived inputs do not satisfy the pairing equation"
    # the committee's own aggregate key cross-checks the decompression and the G1 addition
    full = None
    for pt in inp["points"]:
        full = bls.G1.add(full, pt)
    assert full == bls.g1_decompress(bytes.fromhex(prev["next_sync_committee"]["aggregate_pubkey"][2:]))
    out = {
        "source": "reference src/light_client_update_period_1052.json (next_sync_committee.pubkeys) and _1053.json (sync_aggregate, attested_header.beacon); main.rs:8-55",
        "pubkeys": prev["next_sync_committee"]["pubkeys"], "sync_committee_bits": agg["sync_committee_bits"],
        "sync_committee_signature": agg["sync_committee_signature"], "attested_header": header, "domain": "0x" + DOMAIN.hex(),
        "derived": {"signing_root": "0x" + root.hex(), "participants": int(sum(inp["bits"])),
                    "apk": [str(v) for v in inp["apk"]],
                    "q1": [[str(v) for v in c] for c in inp["q1"]], "q2": [[str(v) for v in c] for c in inp["q2"]]},
    }
    path = os.path.join(HERE, "bundled_inputs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes;", out["derived"]["participants"], "participants; pairing product == 1")


if __name__ == "__main__":
    main()
