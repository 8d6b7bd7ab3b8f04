"""python -m tools.witness_offsets [--ref /root/reference/src]

Evaluates every `pub const NAME: usize` column / public-input offset of the reference's stark files (fp.rs, fp2.rs,
fp6.rs, fp12.rs, g1.rs, fp12_mul.rs, calc_pairing_precomp.rs, miller_loop.rs, final_exponentiate.rs,
ecc_aggregate.rs) with the airgen interpreter and writes them, keyed "<file>.<NAME>", to
starky_bls12_381_b200/witness/offsets.json, together with the Frobenius coefficient tables of native.rs
(native.rs:1052-1057,1069-1125,1148-1199).  These are derived numbers (a layout table), not source; the witness
generators in starky_bls12_381_b200/witness/ index the trace with them exactly as the reference's fill_* functions do.
Runs only where /root/reference is mounted; the output is committed.
"""
import argparse
import json
import os
import re
import sys

from tools.airgen.rustsym import Interp

FILES = ["fp", "fp2", "fp6", "fp12", "g1", "fp12_mul", "calc_pairing_precomp", "miller_loop", "final_exponentiate",
         "ecc_aggregate"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/src")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ap.add_argument("--out", default=os.path.join(root, "starky_bls12_381_b200", "witness", "offsets.json"))
    a = ap.parse_args()
    it = Interp(a.ref)
    out = {}
    for f in FILES:
        for name in it.files[f].const_src:
            try:
                v = it.const(name, f)
            except Exception as e:            # non-integer consts are not offsets
                print("skip %s.%s: %r" % (f, name, e), file=sys.stderr)
                continue
            if isinstance(v, int) and not isinstance(v, bool):
                out["%s.%s" % (f, name)] = v
    nat = it.files["native"].src
    tables = {}
    for ty, fn in (("Fp2", "forbenius_coefficients"), ("Fp6", "forbenius_coefficients_1"), ("Fp6", "forbenius_coefficients_2"),
                   ("Fp12", "forbenius_coefficients")):
        m = [x for x in re.finditer(r"fn %s\(" % fn, nat)]
        # pick the occurrence inside `impl <ty>`
        best = None
        for x in m:
            head = nat.rfind("impl ", 0, x.start())
            if re.match(r"impl\s+%s\s*\{" % ty, nat[head:head + 40]):
                best = x
                break
        assert best is not None, (ty, fn)
        b = nat.index("{", nat.index(")", best.end()))
        depth, q = 0, b
        while True:
            if nat[q] == "{": depth += 1
            elif nat[q] == "}":
                depth -= 1
                if depth == 0: break
            q += 1
        tables["%s.%s" % (ty, fn)] = [str(int(v)) for v in re.findall(r'from_str\("(\d+)"\)', nat[b:q])]
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"offsets": out, "tables": tables}, open(a.out, "w"), indent=0, sort_keys=True)
    print("%d offsets, tables: %s -> %s" % (len(out), {k: len(v) for k, v in tables.items()}, a.out))


if __name__ == "__main__":
    main()
