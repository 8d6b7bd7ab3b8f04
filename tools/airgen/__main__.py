"""python -m tools.airgen [--ref /root/reference/src] [--out starky_bls12_381_b200/air] [stark ...]

Extracts the five starks' constraint programs from the reference source by symbolic execution (rustsym.py), checks
them against SURVEY.md Appendix B's independent fingerprints, and writes per stark
    <name>.air.xz     flat DAG in emission order  (CPU oracle, tests)
    <name>.airbin.xz  grouped bytecode            (GPU quotient evaluator)
plus manifest.json (counts + fingerprints).  Runs only where /root/reference is mounted; outputs are committed.
"""
import argparse
import json
import lzma
import os
import sys
import time

from .compile import compile_program, write_airbin, write_flat_air
from .fingerprint import fingerprint
from .rustsym import Interp

# name -> (reference file, COLUMNS, PUBLIC_INPUTS, num_rows the reference instantiates, constraint_degree)
STARKS = {
    "fp12_mul": ("fp12_mul", 60285, 432, 16, 3),
    "pairing_precomp": ("calc_pairing_precomp", 29376, 4968, 1024, 4),
    "miller_loop": ("miller_loop", 97330, 5064, 1024, 3),
    "final_exp": ("final_exponentiate", 73527, 288, 8192, 5),
    "ecc_agg": ("ecc_aggregate", 3339, 12824, 8192, 4),
}
# SURVEY.md Appendix B (extracted there by an independent scan of the Rust source)
EXPECT = {
    "fp12_mul": dict(K=82560, plain=26916, transition=55644, first=0, last=0, uses_next=8064, cs_refs=377135969582744497, cs_class=5601190230),
    "pairing_precomp": dict(K=113634, plain=28944, transition=84474, first=216, last=0, uses_next=13176, cs_refs=282723424429217662, cs_class=11576248902),
    "miller_loop": dict(K=145574, plain=48012, transition=97562, first=0, last=0, uses_next=14832, cs_refs=337172101024110717, cs_class=17649424608),
    "final_exp": dict(K=360800, plain=119598, transition=224818, first=8192, last=8192, uses_next=45256, cs_refs=2027473019779257126, cs_class=110204239101),
    "ecc_agg": dict(K=20013, plain=1251, transition=18188, first=574, last=0, uses_next=14638, cs_refs=1110302787620028249, cs_class=379379123),
}


def xz(path):
    data = open(path, "rb").read()
    with open(path + ".xz", "wb") as f:
        f.write(lzma.compress(data, preset=6))
    os.remove(path)
    return len(data), os.path.getsize(path + ".xz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/src")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                                                  "starky_bls12_381_b200", "air"))
    ap.add_argument("starks", nargs="*", default=list(STARKS))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    man_path = os.path.join(a.out, "manifest.json")
    manifest = json.load(open(man_path)) if os.path.exists(man_path) else {}
    for name in a.starks:
        file, n_cols, n_pis, rows, degree = STARKS[name]
        t = time.time()
        it = Interp(a.ref)
        cons = it.trace_stark(file, n_cols, n_pis, num_rows=rows)
        fp = fingerprint(it.dag, cons)
        for k, v in EXPECT[name].items():
            if fp[k] != v:
                sys.exit("%s: fingerprint %s = %s, SURVEY Appendix B says %s" % (name, k, fp[k], v))
        prog = compile_program(it.dag, cons, n_cols, n_pis, degree)
        flat = os.path.join(a.out, name + ".air")
        binp = os.path.join(a.out, name + ".airbin")
        write_flat_air(it.dag, cons, n_cols, n_pis, degree, flat)
        write_airbin(prog, binp)
        s1, s2 = xz(flat), xz(binp)
        manifest[name] = dict(fp, n_cols=n_cols, n_public_inputs=n_pis, num_rows=rows, constraint_degree=degree,
                              dag_nodes=len(it.dag.nodes), code_words=len(prog.code), weight_slots=len(prog.slot_off) - 1,
                              groups=len(prog.group_pc) - 1, consts=len(prog.consts))
        print("%-16s K=%d nodes=%d code=%d words slots=%d groups=%d  air %d->%d B  airbin %d->%d B  %.1fs" % (
            name, fp["K"], len(it.dag.nodes), len(prog.code), len(prog.slot_off) - 1, len(prog.group_pc) - 1,
            s1[0], s1[1], s2[0], s2[1], time.time() - t), flush=True)
    json.dump(manifest, open(man_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
