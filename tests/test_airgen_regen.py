"""The committed constraint programs are what tools/airgen extracts from the reference TODAY: re-run the symbolic
execution of the reference's eval_packed_generic bodies (/root/reference/src/*.rs) and byte-compare both emitted forms
(flat DAG for the oracle, grouped bytecode for the GPU) with starky_bls12_381_b200/air/*.xz.  Needs the reference tree,
so it runs in the build container only (/root/reference does not exist on the GPU box)."""
import lzma
import os
import tempfile

import pytest

from starky_bls12_381_b200 import airfiles

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


@pytest.mark.parametrize("name", ["fp12_mul", "ecc_agg", "pairing_precomp", "miller_loop", "final_exp"])
def test_regenerated_program_equals_committed_blob(name):
    from tools.airgen.__main__ import EXPECT, STARKS
    from tools.airgen.compile import compile_program, write_airbin, write_flat_air
    from tools.airgen.fingerprint import fingerprint
    from tools.airgen.rustsym import Interp
    file, n_cols, n_pis, rows, degree = STARKS[name]
    it = Interp(REF)
    cons = it.trace_stark(file, n_cols, n_pis, num_rows=rows)
    fp = fingerprint(it.dag, cons)
    for k, v in EXPECT[name].items():
        assert fp[k] == v, (name, k)
    prog = compile_program(it.dag, cons, n_cols, n_pis, degree)
    with tempfile.TemporaryDirectory() as d:
        flat, binp = os.path.join(d, "x.air"), os.path.join(d, "x.airbin")
        write_flat_air(it.dag, cons, n_cols, n_pis, degree, flat)
        write_airbin(prog, binp)
        for path, kind in ((flat, "air"), (binp, "airbin")):
            want = lzma.decompress(open(os.path.join(airfiles.AIR_DIR, "%s.%s.xz" % (name, kind)), "rb").read())
            assert open(path, "rb").read() == want, (name, kind)
