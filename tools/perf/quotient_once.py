"""One trace commitment + one quotient evaluation of a stark shape on synthetic data (ncu target for quotient_vm_kernel)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import starky_bls12_381_b200 as sb
from starky_bls12_381_b200 import airfiles

name = sys.argv[1] if len(sys.argv) > 1 else "pairing_precomp"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
info = sb.STARKS[name]
airfiles.air_path(name, "airbin")
p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1)
rng = np.random.default_rng(1)
trace = rng.integers(0, 1 << 32, (info.columns, info.num_rows), dtype=np.uint64)
pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
ctx = sb.Context(0)
ctx.lde_commit(p, trace, want_lde=False, want_digests=False)
del trace
al = np.array([0x1234567890ABCDEF % 0xFFFFFFFF00000001, 0x0FEDCBA987654321], np.uint64)
for _ in range(reps):
    q = ctx.quotient_values(p, pis, al)
    print(name, "quotient ms", round(ctx.stage_ms("quotient"), 3), "checksum", hex(int(np.bitwise_xor.reduce(q.reshape(-1)))))
ctx.close()
