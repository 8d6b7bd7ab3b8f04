"""Debugging aid: quotient values of a small program with the run form restricted to subsets of the loop kinds
(SB_RUN_KINDS), against the word-form interpreter (SB_QUOTIENT_VM=1)."""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if len(sys.argv) > 2 and sys.argv[1] == "child":
    import starky_bls12_381_b200 as sb
    name, log_n = sys.argv[2], int(sys.argv[3])
    info = sb.STARKS[name]
    p = sb.standard_params(info.stark_id, log_n)
    rng = np.random.default_rng(1)
    trace = rng.integers(0, 1 << 32, (info.columns, 1 << log_n), dtype=np.uint64)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    ctx = sb.Context(0)
    ctx.lde_commit(p, trace, want_lde=False, want_digests=False)
    al = np.array([0x1234567890ABCDEF % 0xFFFFFFFF00000001, 0x0FEDCBA987654321], np.uint64)
    q = ctx.quotient_values(p, pis, al)
    print(hex(int(np.bitwise_xor.reduce(q.reshape(-1)))), hex(int(q[0, 0])), hex(int(q[1, 5])))
    sys.exit(0)
for name, log_n in (("ecc_agg", 5), ("pairing_precomp", 4)):
    for env in ({"SB_QUOTIENT_VM": "1"}, {"SB_RUN_KINDS": "0"}, {"SB_RUN_KINDS": "1"}, {"SB_RUN_KINDS": "2"}, {"SB_RUN_KINDS": "4"},
                {"SB_RUN_KINDS": "8"}, {"SB_RUN_KINDS": "15"}):
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, __file__, "child", name, str(log_n)], env=e, capture_output=True, text=True)
        print(name, env, out.stdout.strip(), out.stderr.strip()[-300:])
