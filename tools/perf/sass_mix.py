"""Static SASS instruction mix of the main kernels of libstarkyb200.so (cuobjdump -sass, sm_100a) -> profiles/r2_sass_static_mix.txt.
STATIC counts of the compiled code, not executed counts (those come from the ncu source pages).

    python tools/perf/sass_mix.py > profiles/r2_sass_static_mix.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "starky_bls12_381_b200", "libstarkyb200.so")
WANT = ["lde8_kernel", "leaf_sponge_mm_het_kernel", "leaf_sponge_mm_kernelILi1E", "leaf_sponge_mm_kernelILi2E", "leaf_sponge_mm_kernelILi4E",
        "leaf_sponge_dp_kernel", "leaf_sponge_sp_kernel", "quotient_run_kernel"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)?)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
print("# Static SASS instruction mix of the main kernels (cuobjdump -sass starky_bls12_381_b200/libstarkyb200.so, sm_100a;")
print("# mnemonic = opcode with its first modifier), written by tools/perf/sass_mix.py.  STATIC counts of the compiled code, not")
print("# executed counts: the executed mixes quoted in DESIGN.md come from the ncu source pages (profiles/r2_*_final_exp.txt).")
print("# The leaf sponge's MDS layer is on the integer tensor-core path: IMMA.16832.U8.U8 (mma.sync.m16n8k32.u8.u8.s32); no")
print("# tcgen05 / TMA instructions are expected (64-bit modular integer arithmetic: IMAD.WIDE / IADD3 / PRMT / IMMA).")
for want in WANT:
    for name, c in funcs.items():
        if want in name and (not name.endswith("ELi0EEvPKmjjjPmS1_S2_") or "mm_" in name or True):
            if "mm_kernel" in name and "ELi0E" not in name:
                continue        # lab variants (DBG != 0) are not in the library
            tot = sum(c.values())
            print("\n== %s: %d instructions" % (name, tot))
            for op, k in c.most_common(16):
                print("  %-18s %5d  %5.1f %%" % (op, k, 100.0 * k / tot))
            break
