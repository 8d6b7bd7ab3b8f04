"""Offline look at the memory behaviour of the run-form quotient program (csrc/quotient.cu) of one stark: the sequence of
column slices a block touches, cut into the chunks the launcher uses, against an LRU window of K slices.
    python tools/perf/quotient_locality.py final_exp [chunks]
Prints: distinct columns per chunk summed over chunks / C (traffic floor of this chunking), LRU miss factors."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import starky_bls12_381_b200 as sb
from starky_bls12_381_b200 import airfiles
from test_air_programs import run_form


def accesses(rf, C):
    """per group: array of column ids in the order the kernel loads them"""
    code, gpc = rf["code"], [int(x) for x in rf["gpc"]]
    out = []
    for g in range(len(gpc) - 1):
        pc, end, acc = gpc[g], gpc[g + 1], []
        while pc < end:
            w = code[pc]
            op = w & 15
            if op == 12:
                kind, count = (w >> 4) & 7, (w >> 16) & 0xFFF
                nops = {0: 2, 1: 1, 2: 5, 3: 5}[kind]
                cols = []
                for k in range(nops):
                    half = (code[pc + 1 + k // 2] >> (32 * (k % 2))) & 0xFFFFFFFF
                    d = half >> 20
                    d = d - 4096 if d >= 2048 else d
                    v = half & 0xFFFFF
                    if v < 2 * C:
                        cols.append((v % C) + d * np.arange(count))
                if cols:
                    acc.append(np.stack(cols, 1).reshape(-1))
                pc += 4
                continue
            pc += 2 if op == 8 else 1
            vs = [(w >> 8) & 0x3FFFF, (w >> 26) & 0x3FFFF, (w >> 44) & 0x3FFFF]
            nv = {1: 1, 2: 2, 3: 1, 4: 1, 5: 2, 6: 1, 7: 2, 8: 3}.get(op, 0)
            acc.append(np.array([v % C for v in vs[:nv] if v < 2 * C], dtype=np.int64))
        out.append(np.concatenate(acc) if acc else np.zeros(0, np.int64))
    return out, gpc


def lru_misses(seq, cap):
    from collections import OrderedDict
    d, miss = OrderedDict(), 0
    for c in seq:
        c = int(c)
        if c in d:
            d.move_to_end(c)
        else:
            miss += 1
            d[c] = 1
            if len(d) > cap:
                d.popitem(last=False)
    return miss


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "final_exp"
    info = sb.STARKS[name]
    C = info.columns
    rf = run_form(open(airfiles.air_path(name, "airbin"), "rb").read())
    per_group, gpc = accesses(rf, C)
    N = info.num_rows << info.rate_bits
    xt = N // 128
    want = int(sys.argv[2]) if len(sys.argv) > 2 else (148 * 128 + xt - 1) // xt
    n_code = gpc[-1]
    # the launcher's cut: remaining code evenly over the remaining chunks, at group boundaries
    chunks, g = [], 0
    ng = len(gpc) - 1
    for c in range(min(want, ng)):
        if g >= ng:
            break
        pc0 = gpc[g]
        target = pc0 + (n_code - pc0) // (want - c)
        g1 = g + 1
        while g1 < ng and gpc[g1] < target:
            g1 += 1
        if c + 1 == want:
            g1 = ng
        chunks.append((g, g1))
        g = g1
    total = sum(len(a) for a in per_group)
    used = np.unique(np.concatenate(per_group))
    print("%s: C=%d columns (%d referenced), %d groups, %d accesses, %d chunks" % (name, C, len(used), ng, total, len(chunks)))
    seqs = [np.concatenate(per_group[a:b]) if b > a else np.zeros(0, np.int64) for a, b in chunks]
    distinct = sum(len(np.unique(s)) for s in seqs)
    print("  distinct columns per chunk, summed: %d = %.2f x C   (span of a chunk: median %d columns)" % (
        distinct, distinct / C, int(np.median([s.max() - s.min() + 1 for s in seqs if len(s)]))))
    for cap in (16, 32, 64, 128, 256, 1024):
        m = sum(lru_misses(s, cap) for s in seqs)
        print("  LRU window of %4d slices per block: %d misses = %.2f x C" % (cap, m, m / C))


if __name__ == "__main__":
    main()
