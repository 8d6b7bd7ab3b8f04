#!/usr/bin/env python3
"""Derive the sparse ("fast") partial-round form of Poseidon-12 over Goldilocks from the MDS matrix and the round
constants alone ("the fast tables must be derived from MDS + RC"; this is the factorisation of the
Poseidon paper's appendix B that plonky2 ships as FAST_PARTIAL_* tables).

The 22 partial rounds   x <- M . S0(x + c_r)        (S0 = x^7 on word 0 only)
are rewritten as        x <- x + K                  (12-word constant, once)
                        x <- P . x                  (P = diag(1, P^) dense 11x11, once)
                        for r in 0..21:   x0 <- x0^7 (+ a_r for r < 21);   x <- Sp_r . x
with Sp_r = [[25, w^_r], [v_r, I]]:   x0' = 25 x0 + sum_i w^_r[i] x_i,   x_i' = x_i + v_r[i] x0.
Derivation (checked below against the naive rounds and plonky2's two published permutation KATs):
  constants: going backwards, c_r = M . (M^-1 c_r); words 1..11 of M^-1 c_r commute with the previous round's S0 and are
             merged into c_{r-1}; word 0 becomes the post-S-box scalar a_{r-1}; what is left at round 0 is K.
  matrices:  N = [[m00, w], [v, N^]] = [[m00, w N^^-1], [v, I]] . diag(1, N^); diag(1, N^) commutes with S0 of the
             same round and is pushed into the previous round's matrix: N <- diag(1, N^) . M.  What is left is P.

    python tools/gen_poseidon_fast.py   -> writes oracle/poseidon_fast.h and starky_bls12_381_b200/csrc/poseidon_fast.h
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_poseidon_constants import round_constants  # noqa: E402

P = 0xFFFFFFFF00000001
CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
W, RF_HALF, RP = 12, 4, 22


def mds_matrix():
    """out[r] = sum_i CIRC[i] * s[(i + r) % 12] + 8 * s[0] * [r == 0]   ->   M[r][c] = CIRC[(c - r) % 12] (+8 at [0][0])"""
    m = [[CIRC[(c - r) % W] for c in range(W)] for r in range(W)]
    m[0][0] += 8
    return m


def mat_mul(a, b):
    n, k, m = len(a), len(b), len(b[0])
    return [[sum(a[i][t] * b[t][j] for t in range(k)) % P for j in range(m)] for i in range(n)]


def mat_vec(a, v):
    return [sum(x * y for x, y in zip(row, v)) % P for row in a]


def mat_inv(a):
    n = len(a)
    m = [list(row) + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(a)]
    for col in range(n):
        piv = next(r for r in range(col, n) if m[r][col] % P)
        m[col], m[piv] = m[piv], m[col]
        inv = pow(m[col][col], P - 2, P)
        m[col] = [x * inv % P for x in m[col]]
        for r in range(n):
            if r != col and m[r][col]:
                f = m[r][col]
                m[r] = [(x - f * y) % P for x, y in zip(m[r], m[col])]
    return [row[n:] for row in m]


def derive():
    rc = round_constants()
    M = mds_matrix()
    Minv = mat_inv(M)
    c = [rc[12 * (RF_HALF + r):12 * (RF_HALF + r) + 12] for r in range(RP)]
    # ---- constants, backwards
    k = list(c[RP - 1])
    post = [0] * RP                       # a_r, added to word 0 after the S-box of round r (a_21 = 0)
    for r in range(RP - 1, 0, -1):
        w = mat_vec(Minv, k)
        post[r - 1] = w[0]
        k = [c[r - 1][0]] + [(c[r - 1][i] + w[i]) % P for i in range(1, W)]
    first = k
    # ---- matrices, backwards
    N = M
    w_hat, vs = [None] * RP, [None] * RP
    for r in range(RP - 1, -1, -1):
        Nhat = [row[1:] for row in N[1:]]
        w = [N[0][1:]]
        w_hat[r] = mat_mul(w, mat_inv(Nhat))[0]
        vs[r] = [N[i][0] for i in range(1, W)]
        assert N[0][0] == 25
        Nd = [[1] + [0] * 11] + [[0] + row for row in Nhat]
        N = mat_mul(Nd, M)
    # the leftover block-diagonal matrix sits in front of round 0; N now holds diag(1, N^_0) . M, undo the last product
    init = [row[1:] for row in mat_mul(N, Minv)[1:]]            # P^[r][c]: x_r' = sum_c P^[r][c] x_c   (r, c in 1..11)
    # ---- derived tables for the warp-split GPU kernel (leafhash.cuh)
    # the INIT layer is folded into the linear layer of the last full round before it (round 3):
    #   x = diag(1, P^) (M v + FIRST) = D3 v + K3;  D3ROT[j][i] = D3[j][(j + i) % 12] (the kernel reads words rotated)
    Pd = [[1] + [0] * 11] + [[0] + row for row in init]
    d3 = mat_mul(Pd, M)
    k3 = mat_vec(Pd, first)
    d3rot = [[d3[j][(j + i) % W] for i in range(W)] for j in range(W)]
    # two-round look-ahead: sum_i w^_r[i] x_i(r) = sum_i w^_r[i] x_i(r-1) + U[r] y(r-1),  U[r] = sum_i w^_r[i] v_{r-1}[i]
    u = [0] + [sum(a * b for a, b in zip(w_hat[r], vs[r - 1])) % P for r in range(1, RP)]
    return dict(rc=rc, first=first, post=post, init=init, w_hat=w_hat, vs=vs, d3rot=d3rot, k3=k3, u=u, d3=d3)


# ---------------------------------------------------------------------------------------------------------
def sbox(x):
    return pow(x, 7, P)


def permute_naive(s, rc):
    s = list(s)
    M = mds_matrix()
    for r in range(30):
        s = [(x + rc[12 * r + i]) % P for i, x in enumerate(s)]
        if r < 4 or r >= 26:
            s = [sbox(x) for x in s]
        else:
            s[0] = sbox(s[0])
        s = mat_vec(M, s)
    return s


def permute_fast(s, t):
    rc, M = t["rc"], mds_matrix()
    s = list(s)
    for r in range(4):
        s = mat_vec(M, [sbox((x + rc[12 * r + i]) % P) for i, x in enumerate(s)])
    s = [(x + k) % P for x, k in zip(s, t["first"])]
    s = [s[0]] + [sum(t["init"][r][c] * s[c + 1] for c in range(11)) % P for r in range(11)]
    for r in range(RP):
        y = (sbox(s[0]) + t["post"][r]) % P
        s = [(25 * y + sum(w * x for w, x in zip(t["w_hat"][r], s[1:]))) % P] + [(x + v * y) % P for x, v in zip(s[1:], t["vs"][r])]
    for r in range(26, 30):
        s = mat_vec(M, [sbox((x + rc[12 * r + i]) % P) for i, x in enumerate(s)])
    return s


def permute_lookahead(s0, t):
    """The schedule of the warp-split kernel: D3/K3 folding and the two-round look-ahead of the word-0 dot product."""
    rc, M = t["rc"], mds_matrix()
    s = list(s0)
    for r in range(3):
        s = mat_vec(M, [sbox((x + rc[12 * r + i]) % P) for i, x in enumerate(s)])
    v = [sbox((x + rc[36 + i]) % P) for i, x in enumerate(s)]
    s = [(sum(M[0][c] * v[c] for c in range(W)) + t["first"][0]) % P] + \
        [(sum(t["d3rot"][j][i] * v[(j + i) % W] for i in range(W)) + t["k3"][j]) % P for j in range(1, W)]
    x0, xs = s[0], s[1:]
    E = {0: sum(w * x for w, x in zip(t["w_hat"][0], xs)) % P, 1: sum(w * x for w, x in zip(t["w_hat"][1], xs)) % P}
    y_prev = 0
    for r in range(RP):
        y = (sbox(x0) + t["post"][r]) % P
        x0 = (25 * y + E[r] + t["u"][r] * y_prev) % P
        xs = [(x + vv * y) % P for x, vv in zip(xs, t["vs"][r])]
        if r + 2 < RP:
            E[r + 2] = sum(w * x for w, x in zip(t["w_hat"][r + 2], xs)) % P
        y_prev = y
    s = [x0] + xs
    for r in range(26, 30):
        s = mat_vec(M, [sbox((x + rc[12 * r + i]) % P) for i, x in enumerate(s)])
    return s == permute_naive(s0, rc)


def eq_constants(rc):
    """Round constants with the same permutation but only ONE non-zero constant in each partial round (dense-MDS kernels,
    leafhash_mm.cuh).  x <- M . S0(x + c_r): the word-0 part of c_r is needed before the S-box, the rest commutes with S0
    and with M -- M . (0, c^_r) joins the constants of the next round.  Rows 4..25 become (k_r, 0, .., 0), row 26 absorbs
    what is left; rows 0..3 and 27..29 are unchanged; row 30 = zeros ("the round after the last")."""
    M = mds_matrix()
    eq = [list(rc[12 * r:12 * r + 12]) for r in range(30)] + [[0] * 12]
    pend = list(eq[4])
    for r in range(4, 26):
        eq[r] = [pend[0]] + [0] * 11
        pushed = mat_vec(M, [0] + pend[1:])
        pend = [(a + b) % P for a, b in zip(rc[12 * (r + 1):12 * (r + 2)], pushed)]
    eq[26] = pend
    return [v for row in eq for v in row]


def permute_eq(s, eq):
    s = list(s)
    M = mds_matrix()
    for r in range(30):
        s = [(x + eq[12 * r + i]) % P for i, x in enumerate(s)]
        if r < 4 or r >= 26:
            s = [sbox(x) for x in s]
        else:
            s[0] = sbox(s[0])
        s = mat_vec(M, s)
    return s


def header(t):
    def arr(name, vals, per=4):
        out = ["#define %s { \\" % name]
        for i in range(0, len(vals), per):
            out.append("  " + ", ".join("0x%016xULL" % v for v in vals[i:i + per]) + ", \\")
        out.append("}")
        return out
    lines = ["// GENERATED by tools/gen_poseidon_fast.py -- do not edit.",
             "// Sparse partial-round form of Poseidon-12 over Goldilocks, derived from the MDS matrix and the round constants",
             "// (the factorisation plonky2 ships as FAST_PARTIAL_*); see the generator for the derivation and its checks.",
             "#pragma once", "#include <stdint.h>",
             "// x += FIRST (12), x[1..] = INIT . x[1..] (row-major 11x11), then per round r: x0 = x0^7 + POST[r];",
             "// x0' = 25 x0 + sum_i WHAT[r][i] x[i+1];  x[i+1]' = x[i+1] + VS[r][i] x0."]
    lines += arr("POSEIDON_FAST_FIRST", t["first"])
    lines += arr("POSEIDON_FAST_POST", t["post"])
    lines += arr("POSEIDON_FAST_INIT", [v for row in t["init"] for v in row])
    lines += arr("POSEIDON_FAST_WHAT", [v for row in t["w_hat"] for v in row])
    lines += arr("POSEIDON_FAST_VS", [v for row in t["vs"] for v in row])
    lines += ["// warp-split kernel: round 3's linear layer with INIT folded in, x_j = sum_i D3ROT[j][i] v_{(j+i)%12} + K3[j];",
              "// U[r] = sum_i WHAT[r][i] VS[r-1][i] (two-round look-ahead of the word-0 dot product)."]
    lines += arr("POSEIDON_FAST_D3ROT", [v for row in t["d3rot"] for v in row])
    lines += arr("POSEIDON_FAST_D3", [v for row in t["d3"] for v in row])
    lines += arr("POSEIDON_FAST_K3", t["k3"])
    lines += arr("POSEIDON_FAST_U", t["u"])
    lines += ["// dense-MDS kernels: the same permutation with one non-zero constant per partial round (31 x 12, row 30 = zeros)."]
    lines += arr("POSEIDON_RC_EQ", t["rc_eq"])
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    import random
    t = derive()
    kat0 = permute_naive([0] * 12, t["rc"])
    assert kat0[0] == 0x3c18a9786cb0b359 and kat0[11] == 0x1792b1c4342109d7          # plonky2's published vector
    assert permute_naive(list(range(12)), t["rc"])[0] == 0xd64e1e3efc5b8e9e
    rnd = random.Random(1)
    for s in [[0] * 12, list(range(12)), [P - 1] * 12] + [[rnd.randrange(P) for _ in range(12)] for _ in range(20)]:
        assert permute_fast(s, t) == permute_naive(s, t["rc"]), "fast form differs from the naive rounds"
    assert permute_lookahead([rnd.randrange(P) for _ in range(12)], t), "look-ahead schedule differs"
    t["rc_eq"] = eq_constants(t["rc"])
    for s in [[0] * 12, [P - 1] * 12] + [[rnd.randrange(P) for _ in range(12)] for _ in range(10)]:
        assert permute_eq(s, t["rc_eq"]) == permute_naive(s, t["rc"]), "equivalent constants differ from the naive rounds"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel in ("oracle/poseidon_fast.h", "starky_bls12_381_b200/csrc/poseidon_fast.h"):
        with open(os.path.join(root, rel), "w") as f:
            f.write(header(t))
    print("ok: fast form == naive form on 23 states; vs[21][:2] =", [hex(v) for v in t["vs"][21][:2]])
