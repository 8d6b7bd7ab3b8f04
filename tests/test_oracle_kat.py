"""Pins the oracle's field + Poseidon layer against plonky2's own known-answer vectors:
the two Poseidon-12 permutation KATs from plonky2's poseidon_goldilocks unit tests, the round-constant
spot values, and the Goldilocks generator facts."""
import numpy as np

import oracle_lib as O

P = O.P

KAT_ZERO = [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca,
            0xd7709673896996dc, 0x46a84e87642f44ed, 0xd032648251ee0b3c, 0x1c687363b207df62,
            0xdf8565563e8045fe, 0x40f5b37ff4254dae, 0xd070f637b431067c, 0x1792b1c4342109d7]
KAT_IOTA = [0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2,
            0x414754bfebd051f0, 0xcb1f8980294a023f, 0x6eb2a9e4d54a9d0f, 0x1902bc3af467e056,
            0xf045d5eafdc6021f, 0xe4150f77caaa3be5, 0xc9bfd01d39b50cce, 0x5c0a27fcb0e1459b]


def test_poseidon_kats():
    assert [int(x) for x in O.permute([0] * 12)] == KAT_ZERO
    assert [int(x) for x in O.permute(list(range(12)))] == KAT_IOTA


def test_round_constants_recipe():
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("gen", os.path.join(O.ROOT, "tools", "gen_poseidon_constants.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    rc = gen.round_constants()
    assert rc[:4] == [0xb585f766f2144405, 0x7746a55f43921ad7, 0xb2fb0d31cee799b4, 0x0f6760a4803427d7]
    assert rc[11] == 0xc54302f225db2c76 and max(rc) == 0xfeed4db6919e5a7c and len(rc) == 360
    # the committed headers are the generator's output
    for rel in ("oracle/poseidon_rc.h", "starky_bls12_381_b200/csrc/poseidon_rc.h"):
        assert open(os.path.join(O.ROOT, rel)).read() == gen.header(rc)


def test_fast_and_equivalent_round_constant_tables():
    """poseidon_fast.h (sparse partial rounds for the round-1 kernels and the host transcript; POSEIDON_RC_EQ: the same
    permutation with ONE non-zero constant per partial round, used by the matrix-instruction leaf sponge) is the output of
    tools/gen_poseidon_fast.py, and both forms give the oracle's permutation."""
    import importlib.util, os, sys
    sys.path.insert(0, os.path.join(O.ROOT, "tools"))
    spec = importlib.util.spec_from_file_location("genfast", os.path.join(O.ROOT, "tools", "gen_poseidon_fast.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    t = gen.derive()
    t["rc_eq"] = gen.eq_constants(t["rc"])
    for rel in ("oracle/poseidon_fast.h", "starky_bls12_381_b200/csrc/poseidon_fast.h"):
        assert open(os.path.join(O.ROOT, rel)).read() == gen.header(t)
    eq = t["rc_eq"]
    assert len(eq) == 31 * 12 and eq[:48] == t["rc"][:48] and eq[27 * 12:30 * 12] == t["rc"][27 * 12:] and not any(eq[360:])
    for r in range(4, 26):
        assert not any(eq[12 * r + 1:12 * r + 12])                     # partial rounds: word 0 only
    rng = np.random.default_rng(5)
    for s in ([0] * 12, [P - 1] * 12, [int(x) % P for x in rng.integers(0, 1 << 64, 12, dtype=np.uint64)]):
        want = [int(x) for x in O.permute(s)]
        assert gen.permute_eq(s, eq) == want and gen.permute_fast(s, t) == want


def test_goldilocks_facts():
    L = O.lib()
    assert pow(7, (P - 1) >> 32, P) == 1753635133440165772
    for k in (1, 4, 10, 13, 15, 32):
        w = L.orc_gl_root(k)
        assert pow(w, 1 << k, P) == 1 and pow(w, 1 << (k - 1), P) == P - 1
    # 7 generates F_p^*: p-1 = 2^32 * 3 * 5 * 17 * 257 * 65537
    for q in (2, 3, 5, 17, 257, 65537):
        assert pow(7, (P - 1) // q, P) != 1


def test_gl_mul_matches_definition():
    L = O.lib()
    rng = np.random.default_rng(1)
    edge = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, (1 << 63), P >> 1]
    vals = edge + [int(x) % P for x in rng.integers(0, 1 << 64, 400, dtype=np.uint64)]
    for a in vals[:60]:
        for b in vals:
            assert L.orc_gl_mul(a, b) == (a * b) % P == L.orc_gl_mul_slow(a, b)


def test_sponge_modes():
    # hash_or_noop: <= 4 elements are copied; overwrite-mode keeps the tail of a short last chunk
    assert list(O.hash_or_noop([5, 6, 7])) == [5, 6, 7, 0]
    x = list(range(1, 12))
    s = O.permute(x[:8] + [0, 0, 0, 0])
    s[:3] = x[8:]
    assert list(O.hash_no_pad(x)) == list(O.permute(s)[:4])
    l, r = [1, 2, 3, 4], [5, 6, 7, 8]
    assert list(O.two_to_one(l, r)) == list(O.permute(l + r + [0] * 4)[:4])


def test_challenger_pops_from_the_end():
    out = O.challenger_run([1, 2, 3], 3)
    s = O.permute([1, 2, 3] + [0] * 9)
    assert [int(v) for v in out] == [int(s[7]), int(s[6]), int(s[5])]


def test_ntt_roundtrip_and_definition():
    rng = np.random.default_rng(2)
    n = 16
    v = (rng.integers(0, 1 << 63, (3, n), dtype=np.uint64) % np.uint64(P)).astype(np.uint64)
    f = O.ntt_batch(v)
    assert np.array_equal(O.ntt_batch(f, inverse=True), v)
    w = O.lib().orc_gl_root(4)
    for k in range(n):
        assert int(f[0, k]) == sum(int(v[0, i]) * pow(w, i * k, P) for i in range(n)) % P
