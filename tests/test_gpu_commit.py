"""GPU parity: trace commitment (K1 LDE + K2 leaf hash + K3 Merkle) and the stage-level NTT / Poseidon entry points,
through the C ABI, bit-exact against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
import starky_bls12_381_b200 as sb
from helpers import P, pos_to_leaf, random_trace, to_oracle_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def custom(log_n, n_cols, rate_bits, cap_height=4, degree=3):
    return sb.Params(sb.StarkId.CUSTOM, log_n, n_cols, 0, degree, rate_bits, cap_height, 2, 16, 84, 4, 5, 0, 0, 0)


def test_poseidon_permute_kats_and_random(ctx):
    rng = np.random.default_rng(3)
    states = np.zeros((1000, 12), np.uint64)
    states[1] = np.arange(12)
    states[2:] = random_trace(rng, 998, 0, full_width=True).reshape(-1)[:998 * 12].reshape(998, 12) if False else \
        (rng.integers(0, 1 << 63, (998, 12), dtype=np.uint64) % np.uint64(P))
    states[3] = P - 1
    got = ctx.poseidon_permute_batch(states)
    assert [int(x) for x in got[0]][:2] == [0x3c18a9786cb0b359, 0xc4055e3364a246c3]
    assert [int(x) for x in got[1]][:2] == [0xd64e1e3efc5b8e9e, 0x53666633020aaa47]
    for i in range(0, 1000, 37):
        assert np.array_equal(got[i], O.permute(states[i]))
    assert (got < np.uint64(P)).all()


@pytest.mark.parametrize("log_n", [1, 2, 4, 6, 10, 12, 13, 15])
def test_ntt_batch_matches_oracle(ctx, log_n):
    rng = np.random.default_rng(log_n)
    v = random_trace(rng, 5 if log_n < 15 else 2, log_n, full_width=True)
    f = ctx.ntt_batch(v)
    assert np.array_equal(f, O.ntt_batch(v))
    assert np.array_equal(ctx.ntt_batch(f, inverse=True), v)


@pytest.mark.parametrize("leaf_len", [1, 3, 4, 5, 8, 9, 16, 23, 135])
def test_hash_leaves_ragged_lengths(ctx, leaf_len):
    rng = np.random.default_rng(leaf_len)
    cols = random_trace(rng, leaf_len, 6, full_width=True)      # [leaf_len][64]
    got = ctx.hash_leaves(cols)
    for i in range(64):
        assert np.array_equal(got[i], O.hash_or_noop(cols[:, i])), (leaf_len, i)


@pytest.mark.parametrize("kernel", [1, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13])
@pytest.mark.parametrize("leaf_len,count", [(5, 33), (8, 64), (17, 1000), (24, 32), (135, 100), (1001, 70)])
def test_hash_leaves_every_kernel_variant(ctx, monkeypatch, kernel, leaf_len, count):
    """The three leaf-sponge kernels (one thread per leaf / 3 words per thread / 1 word per warp) are bit-identical
    to hash_or_noop on ragged chain lengths and leaf counts that are not multiples of the 32-leaf group."""
    monkeypatch.setenv("SB_LEAF_KERNEL", str(kernel))
    rng = np.random.default_rng(100 * leaf_len + count)
    cols = (rng.integers(0, 1 << 63, (leaf_len, count), dtype=np.uint64) * np.uint64(2)
            + rng.integers(0, 2, (leaf_len, count), dtype=np.uint64)) % np.uint64(P)
    cols[:, 0] = P - 1
    cols[:, -1] = 0
    got = ctx.hash_leaves(cols)
    for i in list(range(0, count, 7)) + [count - 1]:
        assert np.array_equal(got[i], O.hash_or_noop(cols[:, i])), (kernel, leaf_len, i)


@pytest.mark.parametrize("log_n,n_cols,rate_bits,full", [
    (4, 8, 1, False), (4, 61, 1, True), (5, 3, 2, True), (6, 37, 1, False), (7, 20, 2, True),
    (10, 19, 1, False), (10, 11, 2, True), (13, 3, 2, True), (3, 300, 3, False),
    (8, 5, 2, True), (9, 3, 1, True), (11, 2, 2, True), (12, 3, 1, True), (6, 3, 3, True)])
def test_lde_commit_matches_oracle(ctx, log_n, n_cols, rate_bits, full):
    rng = np.random.default_rng(1000 * log_n + n_cols)
    p = custom(log_n, n_cols, rate_bits)
    trace = random_trace(rng, n_cols, log_n, full_width=full)
    got = ctx.lde_commit(p, trace)
    want = O.lde_commit(to_oracle_params(p), trace, want_coeffs=True)
    perm = pos_to_leaf(log_n, rate_bits)
    # device position q holds plonky2 leaf perm[q]
    assert np.array_equal(got["lde"], want["leaves"].T[:, perm])
    assert np.array_equal(got["digests"], want["digests"])
    assert np.array_equal(got["cap"], want["cap"])
    # coefficients are kept in bit-reversed coefficient order
    from helpers import bitrev_perm
    assert np.array_equal(ctx.coeffs(p), want["coeffs"][:, bitrev_perm(log_n)])


@pytest.mark.parametrize("log_n,n_cols,rate_bits", [(5, 300, 3), (7, 203, 2), (6, 1037, 1), (11, 70, 4), (12, 141, 2)])
@pytest.mark.parametrize("layout", ["colmajor", "col_ptrs", "colmajor_old_kernels"])
def test_streamed_leaf_sponge_matches_oracle(ctx, monkeypatch, log_n, n_cols, rate_bits, layout):
    """The host-trace path hashes the leaves slab by slab behind the copy (capi.cu ingest_and_commit_trace): with small
    slabs every shape here is fed to the sponge in several launches (sp kernel: few leaves, dp kernel: many leaves, a
    last slab that is not a multiple of 8 columns) and must give the digests of the one-launch path and of the oracle."""
    rng = np.random.default_rng(77 * log_n + n_cols)
    p = custom(log_n, n_cols, rate_bits)
    trace = random_trace(rng, n_cols, log_n, full_width=True)
    want = O.lde_commit(to_oracle_params(p), trace)
    monkeypatch.setenv("SB_SLAB_BYTES", str(64 << log_n))          # 8 columns per slab
    monkeypatch.setenv("SB_HASH_GROUP_COLS", "64")                 # one sponge launch per 64 columns
    if layout == "colmajor_old_kernels":
        monkeypatch.setenv("SB_STREAM_OLD", "1")                    # round 2's sp / dp kernels carry the same state
    if layout.startswith("colmajor"):
        got = ctx.lde_commit(p, trace)
    else:
        cols = [np.ascontiguousarray(trace[c]).copy() for c in range(n_cols)]
        ptrs = np.array([c.ctypes.data for c in cols], dtype=np.uint64)
        got = ctx.lde_commit(p, ptrs, sb.TraceLayout.COLS_U64_PTRS)
    assert np.array_equal(got["digests"], want["digests"])
    assert np.array_equal(got["cap"], want["cap"])
    monkeypatch.setenv("SB_NO_STREAM_HASH", "1")
    assert np.array_equal(ctx.lde_commit(p, trace)["digests"], want["digests"])


def test_trace_layouts_agree(ctx):
    rng = np.random.default_rng(7)
    p = custom(6, 45, 1)
    trace = random_trace(rng, 45, 6)
    base = ctx.lde_commit(p, trace)["cap"]
    rows64 = np.ascontiguousarray(trace.T)
    rows32 = rows64.astype(np.uint32)
    assert np.array_equal(ctx.lde_commit(p, rows64, sb.TraceLayout.ROWMAJOR_U64)["cap"], base)
    assert np.array_equal(ctx.lde_commit(p, rows32, sb.TraceLayout.ROWMAJOR_U32)["cap"], base)
    cols = [np.ascontiguousarray(trace[c]).copy() for c in range(45)]
    ptrs = np.array([c.ctypes.data for c in cols], dtype=np.uint64)
    assert np.array_equal(ctx.lde_commit(p, ptrs, sb.TraceLayout.COLS_U64_PTRS)["cap"], base)
    ctx.trace_upload(p, trace)
    assert np.array_equal(ctx.lde_commit(p, None, sb.TraceLayout.DEVICE_COLMAJOR_U64)["cap"], base)


def test_lde_is_linear_at_full_width(ctx):
    """Size-independent property at a BASELINE-scale column height: LDE(a + b) = LDE(a) + LDE(b)."""
    rng = np.random.default_rng(11)
    p = custom(13, 4, 2)
    a = random_trace(rng, 4, 13, full_width=True)
    b = random_trace(rng, 4, 13, full_width=True)
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    la, lb, ls = (ctx.lde_commit(p, t)["lde"] for t in (a, b, s))
    assert np.array_equal(ls, ((la.astype(object) + lb.astype(object)) % P).astype(np.uint64))


def test_errors_are_codes_not_aborts(ctx):
    p = custom(20, 4, 1)
    with pytest.raises(sb.SbError) as e:
        ctx.lde_commit(p, np.zeros((4, 16), np.uint64))
    assert e.value.code == -1


def test_openings_stage_matches_direct_evaluation(ctx):
    """sb_openings (SURVEY 8b stage export): P_c(zeta), P_c(g zeta) of the committed trace against a plain Horner evaluation
    of the natural-order coefficients in F_p[X]/(X^2 - 7), in Python integers."""
    rng = np.random.default_rng(31)
    p = custom(6, 9, 1)
    trace = random_trace(rng, p.n_cols, p.log_n, full_width=True)
    ctx.lde_commit(p, trace, want_lde=False, want_digests=False)
    coeffs = O.ntt_batch(trace, inverse=True)                       # natural-order coefficients of every column
    zeta = [int(x) for x in rng.integers(1, 1 << 62, 2)]
    loc, nxt = ctx.openings(p, np.array(zeta, np.uint64))
    g = int(O.lib().orc_gl_root(p.log_n))

    def ext_mul(x, y):
        return ((x[0] * y[0] + 7 * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)

    def horner(c, z):
        acc = (0, 0)
        for a in reversed([int(v) for v in c]):
            acc = ext_mul(acc, z)
            acc = ((acc[0] + a) % P, acc[1])
        return acc
    zn = (zeta[0] * g % P, zeta[1] * g % P)
    for col in range(p.n_cols):
        assert tuple(int(v) for v in loc[col]) == horner(coeffs[col], tuple(zeta))
        assert tuple(int(v) for v in nxt[col]) == horner(coeffs[col], zn)


@pytest.mark.parametrize("log_n,rate_bits", [(5, 1), (6, 2), (10, 1), (13, 2)])
def test_fri_commit_stage_matches_oracle(ctx, log_n, rate_bits):
    """sb_fri_commit (SURVEY 8b stage export): round caps and final polynomial of fri_committed_trees for an injected
    polynomial and injected folding challenges == the oracle's (0, 1 and 2 arity-16 rounds)."""
    rng = np.random.default_rng(1000 + log_n)
    p = custom(log_n, 8, rate_bits)
    op = to_oracle_params(p)
    n = 1 << log_n
    coeffs = rng.integers(0, 1 << 63, (n, 2), dtype=np.uint64) % np.uint64(P)
    betas = rng.integers(0, 1 << 63, (4, 2), dtype=np.uint64) % np.uint64(P)
    caps, fin = ctx.fri_commit(p, coeffs, betas)
    want_caps, want_fin = O.fri_commit(op, coeffs, betas)
    assert caps.shape[0] == {5: 0, 6: 1, 10: 1, 13: 2}[log_n]      # ConstantArityBits(4, 5): MillerLoop-like shapes fold once
    assert np.array_equal(caps, want_caps) and np.array_equal(fin, want_fin)
