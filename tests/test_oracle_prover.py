"""CPU-only: the oracle prover / verifier restatement is self-consistent on VALID traces of small AIRs
(prove -> verify accepts; any tampering or an invalid trace is rejected) and the constraint evaluators agree."""
import numpy as np
import pytest

import oracle_lib as O
import toy_air

P = O.P


@pytest.fixture(scope="module")
def airs(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("toyair"))
    return dict(fib=toy_air.fibonacci(d), limbs=toy_air.limbs(d, 4))


def params(air, log_n, **kw):
    return O.make_params(n_cols=air["n_cols"], n_pis=air["n_pis"], degree=air["degree"], rate_bits=air["rate_bits"],
                         log_n=log_n, **kw)


@pytest.mark.parametrize("name,log_n", [("fib", 5), ("fib", 10), ("limbs", 6)])
def test_valid_trace_proves_and_verifies(airs, name, log_n):
    air = airs[name]
    trace, pis = air["witness"](log_n)
    p = params(air, log_n)
    # every constraint vanishes on every row of a valid trace (transitions: all but the last row)
    n = 1 << log_n
    for row in (0, 1, n // 2, n - 2):
        vals = O.eval_constraints_row(air["flat"], trace[:, row], trace[:, (row + 1) % n], pis)
        hdr = np.fromfile(air["flat"], dtype=np.uint32, count=10)
        off = 40 + 8 * int(hdr[5]) + 12 * int(hdr[6])
        cls = np.fromfile(air["flat"], dtype=np.uint32, offset=off, count=2 * int(hdr[7])).reshape(-1, 2)[:, 0]
        for k, v in enumerate(vals):
            if cls[k] in (1, 2) or (cls[k] == 3 and row == 0):
                assert v == 0, (name, row, k)
    rc, words = O.prove(air["flat"], p, trace, pis)
    assert rc == 0, O.err()
    assert O.verify(air["flat"], p, words) == 0, O.err()
    l = O.layout(p)
    # tamper with one word of every region
    for off in (l.off_trace_cap, l.off_quotient_cap + 5, l.off_local + 1, l.off_next, l.off_quot_open, l.off_final_poly,
                l.off_pow, l.off_queries + l.q_trace_leaf, l.off_queries + l.q_trace_path + 2, l.off_queries + l.q_quot_leaf,
                l.off_queries + 3 * l.query_stride + l.q_quot_path, l.off_pis):
        bad = words.copy()
        bad[off] = (int(bad[off]) + 1) % P
        assert O.verify(air["flat"], p, bad) != 0, off
    if l.n_fri_rounds:
        for off in (l.off_fri_caps, l.off_queries + l.q_steps, l.off_queries + l.q_steps + 33):
            bad = words.copy()
            bad[off] = (int(bad[off]) + 1) % P
            assert O.verify(air["flat"], p, bad) != 0, off


def test_invalid_trace_is_rejected(airs):
    # quotient factor 3 < domain blow-up 4: trim_to_len catches the non-divisible quotient (starky's panic)
    air = airs["limbs"]
    trace, pis = air["witness"](6)
    trace[5, 17] = (int(trace[5, 17]) + 1) % P
    p = params(air, 6)
    rc, _ = O.prove(air["flat"], p, trace, pis)
    assert rc == -4 and "not divisible" in O.err()
    # with the benchmarking flag the prover truncates instead; the verifier must then reject
    p2 = params(air, 6, flags=1)
    rc, words = O.prove(air["flat"], p2, trace, pis)
    assert rc == 0
    assert O.verify(air["flat"], p2, words) != 0
    # quotient factor 2 == blow-up 2: nothing is trimmed, so (as in starky) the prover cannot notice; the verifier does
    air = airs["fib"]
    trace, pis = air["witness"](6)
    trace[1, 17] = (int(trace[1, 17]) + 1) % P
    p = params(air, 6)
    rc, words = O.prove(air["flat"], p, trace, pis)
    assert rc == 0
    assert O.verify(air["flat"], p, words) != 0


def test_pow_witness_is_smallest_and_fixed_witness_replays(airs):
    air = airs["fib"]
    trace, pis = air["witness"](5)
    p = params(air, 5, pow_bits=10)
    rc, words = O.prove(air["flat"], p, trace, pis)
    assert rc == 0
    l = O.layout(p)
    w = int(words[l.off_pow])
    p2 = params(air, 5, pow_bits=10, flags=2, fixed_pow_witness=w)
    rc, words2 = O.prove(air["flat"], p2, trace, pis)
    assert rc == 0 and np.array_equal(words, words2)
    for cand in range(w):                          # no smaller witness works
        p3 = params(air, 5, pow_bits=10, flags=2, fixed_pow_witness=cand)
        rc, _ = O.prove(air["flat"], p3, trace, pis)
        assert rc == -6
        if cand > 40: break
