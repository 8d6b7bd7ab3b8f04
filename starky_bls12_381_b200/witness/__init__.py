"""Witness (trace) generation for the reference's starks — the host-side mirror of `XStark::generate_trace` and of the
public-input assembly in /root/reference/src/aggregate_proof.rs:24-227.  Trace generation is row-sequential CPU work
in the reference and stays on the host here; it exists so that the GPU prover can be exercised
on VALID traces of the real starks (all constraints vanish, the proof verifies) without the Rust toolchain.

Every generator returns (trace, public_inputs): trace is column-major uint64 [COLUMNS, num_rows] — the layout of
`trace_rows_to_poly_values(trace)`, ready for `prove()` — and public_inputs is uint64 [PUBLIC_INPUTS].
"""
import numpy as np

from . import native as N
from .fills import F12, fill_trace_fp12_multiplication
from .native import NS, limbs

_FP12MUL = NS("fp12_mul")


def _flat(vals):
    out = []
    for v in vals:
        out += limbs(v)
    return out


def random_fp(rng):
    return int.from_bytes(rng.bytes(48), "little") % N.P


def random_fp12(rng):
    return tuple(random_fp(rng) for _ in range(12))


def fp12_mul_trace(x, y, num_rows=16):
    """FP12MulStark::generate_trace (fp12_mul.rs:44-48) + fp12_mul_main's public inputs (aggregate_proof.rs:124-151)."""
    tr = np.zeros((num_rows, _FP12MUL.TOTAL_COLUMNS), dtype=np.uint64)
    fill_trace_fp12_multiplication(tr, x, y, 0, 11, 0)
    pis = np.array(_flat(x) + _flat(y) + _flat(N.fp12_mul(x, y)), dtype=np.uint64)
    assert pis.size == _FP12MUL.PUBLIC_INPUTS
    return np.ascontiguousarray(tr.T), pis


_ML = NS("miller_loop")


def fill_trace_miller_loop(tr, x, y, ell_coeffs, start_row, end_row, col):
    """miller_loop.rs:87-146."""
    from .fills import (fill_trace_fp2_fp_mul, fill_trace_multiply_by_014, negate6_rows)
    M = _ML
    rows = slice(start_row, end_row + 1)
    tr[rows, col + M.PX_OFFSET:col + M.PX_OFFSET + 12] = limbs(x)
    tr[rows, col + M.PY_OFFSET:col + M.PY_OFFSET + 12] = limbs(y)
    f12 = N.FP12_ONE
    i = N.BLS_X.bit_length() - 2
    bitone = False
    n_ops = min((end_row + 1 - start_row) // 12, len(ell_coeffs))
    for j in range(n_ops):
        s_row, e_row = start_row + j * 12, start_row + (j + 1) * 12 - 1
        blk = slice(s_row, e_row + 1)
        if j == 0:
            tr[blk, col + M.FIRST_BIT_SELECTOR_OFFSET] = 1
        if i == 0:
            tr[blk, col + M.LAST_BIT_SELECTOR_OFFSET] = 1
        if bitone:
            tr[blk, col + M.BIT1_SELECTOR_OFFSET] = 1
        tr[blk, col + M.ELL_COEFFS_INDEX_OFFEST + j] = 1
        e = ell_coeffs[j]
        for k in range(3):
            tr[blk, col + M.ELL_COEFFS_OFFSET + k * 24:col + M.ELL_COEFFS_OFFSET + k * 24 + 24] = _flat(e[k])
        tr[blk, col + M.F12_OFFSET:col + M.F12_OFFSET + 144] = _flat(f12)
        if j != 0:
            tr[s_row, col + M.FIRST_ROW_SELECTOR_OFFSET] = 1
        fill_trace_fp2_fp_mul(tr, e[1], x, s_row, e_row, col + M.O1_CALC_OFFSET)
        o1 = N.fp2_mul_fp(e[1], x)
        fill_trace_fp2_fp_mul(tr, e[2], y, s_row, e_row, col + M.O4_CALC_OFFSET)
        o4 = N.fp2_mul_fp(e[2], y)
        fill_trace_multiply_by_014(tr, f12, e[0], o1, o4, s_row, e_row, col + M.F12_MUL_BY_014_OFFSET)
        f12 = N.fp12_multiply_by_014(f12, e[0], o1, o4)
        fill_trace_fp12_multiplication(tr, f12, f12, s_row, e_row, col + M.F12_SQ_CALC_OFFSET)
        f12_sq = N.fp12_mul(f12, f12)
        if ((N.BLS_X >> i) & 1) and not bitone:
            bitone = True
        elif j < len(ell_coeffs) - 1:
            f12 = f12_sq
            i -= 1
            bitone = False
    f12 = N.fp12_conjugate(f12)
    tr[rows, col + M.MILLER_LOOP_RES_OFFSET:col + M.MILLER_LOOP_RES_OFFSET + 144] = _flat(f12)
    negate6_rows(tr, tuple(f12[6:]), start_row, end_row, col + M.RES_CONJUGATE_OFFSET)


def miller_loop_trace(x, y, q, num_rows=1024):
    """MillerLoopStark::generate_trace (miller_loop.rs:157-160) + miller_loop_main's public inputs
    (aggregate_proof.rs:71-121).  x, y: the G1 point (Fp); q = (qx, qy, qz): the G2 point (Fp2 each)."""
    ell = N.calc_pairing_precomp(*q)
    res = N.miller_loop(x, y, *q)
    tr = np.zeros((num_rows, _ML.TOTAL_COLUMNS), dtype=np.uint64)
    fill_trace_miller_loop(tr, x, y, ell, 0, num_rows - 1, 0)
    pi = limbs(x) + limbs(y)
    for cs in ell:
        for f2 in cs:
            pi += _flat(f2)
    pi += _flat(res)
    pis = np.array(pi, dtype=np.uint64)
    assert pis.size == _ML.PUBLIC_INPUTS
    return np.ascontiguousarray(tr.T), pis


_FE = NS("final_exponentiate")


def final_exp_trace(x, num_rows=8192):
    """FinalExponentiateStark::generate_trace (final_exponentiate.rs:137-281) + final_exponentiate_main's public inputs
    (aggregate_proof.rs:153-184)."""
    from .fills import (fill_trace_cyclotomic_exp, fill_trace_cyclotomic_sq, fill_trace_fp12_conjugate,
                        fill_trace_fp12_forbenius_map)
    E = _FE
    tr = np.zeros((num_rows, E.TOTAL_COLUMNS), dtype=np.uint64)
    idx = np.arange(num_rows)
    tr[idx, E.FINAL_EXP_ROW_SELECTORS + idx] = 1
    tr[:, E.FINAL_EXP_INPUT_OFFSET:E.FINAL_EXP_INPUT_OFFSET + 144] = _flat(x)
    OP = E.FINAL_EXP_OP_OFFSET

    def out(res, name):
        c = getattr(E, "FINAL_EXP_%s_OFFSET" % name)
        tr[:, c:c + 144] = _flat(res)
        return res

    def span(name, nxt):
        return getattr(E, name + "_ROW"), (getattr(E, nxt + "_ROW") if nxt != "TOTAL" else E.TOTAL_ROW) - 1

    def frob(v, pw, name, nxt):                      # fill_trace_forbenius
        s, e = span(name, nxt)
        tr[s:e + 1, E.FINAL_EXP_FORBENIUS_MAP_SELECTOR] = 1
        fill_trace_fp12_forbenius_map(tr, v, pw, s, e, OP)
        return out(N.fp12_frobenius(v, pw), name)

    def mul(a, b, name, nxt):                        # fill_trace_mul
        s, e = span(name, nxt)
        tr[s:e + 1, E.FINAL_EXP_MUL_SELECTOR] = 1
        fill_trace_fp12_multiplication(tr, a, b, s, e, OP)
        return out(N.fp12_mul(a, b), name)

    def div(a, b, name, nxt):                        # fill_trace_div: res = a / b, trace proves res * b
        s, e = span(name, nxt)
        res = N.fp12_mul(a, N.fp12_inv(b))
        tr[s:e + 1, E.FINAL_EXP_MUL_SELECTOR] = 1
        fill_trace_fp12_multiplication(tr, res, b, s, e, OP)
        return out(res, name)

    def cexp(v, name, nxt):                          # fill_trace_cyc_exp
        s, e = span(name, nxt)
        tr[s:e + 1, E.FINAL_EXP_CYCLOTOMIC_EXP_SELECTOR] = 1
        return out(fill_trace_cyclotomic_exp(tr, v, s, e, OP), name)

    def conj(v, name):                               # fill_trace_conjugate
        row = getattr(E, name + "_ROW")
        tr[row, E.FINAL_EXP_CONJUGATE_SELECTOR] = 1
        return out(fill_trace_fp12_conjugate(tr, v, row, OP), name)

    def csq(v, name, nxt):                           # fill_trace_cyc_sq
        s, e = span(name, nxt)
        tr[s:e + 1, E.FINAL_EXP_CYCLOTOMIC_SQ_SELECTOR] = 1
        fill_trace_cyclotomic_sq(tr, v, s, e, OP)
        return out(N.fp12_cyclotomic_square(v), name)

    t0 = frob(x, 6, "T0", "T1")
    t1 = div(t0, x, "T1", "T2")
    t2 = frob(t1, 2, "T2", "T3")
    t3 = mul(t2, t1, "T3", "T4")
    t4 = cexp(t3, "T4", "T5")
    t5 = conj(t4, "T5")
    t6 = csq(t3, "T6", "T7")
    t7 = conj(t6, "T7")
    t8 = mul(t7, t5, "T8", "T9")
    t9 = cexp(t8, "T9", "T10")
    t10 = conj(t9, "T10")
    t11 = cexp(t10, "T11", "T12")
    t12 = conj(t11, "T12")
    t13 = cexp(t12, "T13", "T14")
    t14 = conj(t13, "T14")
    t15 = csq(t5, "T15", "T16")
    t16 = mul(t14, t15, "T16", "T17")
    t17 = cexp(t16, "T17", "T18")
    t18 = conj(t17, "T18")
    t19 = mul(t5, t12, "T19", "T20")
    t20 = frob(t19, 2, "T20", "T21")
    t21 = mul(t10, t3, "T21", "T22")
    t22 = frob(t21, 3, "T22", "T23")
    t23 = conj(t3, "T23")
    t24 = mul(t16, t23, "T24", "T25")
    t25 = frob(t24, 1, "T25", "T26")
    t26 = conj(t8, "T26")
    t27 = mul(t18, t26, "T27", "T28")
    t28 = mul(t27, t3, "T28", "T29")
    t29 = mul(t20, t22, "T29", "T30")
    t30 = mul(t29, t25, "T30", "T31")
    t31 = mul(t30, t28, "T31", "TOTAL")
    assert t31 == N.fp12_final_exponentiate(x)
    pis = np.array(_flat(x) + _flat(t31), dtype=np.uint64)
    assert pis.size == E.PUBLIC_INPUTS
    return np.ascontiguousarray(tr.T), pis
