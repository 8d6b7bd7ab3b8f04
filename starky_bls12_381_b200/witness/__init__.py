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


_PP = NS("calc_pairing_precomp")


def pairing_precomp_trace(x, y, z, num_rows=1024):
    """PairingPrecompStark::generate_trace (calc_pairing_precomp.rs:150-366) + calc_pairing_precomp_main's public inputs
    (aggregate_proof.rs:24-69).  x, y, z: the projective G2 point (Fp2 each)."""
    from .fills import (add_red_rows, fill_multiply_by_b_trace, fill_trace_fp2_fp_mul, fill_trace_negate_fp2,
                        generate_trace_fp2_mul, sub_red_rows, _rep, F2)
    A = _PP
    tr = np.zeros((num_rows, A.TOTAL_COLUMNS), dtype=np.uint64)
    z_inv = N.fp2_inv(z)
    generate_trace_fp2_mul(tr, z, z_inv, 0, num_rows - 1, A.Z_MULT_Z_INV_OFFSET)
    generate_trace_fp2_mul(tr, x, z_inv, 0, num_rows - 1, A.X_MULT_Z_INV_OFFSET)
    generate_trace_fp2_mul(tr, y, z_inv, 0, num_rows - 1, A.Y_MULT_Z_INV_OFFSET)
    qx, qy, qz = N.fp2_mul(x, z_inv), N.fp2_mul(y, z_inv), (1, 0)          # calc_qs (native.rs:277-286)
    tr[:, A.QX_OFFSET:A.QX_OFFSET + 24] = _flat(qx)
    tr[:, A.QY_OFFSET:A.QY_OFFSET + 24] = _flat(qy)
    tr[:, A.QZ_OFFSET:A.QZ_OFFSET + 24] = _flat(qz)
    rx, ry, rz = qx, qy, qz
    bit_pos, bit1, num_coeffs = 62, False, 68

    def negate_rows(v, s, e, col):
        fill_trace_negate_fp2(tr, v, s, col)
        _rep(tr, s, e, col, F2.FP2_ADDITION_TOTAL)

    for n in range(num_rows // 12 + 1):
        s, end_row = n * 12, (n + 1) * 12
        blk = slice(s, min(end_row, num_rows))
        if n == 0:
            tr[blk, A.FIRST_LOOP_SELECTOR_OFFSET] = 1
        tr[blk, A.RX_OFFSET:A.RX_OFFSET + 24] = _flat(rx)
        tr[blk, A.RY_OFFSET:A.RY_OFFSET + 24] = _flat(ry)
        tr[blk, A.RZ_OFFSET:A.RZ_OFFSET + 24] = _flat(rz)
        if bit1:
            tr[blk, A.BIT1_SELECTOR_OFFSET] = 1
        if n < num_coeffs:
            tr[blk, A.ELL_COEFFS_IDX_OFFSET + n] = 1
        tr[s, A.FIRST_ROW_SELECTOR_OFFSET] = 1
        if end_row > num_rows:
            break
        e = end_row - 1
        if not bit1:
            v = N.calc_precomp_stuff_loop0(rx, ry, rz)
            generate_trace_fp2_mul(tr, ry, ry, s, e, A.T0_CALC_OFFSET)
            generate_trace_fp2_mul(tr, rz, rz, s, e, A.T1_CALC_OFFSET)
            fill_trace_fp2_fp_mul(tr, v[4], 3, s, e, A.X0_CALC_OFFSET)
            fill_multiply_by_b_trace(tr, v[5], s, e, A.T2_CALC_OFFSET)
            fill_trace_fp2_fp_mul(tr, v[6], 3, s, e, A.T3_CALC_OFFSET)
            generate_trace_fp2_mul(tr, ry, rz, s, e, A.X1_CALC_OFFSET)
            fill_trace_fp2_fp_mul(tr, v[8], 2, s, e, A.T4_CALC_OFFSET)
            sub_red_rows(tr, v[6], v[3], s, e, A.X2_CALC_OFFSET)
            generate_trace_fp2_mul(tr, rx, rx, s, e, A.X3_CALC_OFFSET)
            fill_trace_fp2_fp_mul(tr, v[10], 3, s, e, A.X4_CALC_OFFSET)
            negate_rows(v[9], s, e, A.X5_CALC_OFFSET)
            sub_red_rows(tr, v[3], v[7], s, e, A.X6_CALC_OFFSET)
            generate_trace_fp2_mul(tr, rx, ry, s, e, A.X7_CALC_OFFSET)
            generate_trace_fp2_mul(tr, v[14], v[15], s, e, A.X8_CALC_OFFSET)
            add_red_rows(tr, v[3], v[7], s, e, A.X9_CALC_OFFSET)
            fill_trace_fp2_fp_mul(tr, v[17], N.HALF, s, e, A.X10_CALC_OFFSET)
            generate_trace_fp2_mul(tr, v[18], v[18], s, e, A.X11_CALC_OFFSET)
            generate_trace_fp2_mul(tr, v[6], v[6], s, e, A.X12_CALC_OFFSET)
            fill_trace_fp2_fp_mul(tr, v[20], 3, s, e, A.X13_CALC_OFFSET)
            fill_trace_fp2_fp_mul(tr, v[16], N.HALF, s, e, A.NEW_RX_OFFSET)
            sub_red_rows(tr, v[19], v[21], s, e, A.NEW_RY_OFFSET)
            generate_trace_fp2_mul(tr, v[3], v[9], s, e, A.NEW_RZ_OFFSET)
            rx, ry, rz = v[0], v[1], v[2]
            bit1 = bool((N.BLS_X >> bit_pos) & 1)
            bit_pos = bit_pos if bit1 else max(bit_pos - 1, 0)
        else:
            w = N.calc_precomp_stuff_loop1(rx, ry, rz, qx, qy)
            generate_trace_fp2_mul(tr, qy, rz, s, e, A.BIT1_T0_CALC_OFFSET)
            sub_red_rows(tr, ry, w[3], s, e, A.BIT1_T1_CALC_OFFSET)
            generate_trace_fp2_mul(tr, qx, rz, s, e, A.BIT1_T2_CALC_OFFSET)
            sub_red_rows(tr, rx, w[5], s, e, A.BIT1_T3_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[4], qx, s, e, A.BIT1_T4_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[6], qy, s, e, A.BIT1_T5_CALC_OFFSET)
            sub_red_rows(tr, w[7], w[8], s, e, A.BIT1_T6_CALC_OFFSET)
            negate_rows(w[4], s, e, A.BIT1_T7_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[6], w[6], s, e, A.BIT1_T8_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[11], w[6], s, e, A.BIT1_T9_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[11], rx, s, e, A.BIT1_T10_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[4], w[4], s, e, A.BIT1_T11_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[14], rz, s, e, A.BIT1_T12_CALC_OFFSET)
            fill_trace_fp2_fp_mul(tr, w[13], 2, s, e, A.BIT1_T13_CALC_OFFSET)
            sub_red_rows(tr, w[12], w[16], s, e, A.BIT1_T14_CALC_OFFSET)
            add_red_rows(tr, w[17], w[15], s, e, A.BIT1_T15_CALC_OFFSET)
            sub_red_rows(tr, w[13], w[18], s, e, A.BIT1_T16_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[19], w[4], s, e, A.BIT1_T17_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[12], ry, s, e, A.BIT1_T18_CALC_OFFSET)
            generate_trace_fp2_mul(tr, w[6], w[18], s, e, A.BIT1_RX_CALC_OFFSET)
            sub_red_rows(tr, w[20], w[21], s, e, A.BIT1_RY_CALC_OFFSET)
            generate_trace_fp2_mul(tr, rz, w[12], s, e, A.BIT1_RZ_CALC_OFFSET)
            rx, ry, rz = w[0], w[1], w[2]
            bit1 = False
            bit_pos = max(bit_pos - 1, 0)
    pi = _flat(x) + _flat(y) + _flat(z)
    for cs in N.calc_pairing_precomp(x, y, z):
        for f2 in cs:
            pi += _flat(f2)
    pis = np.array(pi, dtype=np.uint64)
    assert pis.size == A.PUBLIC_INPUTS
    return np.ascontiguousarray(tr.T), pis


_EC = NS("ecc_aggregate")
_G1 = NS("g1")


def fill_trace_g1_addition(tr, pt1, pt2, start_row, col):
    """g1.rs:26-255: chord addition of two affine points, x3 = l^2 - x2 - x1, y3 = l (x1 - x3) - y1."""
    from .fills import (F, fill_multiplication_trace_no_mod_reduction, fill_range_check_trace, fill_reduction_trace,
                        fill_trace_addition_fp, fill_trace_subtraction_fp, _rep)
    G, P = _G1, N.P
    x1, y1, x2, y2 = pt1[0], pt1[1], pt2[0], pt2[1]
    lam = N.fp_mul(N.fp_sub(y2, y1), N.fp_inv(N.fp_sub(x2, x1)))
    x3 = N.fp_sub(N.fp_sub(N.fp_mul(lam, lam), x2), x1)
    y3 = N.fp_sub(N.fp_mul(lam, N.fp_sub(x1, x3)), y1)
    s, e = start_row, start_row + 11
    tr[s:e + 1, col + G.G1_POINT_ADDITION_X1:col + G.G1_POINT_ADDITION_X1 + 72] = _flat((x1, y1, x2, y2, x3, y3))
    MULW = F.FP_MULTIPLICATION_TOTAL_COLUMNS

    def add_rows(a, b, c):
        fill_trace_addition_fp(tr, a, b, s, c)
        _rep(tr, s, e, c, F.FP_ADDITION_TOTAL)

    def sub_rows(a, b, c):
        fill_trace_subtraction_fp(tr, a, b, s, c)
        _rep(tr, s, e, c, F.FP_SUBTRACTION_TOTAL)

    def mul_red(a, b, c):
        fill_multiplication_trace_no_mod_reduction(tr, a, b, s, e, c)
        res = fill_reduction_trace(tr, a * b, s, e, c + MULW)
        fill_range_check_trace(tr, res, e, c + MULW + F.REDUCTION_TOTAL)
        return res

    add_rows(x2, P, col + G.X2_X1_DIFF)
    x2_x1 = x2 + P - x1
    sub_rows(x2 + P, x1, col + G.X2_X1_DIFF + F.FP_ADDITION_TOTAL)
    add_rows(y2, P, col + G.Y2_Y1_DIFF)
    y2_y1 = y2 + P - y1
    sub_rows(y2 + P, y1, col + G.Y2_Y1_DIFF + F.FP_ADDITION_TOTAL)
    x2_x1_sq = mul_red(x2_x1, x2_x1, col + G.X2_X1_SQ)
    y2_y1_sq = mul_red(y2_y1, y2_y1, col + G.Y2_Y1_SQ)
    add_rows(x1, x2, col + G.X1_X2_X3_SUM)
    add_rows(x1 + x2, x3, col + G.X1_X2_X3_SUM + F.FP_ADDITION_TOTAL)
    lhs = mul_red(x1 + x2 + x3, x2_x1_sq, col + G.X1_X2_X3_X2_X1_SQ)
    assert lhs == y2_y1_sq
    add_rows(y1, y3, col + G.Y1_Y3)
    add_rows(x1, P, col + G.X1_X3)
    x1_x3 = x1 + P - x3
    sub_rows(x1 + P, x3, col + G.X1_X3 + F.FP_ADDITION_TOTAL)
    a = mul_red(y1 + y3, x2_x1, col + G.Y1_Y3_X2_X1)
    b = mul_red(y2_y1, x1_x3, col + G.Y2_Y1_X1_X3)
    assert a == b
    return (x3, y3)


def ecc_aggregate_trace(points, bits, num_rows=8192):
    """ECCAggStark::generate_trace (ecc_aggregate.rs:37-82) + ec_aggregate_main's public inputs
    (aggregate_proof.rs:186-227).  points: 512 affine G1 points (x, y); bits: 512 participation bits."""
    E = _EC
    assert len(points) == E.NUM_POINTS == len(bits)
    assert (len(points) - 1) * 12 < num_rows, "stark doesn't have enough rows"
    tr = np.zeros((num_rows, E.TOTAL_COLUMNS), dtype=np.uint64)
    idx = np.arange(num_rows)
    tr[idx, E.ROW_NUM + idx % 12] = 1
    row = 0
    for i in range(E.NUM_POINTS):
        if i >= 2:
            row += 12
        tr[row:row + 12, E.PIS_IDX + i] = 1
    row = 0
    res = fill_trace_g1_addition(tr, points[0], points[1], row, E.OP)
    tr[row:row + 12, E.A_IS_INF] = int(not bits[0])
    tr[row:row + 12, E.B_IS_INF] = int(not bits[1])
    if not bits[0]:
        res = points[1]
    elif not bits[1]:
        res = points[0]
    for i in range(2, E.NUM_POINTS):
        row += 12
        tmp = fill_trace_g1_addition(tr, res, points[i], row, E.OP)
        tr[row:row + 12, E.A_IS_INF] = 0
        tr[row:row + 12, E.B_IS_INF] = int(not bits[i])
        if bits[i]:
            res = tmp
    pi = []
    for pt in points:
        pi += limbs(pt[0]) + limbs(pt[1])
    pi += [int(b) for b in bits]
    pi += limbs(res[0]) + limbs(res[1])
    pis = np.array(pi, dtype=np.uint64)
    assert pis.size == E.PUBLIC_INPUTS
    return np.ascontiguousarray(tr.T), pis, res
