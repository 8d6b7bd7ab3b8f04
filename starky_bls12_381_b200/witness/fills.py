"""Trace-filling gadgets: host-side restatement of the reference's fill_* functions
(/root/reference/src/fp.rs:185-428, fp2.rs:187-520, fp6.rs:124-443, fp12.rs:132-426), one Python function per Rust
function, same argument meaning.  `tr` is the row-major trace (numpy uint64 [rows, COLUMNS], the reference's
Vec<[F; COLUMNS]>); field elements are ints / tuples of ints (witness/native.py).

Gadgets the reference fills with the same values on every row of a block (`for row in start_row..end_row+1 { fill_x(row) }`)
are filled once and the column block is replicated (`_rep`).
"""
from .native import (NS, P, fp2_add, fp2_frobenius, fp2_mul, fp2_mul_by_nonresidue, fp2_mul_fp, fp2_neg, fp2_sub, fp4_square,
                     fp6_add, fp6_frobenius, fp6_mul, fp6_mul_by_nonresidue, fp6_multiply_by_01, fp6_multiply_by_1, fp6_neg,
                     fp6_parts, fp6_sub, fp12_cyclotomic_square, fp12_frobenius, fp12_mul, fp_neg, limbs, FP2_FROB, FP6_FROB_1,
                     FP6_FROB_2, FP12_FROB, BLS_X)

F = NS("fp")
F2 = NS("fp2")
F6 = NS("fp6")
F12 = NS("fp12")
M32 = 0xFFFFFFFF
RC_ADD = (1 << 382) - P          # fp.rs:1343 (range check addend 2^382 - p)
RED = F.FP_SINGLE_REDUCE_TOTAL + F.RANGE_CHECK_TOTAL


def put(tr, row, col, vals):
    """assign_u32_in_series (utils.rs:12-20)."""
    tr[row, col:col + len(vals)] = vals


def _rep(tr, start_row, end_row, col, width):
    if end_row > start_row:
        tr[start_row + 1:end_row + 1, col:col + width] = tr[start_row, col:col + width]


def add_carries(x, y, n):
    """add_u32_slices / add_u32_slices_12 (native.rs:69-100): limbs of x+y mod 2^(32n) and the carry out of each limb."""
    s = x + y
    out = limbs(s & ((1 << (32 * n)) - 1), n)
    car, c = [], 0
    for i in range(n):
        c = (((x >> (32 * i)) & M32) + ((y >> (32 * i)) & M32) + c) >> 32
        car.append(c)
    return out, car


def sub_borrows(x, y, n):
    """sub_u32_slices / _12 (native.rs:102-141), x >= y."""
    assert x >= y
    out = limbs(x - y, n)
    bor, b = [], 0
    for i in range(n):
        xi, yi = (x >> (32 * i)) & M32, (y >> (32 * i)) & M32
        b = 0 if xi >= yi + b else 1
        bor.append(b)
    return out, bor


# ------------------------------------------------------------------ fp.rs
def fill_addition_trace(tr, x, y, row, col):                                      # fp.rs:185-201 (24 limbs)
    tr[row, col + F.ADDITION_CHECK_OFFSET] = 1
    s, c = add_carries(x, y, 24)
    put(tr, row, col + F.ADDITION_X_OFFSET, limbs(x, 24))
    put(tr, row, col + F.ADDITION_Y_OFFSET, limbs(y, 24))
    put(tr, row, col + F.ADDITION_SUM_OFFSET, s)
    put(tr, row, col + F.ADDITION_CARRY_OFFSET, c)


def fill_trace_addition_fp(tr, x, y, row, col):                                   # fp.rs:204-220
    tr[row, col + F.FP_ADDITION_CHECK_OFFSET] = 1
    s, c = add_carries(x, y, 12)
    put(tr, row, col + F.FP_ADDITION_X_OFFSET, limbs(x))
    put(tr, row, col + F.FP_ADDITION_Y_OFFSET, limbs(y))
    put(tr, row, col + F.FP_ADDITION_SUM_OFFSET, s)
    put(tr, row, col + F.FP_ADDITION_CARRY_OFFSET, c)


def fill_trace_negate_fp(tr, x, row, col):                                        # fp.rs:223-234
    fill_trace_addition_fp(tr, x, fp_neg(x), row, col)


def fill_subtraction_trace(tr, x, y, row, col):                                   # fp.rs:237-253 (24 limbs)
    tr[row, col + F.SUBTRACTION_CHECK_OFFSET] = 1
    d, b = sub_borrows(x, y, 24)
    put(tr, row, col + F.SUBTRACTION_X_OFFSET, limbs(x, 24))
    put(tr, row, col + F.SUBTRACTION_Y_OFFSET, limbs(y, 24))
    put(tr, row, col + F.SUBTRACTION_DIFF_OFFSET, d)
    put(tr, row, col + F.SUBTRACTION_BORROW_OFFSET, b)


def fill_trace_subtraction_fp(tr, x, y, row, col):                                # fp.rs:256-272
    tr[row, col + F.FP_SUBTRACTION_CHECK_OFFSET] = 1
    d, b = sub_borrows(x, y, 12)
    assert b[11] == 0
    put(tr, row, col + F.FP_SUBTRACTION_X_OFFSET, limbs(x))
    put(tr, row, col + F.FP_SUBTRACTION_Y_OFFSET, limbs(y))
    put(tr, row, col + F.FP_SUBTRACTION_DIFF_OFFSET, d)
    put(tr, row, col + F.FP_SUBTRACTION_BORROW_OFFSET, b)


def fill_trace_multiply_single_fp(tr, x, y, row, col):                            # fp.rs:275-291 ; native.rs:143-155
    tr[row, col + F.FP_MULTIPLY_SINGLE_CHECK_OFFSET] = 1
    xl = limbs(x)
    res, car, c = [], [], 0
    for i in range(12):
        t = xl[i] * y + c
        res.append(t & M32)
        c = t >> 32
        car.append(c)
    assert c == 0
    put(tr, row, col + F.FP_MULTIPLY_SINGLE_X_OFFSET, xl)
    tr[row, col + F.FP_MULTIPLY_SINGLE_Y_OFFSET] = y
    put(tr, row, col + F.FP_MULTIPLY_SINGLE_SUM_OFFSET, res)
    put(tr, row, col + F.FP_MULTIPLY_SINGLE_CARRY_OFFSET, car)


def fill_trace_reduce_single(tr, x, row, col):                                    # fp.rs:294-312
    div, rem = divmod(x, P)
    assert div <= M32
    fill_trace_multiply_single_fp(tr, P, div, row, col + F.FP_SINGLE_REDUCE_MULTIPLICATION_OFFSET)
    put(tr, row, col + F.FP_SINGLE_REDUCE_X_OFFSET, limbs(x))
    put(tr, row, col + F.FP_SINGLE_REDUCED_OFFSET, limbs(rem))
    fill_trace_addition_fp(tr, div * P, rem, row, col + F.FP_SINGLE_REDUCTION_ADDITION_OFFSET)
    return rem


def fill_range_check_trace(tr, x, row, col):                                      # fp.rs:315-331
    s, c = add_carries(x, RC_ADD, 12)
    tr[row, col + F.RANGE_CHECK_SELECTOR_OFFSET] = 1
    put(tr, row, col + F.RANGE_CHECK_SUM_OFFSET, s)
    put(tr, row, col + F.RANGE_CHECK_SUM_CARRY_OFFSET, c)
    put(tr, row, col + F.RANGE_CHECK_BIT_DECOMP_OFFSET, [(s[11] >> i) & 1 for i in range(32)])


def fill_multiplication_trace_no_mod_reduction(tr, x, y, start_row, end_row, col):   # fp.rs:334-383
    tr[start_row, col + F.MULTIPLICATION_FIRST_ROW_OFFSET] = 1
    tr[start_row:start_row + 11, col + F.MULTIPLICATION_SELECTOR_OFFSET] = 1
    xl, yl = limbs(x), limbs(y)
    nrows = end_row + 1 - start_row
    tr[start_row:end_row + 1, col + F.X_INPUT_OFFSET:col + F.X_INPUT_OFFSET + 12] = xl
    tr[start_row:end_row + 1, col + F.Y_INPUT_OFFSET:col + F.Y_INPUT_OFFSET + 12] = yl
    for r in range(nrows):
        sel = 1 << r            # get_selector_bits_from_u32 keeps the low 12 bits (native.rs:250-259)
        put(tr, start_row + r, col + F.SELECTOR_OFFSET, [(sel >> i) & 1 for i in range(12)])
    prev = 0
    for i in range(12):
        # multiply_by_slice (native.rs:50-66)
        xy, car, c = [], [], 0
        for j in range(12):
            t = xl[j] * yl[i] + c
            xy.append(t & M32)
            c = t >> 32
            car.append(c)
        xy.append(c)
        r = start_row + i
        put(tr, r, col + F.XY_OFFSET, xy)
        put(tr, r, col + F.XY_CARRIES_OFFSET, car)
        shifted = (x * yl[i]) << (32 * i)
        put(tr, r, col + F.SHIFTED_XY_OFFSET, limbs(shifted, 24))
        s, cs = add_carries(shifted, prev, 24)
        put(tr, r, col + F.SUM_OFFSET, s)
        put(tr, r, col + F.SUM_CARRIES_OFFSET, cs)
        prev = shifted + prev
        assert prev < (1 << 768)


def fill_reduction_trace(tr, x, start_row, end_row, col):                         # fp.rs:386-424
    div, rem = divmod(x, P)
    fill_multiplication_trace_no_mod_reduction(tr, div, P, start_row, end_row, col + F.REDUCE_MULTIPLICATION_OFFSET)
    xl = limbs(x, 24)
    rl = limbs(rem)
    tr[start_row:end_row + 1, col + F.REDUCE_X_OFFSET:col + F.REDUCE_X_OFFSET + 24] = xl
    tr[start_row:end_row + 1, col + F.REDUCED_OFFSET:col + F.REDUCED_OFFSET + 12] = rl
    fill_addition_trace(tr, div * P, rem, start_row + 11, col + F.REDUCTION_ADDITION_OFFSET)
    return rem


# ------------------------------------------------------------------ fp2.rs
def fill_trace_addition_fp2(tr, x, y, row, col):                                  # fp2.rs:187-199
    fill_trace_addition_fp(tr, x[0], y[0], row, col + F2.FP2_ADDITION_0_OFFSET)
    fill_trace_addition_fp(tr, x[1], y[1], row, col + F2.FP2_ADDITION_1_OFFSET)


def fill_trace_subtraction_fp2(tr, x, y, row, col):                               # fp2.rs:202-214
    fill_trace_subtraction_fp(tr, x[0], y[0], row, col + F2.FP2_SUBTRACTION_0_OFFSET)
    fill_trace_subtraction_fp(tr, x[1], y[1], row, col + F2.FP2_SUBTRACTION_1_OFFSET)


def fill_trace_multiply_single_fp2(tr, x, y, row, col):                           # fp2.rs:217-229 (offsets as written there)
    fill_trace_multiply_single_fp(tr, x[0], y[0], row, col + F2.FP2_SUBTRACTION_0_OFFSET)
    fill_trace_multiply_single_fp(tr, x[1], y[1], row, col + F2.FP2_SUBTRACTION_1_OFFSET)


def fill_trace_negate_fp2(tr, x, row, col):                                       # fp2.rs:232-243
    fill_trace_addition_fp2(tr, x, fp2_neg(x), row, col)


def generate_trace_fp2_mul(tr, x, y, start_row, end_row, col):                    # fp2.rs:246-321
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F2.FP2_FP2_SELECTOR_OFFSET] = 1
    tr[rows, col + F2.FP2_FP2_X_INPUT_OFFSET:col + F2.FP2_FP2_X_INPUT_OFFSET + 24] = limbs(x[0]) + limbs(x[1])
    tr[rows, col + F2.FP2_FP2_Y_INPUT_OFFSET:col + F2.FP2_FP2_Y_INPUT_OFFSET + 24] = limbs(y[0]) + limbs(y[1])
    tr[end_row, col + F2.FP2_FP2_SELECTOR_OFFSET] = 0
    fill_multiplication_trace_no_mod_reduction(tr, x[0], y[0], start_row, end_row, col + F2.X_0_Y_0_MULTIPLICATION_OFFSET)
    fill_multiplication_trace_no_mod_reduction(tr, x[1], y[1], start_row, end_row, col + F2.X_1_Y_1_MULTIPLICATION_OFFSET)
    x0y0, x1y1 = x[0] * y[0], x[1] * y[1]
    fill_addition_trace(tr, x0y0, P * P, start_row + 11, col + F2.Z1_ADD_MODULUS_OFFSET)
    fill_subtraction_trace(tr, x0y0 + P * P, x1y1, start_row + 11, col + F2.Z1_SUBTRACTION_OFFSET)
    rem = fill_reduction_trace(tr, x0y0 + P * P - x1y1, start_row, end_row, col + F2.Z1_REDUCE_OFFSET)
    fill_range_check_trace(tr, rem, start_row, col + F2.Z1_RANGECHECK_OFFSET)
    fill_multiplication_trace_no_mod_reduction(tr, x[0], y[1], start_row, end_row, col + F2.X_0_Y_1_MULTIPLICATION_OFFSET)
    fill_multiplication_trace_no_mod_reduction(tr, x[1], y[0], start_row, end_row, col + F2.X_1_Y_0_MULTIPLICATION_OFFSET)
    x0y1, x1y0 = x[0] * y[1], x[1] * y[0]
    fill_addition_trace(tr, x0y1, x1y0, start_row + 11, col + F2.Z2_ADDITION_OFFSET)
    rem = fill_reduction_trace(tr, x0y1 + x1y0, start_row, end_row, col + F2.Z2_REDUCE_OFFSET)
    fill_range_check_trace(tr, rem, start_row, col + F2.Z2_RANGECHECK_OFFSET)


def fill_trace_fp2_fp_mul(tr, x, y, start_row, end_row, col):                     # fp2.rs:324-343
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F2.FP2_FP_MUL_SELECTOR_OFFSET] = 1
    tr[rows, col + F2.FP2_FP_X_INPUT_OFFSET:col + F2.FP2_FP_X_INPUT_OFFSET + 24] = limbs(x[0]) + limbs(x[1])
    tr[rows, col + F2.FP2_FP_Y_INPUT_OFFSET:col + F2.FP2_FP_Y_INPUT_OFFSET + 12] = limbs(y)
    tr[end_row, col + F2.FP2_FP_MUL_SELECTOR_OFFSET] = 0
    fill_multiplication_trace_no_mod_reduction(tr, x[0], y, start_row, end_row, col + F2.X0_Y_MULTIPLICATION_OFFSET)
    rem = fill_reduction_trace(tr, x[0] * y, start_row, end_row, col + F2.X0_Y_REDUCE_OFFSET)
    fill_range_check_trace(tr, rem, start_row, col + F2.X0_Y_RANGECHECK_OFFSET)
    fill_multiplication_trace_no_mod_reduction(tr, x[1], y, start_row, end_row, col + F2.X1_Y_MULTIPLICATION_OFFSET)
    rem = fill_reduction_trace(tr, x[1] * y, start_row, end_row, col + F2.X1_Y_REDUCE_OFFSET)
    fill_range_check_trace(tr, rem, start_row, col + F2.X1_Y_RANGECHECK_OFFSET)


def fill_trace_subtraction_with_reduction(tr, x, y, row, col):                    # fp2.rs:346-371
    fill_trace_addition_fp2(tr, x, (P, P), row, col)
    xm = (x[0] + P, x[1] + P)
    fill_trace_subtraction_fp2(tr, xm, y, row, col + F2.FP2_ADDITION_TOTAL)
    base = col + F2.FP2_ADDITION_TOTAL + F2.FP2_SUBTRACTION_TOTAL
    rem = fill_trace_reduce_single(tr, xm[0] - y[0], row, base)
    fill_range_check_trace(tr, rem, row, base + F.FP_SINGLE_REDUCE_TOTAL)
    rem = fill_trace_reduce_single(tr, xm[1] - y[1], row, base + RED)
    fill_range_check_trace(tr, rem, row, base + F.FP_SINGLE_REDUCE_TOTAL * 2 + F.RANGE_CHECK_TOTAL)


SUB_RED_FP2 = F2.FP2_ADDITION_TOTAL + F2.FP2_SUBTRACTION_TOTAL + 2 * RED
ADD_RED_FP2 = F2.FP2_ADDITION_TOTAL + 2 * RED


def fill_multiply_by_b_trace(tr, x, start_row, end_row, col):                     # fp2.rs:374-410
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F2.MULTIPLY_B_SELECTOR_OFFSET] = 1
    tr[rows, col + F2.MULTIPLY_B_X_OFFSET:col + F2.MULTIPLY_B_X_OFFSET + 24] = limbs(x[0]) + limbs(x[1])
    tr[end_row, col + F2.MULTIPLY_B_SELECTOR_OFFSET] = 0
    fill_multiplication_trace_no_mod_reduction(tr, x[0], 4, start_row, end_row, col + F2.MULTIPLY_B_X0_B_MUL_OFFSET)
    fill_multiplication_trace_no_mod_reduction(tr, x[1], 4, start_row, end_row, col + F2.MULTIPLY_B_X1_B_MUL_OFFSET)
    x0y, x1y = x[0] * 4, x[1] * 4
    fill_addition_trace(tr, x0y, P * P, start_row + 11, col + F2.MULTIPLY_B_ADD_MODSQ_OFFSET)
    fill_subtraction_trace(tr, x0y + P * P, x1y, start_row + 11, col + F2.MULTIPLY_B_SUB_OFFSET)
    rem = fill_reduction_trace(tr, x0y + P * P - x1y, start_row, end_row, col + F2.MULTIPLY_B_Z0_REDUCE_OFFSET)
    fill_range_check_trace(tr, rem, start_row, col + F2.MULTIPLY_B_Z0_RANGECHECK_OFFSET)
    fill_addition_trace(tr, x0y, x1y, start_row + 11, col + F2.MULTIPLY_B_ADD_OFFSET)
    rem = fill_reduction_trace(tr, x0y + x1y, start_row, end_row, col + F2.MULTIPLY_B_Z1_REDUCE_OFFSET)
    fill_range_check_trace(tr, rem, start_row, col + F2.MULTIPLY_B_Z1_RANGECHECK_OFFSET)


def fill_trace_addition_with_reduction(tr, x, y, row, col):                       # fp2.rs:413-429
    fill_trace_addition_fp2(tr, x, y, row, col)
    base = col + F2.FP2_ADDITION_TOTAL
    rem = fill_trace_reduce_single(tr, x[0] + y[0], row, base)
    fill_range_check_trace(tr, rem, row, base + F.FP_SINGLE_REDUCE_TOTAL)
    rem = fill_trace_reduce_single(tr, x[1] + y[1], row, base + RED)
    fill_range_check_trace(tr, rem, row, base + F.FP_SINGLE_REDUCE_TOTAL * 2 + F.RANGE_CHECK_TOTAL)


def fill_trace_non_residue_multiplication(tr, x, row, col):                       # fp2.rs:432-456
    tr[row, col + F2.FP2_NON_RESIDUE_MUL_CHECK_OFFSET] = 1
    put(tr, row, col + F2.FP2_NON_RESIDUE_MUL_INPUT_OFFSET, limbs(x[0]) + limbs(x[1]))
    fill_trace_addition_fp(tr, x[0], P, row, col + F2.FP2_NON_RESIDUE_MUL_C0_C1_SUB_OFFSET)
    fill_trace_subtraction_fp(tr, x[0] + P, x[1], row, col + F2.FP2_NON_RESIDUE_MUL_C0_C1_SUB_OFFSET + F.FP_ADDITION_TOTAL)
    rem = fill_trace_reduce_single(tr, x[0] + P - x[1], row, col + F2.FP2_NON_RESIDUE_MUL_Z0_REDUCE_OFFSET)
    fill_range_check_trace(tr, rem, row, col + F2.FP2_NON_RESIDUE_MUL_Z0_RANGECHECK_OFFSET)
    fill_trace_addition_fp(tr, x[0], x[1], row, col + F2.FP2_NON_RESIDUE_MUL_C0_C1_ADD_OFFSET)
    rem = fill_trace_reduce_single(tr, x[0] + x[1], row, col + F2.FP2_NON_RESIDUE_MUL_Z1_REDUCE_OFFSET)
    fill_range_check_trace(tr, rem, row, col + F2.FP2_NON_RESIDUE_MUL_Z1_RANGECHECK_OFFSET)


def _rows(fn, width):
    """`for row in start_row..end_row+1 { fn(trace, ..., row, col) }` with identical values on every row."""
    def run(tr, *args):
        *vals, start_row, end_row, col = args
        fn(tr, *vals, start_row, col)
        _rep(tr, start_row, end_row, col, width)
    return run


add_red_rows = _rows(fill_trace_addition_with_reduction, ADD_RED_FP2)
sub_red_rows = _rows(fill_trace_subtraction_with_reduction, SUB_RED_FP2)
nonres_rows = _rows(fill_trace_non_residue_multiplication, F2.FP2_NON_RESIDUE_MUL_TOTAL)


def fill_trace_fp4_sq(tr, x, y, start_row, end_row, col):                         # fp2.rs:459-502
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F2.FP4_SQ_INPUT_X_OFFSET:col + F2.FP4_SQ_INPUT_X_OFFSET + 24] = limbs(x[0]) + limbs(x[1])
    tr[rows, col + F2.FP4_SQ_INPUT_Y_OFFSET:col + F2.FP4_SQ_INPUT_Y_OFFSET + 24] = limbs(y[0]) + limbs(y[1])
    tr[rows, col + F2.FP4_SQ_SELECTOR_OFFSET] = 1
    tr[end_row, col + F2.FP4_SQ_SELECTOR_OFFSET] = 0
    t0 = fp2_mul(x, x)
    generate_trace_fp2_mul(tr, x, x, start_row, end_row, col + F2.FP4_SQ_T0_CALC_OFFSET)
    t1 = fp2_mul(y, y)
    generate_trace_fp2_mul(tr, y, y, start_row, end_row, col + F2.FP4_SQ_T1_CALC_OFFSET)
    t2 = fp2_mul_by_nonresidue(t1)
    nonres_rows(tr, t1, start_row, end_row, col + F2.FP4_SQ_T2_CALC_OFFSET)
    add_red_rows(tr, t2, t0, start_row, end_row, col + F2.FP4_SQ_X_CALC_OFFSET)
    t3 = fp2_add(x, y)
    add_red_rows(tr, x, y, start_row, end_row, col + F2.FP4_SQ_T3_CALC_OFFSET)
    t4 = fp2_mul(t3, t3)
    generate_trace_fp2_mul(tr, t3, t3, start_row, end_row, col + F2.FP4_SQ_T4_CALC_OFFSET)
    t5 = fp2_sub(t4, t0)
    sub_red_rows(tr, t4, t0, start_row, end_row, col + F2.FP4_SQ_T5_CALC_OFFSET)
    sub_red_rows(tr, t5, t1, start_row, end_row, col + F2.FP4_SQ_Y_CALC_OFFSET)


def fill_trace_fp2_forbenius_map(tr, x, pw, start_row, end_row, col):             # fp2.rs:505-531
    div, rem = pw // 2, pw % 2
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F2.FP2_FORBENIUS_MAP_INPUT_OFFSET:col + F2.FP2_FORBENIUS_MAP_INPUT_OFFSET + 24] = limbs(x[0]) + limbs(x[1])
    tr[rows, col + F2.FP2_FORBENIUS_MAP_SELECTOR_OFFSET] = 1
    tr[rows, col + F2.FP2_FORBENIUS_MAP_POW_OFFSET] = pw
    tr[rows, col + F2.FP2_FORBENIUS_MAP_DIV_OFFSET] = div
    tr[rows, col + F2.FP2_FORBENIUS_MAP_REM_OFFSET] = rem
    tr[end_row, col + F2.FP2_FORBENIUS_MAP_SELECTOR_OFFSET] = 0
    k = FP2_FROB[rem]
    fill_multiplication_trace_no_mod_reduction(tr, x[1], k, start_row, end_row, col + F2.FP2_FORBENIUS_MAP_T0_CALC_OFFSET)
    tr[start_row + 11, col + F2.FP2_FORBENIUS_MAP_MUL_RES_ROW] = 1
    base = col + F2.FP2_FORBENIUS_MAP_T0_CALC_OFFSET + F.FP_MULTIPLICATION_TOTAL_COLUMNS
    res = fill_reduction_trace(tr, x[1] * k, start_row, end_row, base)
    fill_range_check_trace(tr, res, start_row, base + F.REDUCTION_TOTAL)
    _rep(tr, start_row, end_row, base + F.REDUCTION_TOTAL, F.RANGE_CHECK_TOTAL)
    assert (x[0], res) == fp2_frobenius(x, pw)


# ------------------------------------------------------------------ fp6.rs
def fill_trace_addition_fp6(tr, x, y, row, col):                                  # fp6.rs:124-132
    for i, off in enumerate((F6.FP6_ADDITION_0_OFFSET, F6.FP6_ADDITION_1_OFFSET, F6.FP6_ADDITION_2_OFFSET)):
        fill_trace_addition_fp2(tr, x[2 * i:2 * i + 2], y[2 * i:2 * i + 2], row, col + off)


def fill_trace_subtraction_fp6(tr, x, y, row, col):                               # fp6.rs:175-183
    for i, off in enumerate((F6.FP6_SUBTRACTION_0_OFFSET, F6.FP6_SUBTRACTION_1_OFFSET, F6.FP6_SUBTRACTION_2_OFFSET)):
        fill_trace_subtraction_fp2(tr, x[2 * i:2 * i + 2], y[2 * i:2 * i + 2], row, col + off)


def fill_trace_addition_with_reduction_fp6(tr, x, y, row, col):                   # fp6.rs:135-148
    fill_trace_addition_fp6(tr, x, y, row, col)
    for i in range(6):
        base = col + F6.FP6_ADDITION_TOTAL + RED * i
        rem = fill_trace_reduce_single(tr, x[i] + y[i], row, base)
        fill_range_check_trace(tr, rem, row, base + F.FP_SINGLE_REDUCE_TOTAL)


def fill_trace_subtraction_with_reduction_fp6(tr, x, y, row, col):                # fp6.rs:151-172
    fill_trace_addition_fp6(tr, x, (P,) * 6, row, col)
    xm = tuple(v + P for v in x)
    fill_trace_subtraction_fp6(tr, xm, y, row, col + F6.FP6_ADDITION_TOTAL)
    for i in range(6):
        base = col + F6.FP6_ADDITION_TOTAL + F6.FP6_SUBTRACTION_TOTAL + RED * i
        rem = fill_trace_reduce_single(tr, xm[i] - y[i], row, base)
        fill_range_check_trace(tr, rem, row, base + F.FP_SINGLE_REDUCE_TOTAL)


def fill_trace_negate_fp6(tr, x, row, col):                                       # fp6.rs:186-196
    fill_trace_addition_fp6(tr, x, fp6_neg(x), row, col)


def fill_trace_non_residue_multiplication_fp6(tr, x, row, col):                   # fp6.rs:199-210
    tr[row, col + F6.FP6_NON_RESIDUE_MUL_CHECK_OFFSET] = 1
    for i in range(6):
        put(tr, row, col + F6.FP6_NON_RESIDUE_MUL_INPUT_OFFSET + i * 12, limbs(x[i]))
    fill_trace_non_residue_multiplication(tr, (x[4], x[5]), row, col + F6.FP6_NON_RESIDUE_MUL_C2)


ADD_RED_FP6 = F6.FP6_ADDITION_TOTAL + 6 * RED
SUB_RED_FP6 = F6.FP6_ADDITION_TOTAL + F6.FP6_SUBTRACTION_TOTAL + 6 * RED
add_red6_rows = _rows(fill_trace_addition_with_reduction_fp6, ADD_RED_FP6)
sub_red6_rows = _rows(fill_trace_subtraction_with_reduction_fp6, SUB_RED_FP6)
nonres6_rows = _rows(fill_trace_non_residue_multiplication_fp6, F6.FP6_NON_RESIDUE_MUL_TOTAL)
negate6_rows = _rows(fill_trace_negate_fp6, F6.FP6_ADDITION_TOTAL)


def _flat(vals):
    out = []
    for v in vals:
        out += limbs(v)
    return out


def fill_trace_fp6_multiplication(tr, x, y, start_row, end_row, col):             # fp6.rs:213-303
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F6.FP6_MUL_X_INPUT_OFFSET:col + F6.FP6_MUL_X_INPUT_OFFSET + 72] = _flat(x)
    tr[rows, col + F6.FP6_MUL_Y_INPUT_OFFSET:col + F6.FP6_MUL_Y_INPUT_OFFSET + 72] = _flat(y)
    tr[rows, col + F6.FP6_MUL_SELECTOR_OFFSET] = 1
    tr[end_row, col + F6.FP6_MUL_SELECTOR_OFFSET] = 0
    c0, c1, c2 = fp6_parts(x)
    r0, r1, r2 = fp6_parts(y)
    s, e = start_row, end_row
    t0 = fp2_mul(c0, r0); generate_trace_fp2_mul(tr, c0, r0, s, e, col + F6.FP6_MUL_T0_CALC_OFFSET)
    t1 = fp2_mul(c1, r1); generate_trace_fp2_mul(tr, c1, r1, s, e, col + F6.FP6_MUL_T1_CALC_OFFSET)
    t2 = fp2_mul(c2, r2); generate_trace_fp2_mul(tr, c2, r2, s, e, col + F6.FP6_MUL_T2_CALC_OFFSET)
    t3 = fp2_add(c1, c2); add_red_rows(tr, c1, c2, s, e, col + F6.FP6_MUL_T3_CALC_OFFSET)
    t4 = fp2_add(r1, r2); add_red_rows(tr, r1, r2, s, e, col + F6.FP6_MUL_T4_CALC_OFFSET)
    t5 = fp2_mul(t3, t4); generate_trace_fp2_mul(tr, t3, t4, s, e, col + F6.FP6_MUL_T5_CALC_OFFSET)
    t6 = fp2_sub(t5, t1); sub_red_rows(tr, t5, t1, s, e, col + F6.FP6_MUL_T6_CALC_OFFSET)
    t7 = fp2_sub(t6, t2); sub_red_rows(tr, t6, t2, s, e, col + F6.FP6_MUL_T7_CALC_OFFSET)
    t8 = fp2_mul_by_nonresidue(t7); nonres_rows(tr, t7, s, e, col + F6.FP6_MUL_T8_CALC_OFFSET)
    add_red_rows(tr, t8, t0, s, e, col + F6.FP6_MUL_X_CALC_OFFSET)
    t9 = fp2_add(c0, c1); add_red_rows(tr, c0, c1, s, e, col + F6.FP6_MUL_T9_CALC_OFFSET)
    t10 = fp2_add(r0, r1); add_red_rows(tr, r0, r1, s, e, col + F6.FP6_MUL_T10_CALC_OFFSET)
    t11 = fp2_mul(t9, t10); generate_trace_fp2_mul(tr, t9, t10, s, e, col + F6.FP6_MUL_T11_CALC_OFFSET)
    t12 = fp2_sub(t11, t0); sub_red_rows(tr, t11, t0, s, e, col + F6.FP6_MUL_T12_CALC_OFFSET)
    t13 = fp2_sub(t12, t1); sub_red_rows(tr, t12, t1, s, e, col + F6.FP6_MUL_T13_CALC_OFFSET)
    t14 = fp2_mul_by_nonresidue(t2); nonres_rows(tr, t2, s, e, col + F6.FP6_MUL_T14_CALC_OFFSET)
    add_red_rows(tr, t13, t14, s, e, col + F6.FP6_MUL_Y_CALC_OFFSET)
    t15 = fp2_add(c0, c2); add_red_rows(tr, c0, c2, s, e, col + F6.FP6_MUL_T15_CALC_OFFSET)
    t16 = fp2_add(r0, r2); add_red_rows(tr, r0, r2, s, e, col + F6.FP6_MUL_T16_CALC_OFFSET)
    t17 = fp2_mul(t15, t16); generate_trace_fp2_mul(tr, t15, t16, s, e, col + F6.FP6_MUL_T17_CALC_OFFSET)
    t18 = fp2_sub(t17, t0); sub_red_rows(tr, t17, t0, s, e, col + F6.FP6_MUL_T18_CALC_OFFSET)
    t19 = fp2_sub(t18, t2); sub_red_rows(tr, t18, t2, s, e, col + F6.FP6_MUL_T19_CALC_OFFSET)
    add_red_rows(tr, t19, t1, s, e, col + F6.FP6_MUL_Z_CALC_OFFSET)


def fill_trace_multiply_by_1(tr, x, b1, start_row, end_row, col):                 # fp6.rs:306-333
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F6.MULTIPLY_BY_1_INPUT_OFFSET:col + F6.MULTIPLY_BY_1_INPUT_OFFSET + 72] = _flat(x)
    tr[rows, col + F6.MULTIPLY_BY_1_B1_OFFSET:col + F6.MULTIPLY_BY_1_B1_OFFSET + 24] = _flat(b1)
    tr[rows, col + F6.MULTIPLY_BY_1_SELECTOR_OFFSET] = 1
    tr[end_row, col + F6.MULTIPLY_BY_1_SELECTOR_OFFSET] = 0
    c0, c1, c2 = fp6_parts(x)
    s, e = start_row, end_row
    t0 = fp2_mul(c2, b1); generate_trace_fp2_mul(tr, c2, b1, s, e, col + F6.MULTIPLY_BY_1_T0_CALC_OFFSET)
    nonres_rows(tr, t0, s, e, col + F6.MULTIPLY_BY_1_X_CALC_OFFSET)
    generate_trace_fp2_mul(tr, c0, b1, s, e, col + F6.MULTIPLY_BY_1_Y_CALC_OFFSET)
    generate_trace_fp2_mul(tr, c1, b1, s, e, col + F6.MULTIPLY_BY_1_Z_CALC_OFFSET)


def fill_trace_multiply_by_01(tr, x, b0, b1, start_row, end_row, col):            # fp6.rs:336-394
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F6.MULTIPLY_BY_01_INPUT_OFFSET:col + F6.MULTIPLY_BY_01_INPUT_OFFSET + 72] = _flat(x)
    tr[rows, col + F6.MULTIPLY_BY_01_B0_OFFSET:col + F6.MULTIPLY_BY_01_B0_OFFSET + 24] = _flat(b0)
    tr[rows, col + F6.MULTIPLY_BY_01_B1_OFFSET:col + F6.MULTIPLY_BY_01_B1_OFFSET + 24] = _flat(b1)
    tr[rows, col + F6.MULTIPLY_BY_01_SELECTOR_OFFSET] = 1
    tr[end_row, col + F6.MULTIPLY_BY_01_SELECTOR_OFFSET] = 0
    c0, c1, c2 = fp6_parts(x)
    s, e = start_row, end_row
    t0 = fp2_mul(c0, b0); generate_trace_fp2_mul(tr, c0, b0, s, e, col + F6.MULTIPLY_BY_01_T0_CALC_OFFSET)
    t1 = fp2_mul(c1, b1); generate_trace_fp2_mul(tr, c1, b1, s, e, col + F6.MULTIPLY_BY_01_T1_CALC_OFFSET)
    t2 = fp2_mul(c2, b1); generate_trace_fp2_mul(tr, c2, b1, s, e, col + F6.MULTIPLY_BY_01_T2_CALC_OFFSET)
    t3 = fp2_mul_by_nonresidue(t2); nonres_rows(tr, t2, s, e, col + F6.MULTIPLY_BY_01_T3_CALC_OFFSET)
    add_red_rows(tr, t3, t0, s, e, col + F6.MULTIPLY_BY_01_X_CALC_OFFSET)
    t4 = fp2_add(b0, b1); add_red_rows(tr, b0, b1, s, e, col + F6.MULTIPLY_BY_01_T4_CALC_OFFSET)
    t5 = fp2_add(c0, c1); add_red_rows(tr, c0, c1, s, e, col + F6.MULTIPLY_BY_01_T5_CALC_OFFSET)
    t6 = fp2_mul(t4, t5); generate_trace_fp2_mul(tr, t4, t5, s, e, col + F6.MULTIPLY_BY_01_T6_CALC_OFFSET)
    t7 = fp2_sub(t6, t0); sub_red_rows(tr, t6, t0, s, e, col + F6.MULTIPLY_BY_01_T7_CALC_OFFSET)
    sub_red_rows(tr, t7, t1, s, e, col + F6.MULTIPLY_BY_01_Y_CALC_OFFSET)
    t8 = fp2_mul(c2, b0); generate_trace_fp2_mul(tr, c2, b0, s, e, col + F6.MULTIPLY_BY_01_T8_CALC_OFFSET)
    add_red_rows(tr, t8, t1, s, e, col + F6.MULTIPLY_BY_01_Z_CALC_OFFSET)


def fill_trace_fp6_forbenius_map(tr, x, pw, start_row, end_row, col):             # fp6.rs:397-431
    div, rem = pw // 6, pw % 6
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F6.FP6_FORBENIUS_MAP_INPUT_OFFSET:col + F6.FP6_FORBENIUS_MAP_INPUT_OFFSET + 72] = _flat(x)
    tr[rows, col + F6.FP6_FORBENIUS_MAP_SELECTOR_OFFSET] = 1
    tr[rows, col + F6.FP6_FORBENIUS_MAP_POW_OFFSET] = pw
    tr[rows, col + F6.FP6_FORBENIUS_MAP_DIV_OFFSET] = div
    tr[rows, col + F6.FP6_FORBENIUS_MAP_REM_OFFSET] = rem
    tr[rows, col + F6.FP6_FORBENIUS_MAP_BIT0_OFFSET] = rem & 1
    tr[rows, col + F6.FP6_FORBENIUS_MAP_BIT1_OFFSET] = (rem >> 1) & 1
    tr[rows, col + F6.FP6_FORBENIUS_MAP_BIT2_OFFSET] = rem >> 2
    tr[end_row, col + F6.FP6_FORBENIUS_MAP_SELECTOR_OFFSET] = 0
    c0, c1, c2 = fp6_parts(x)
    s, e = start_row, end_row
    fill_trace_fp2_forbenius_map(tr, c0, pw, s, e, col + F6.FP6_FORBENIUS_MAP_X_CALC_OFFSET)
    t0 = fp2_frobenius(c1, pw)
    fill_trace_fp2_forbenius_map(tr, c1, pw, s, e, col + F6.FP6_FORBENIUS_MAP_T0_CALC_OFFSET)
    generate_trace_fp2_mul(tr, t0, FP6_FROB_1[pw % 6], s, e, col + F6.FP6_FORBENIUS_MAP_Y_CALC_OFFSET)
    t1 = fp2_frobenius(c2, pw)
    fill_trace_fp2_forbenius_map(tr, c2, pw, s, e, col + F6.FP6_FORBENIUS_MAP_T1_CALC_OFFSET)
    generate_trace_fp2_mul(tr, t1, FP6_FROB_2[pw % 6], s, e, col + F6.FP6_FORBENIUS_MAP_Z_CALC_OFFSET)


# ------------------------------------------------------------------ fp12.rs
def fill_trace_multiply_by_014(tr, x, o0, o1, o4, start_row, end_row, col):       # fp12.rs:132-183
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F12.MULTIPLY_BY_014_INPUT_OFFSET:col + F12.MULTIPLY_BY_014_INPUT_OFFSET + 144] = _flat(x)
    tr[rows, col + F12.MULTIPLY_BY_014_O0_OFFSET:col + F12.MULTIPLY_BY_014_O0_OFFSET + 24] = _flat(o0)
    tr[rows, col + F12.MULTIPLY_BY_014_O1_OFFSET:col + F12.MULTIPLY_BY_014_O1_OFFSET + 24] = _flat(o1)
    tr[rows, col + F12.MULTIPLY_BY_014_O4_OFFSET:col + F12.MULTIPLY_BY_014_O4_OFFSET + 24] = _flat(o4)
    tr[rows, col + F12.MULTIPLY_BY_014_SELECTOR_OFFSET] = 1
    tr[end_row, col + F12.MULTIPLY_BY_014_SELECTOR_OFFSET] = 0
    c0, c1 = tuple(x[:6]), tuple(x[6:])
    s, e = start_row, end_row
    t0 = fp6_multiply_by_01(c0, o0, o1)
    fill_trace_multiply_by_01(tr, c0, o0, o1, s, e, col + F12.MULTIPLY_BY_014_T0_CALC_OFFSET)
    t1 = fp6_multiply_by_1(c1, o4)
    fill_trace_multiply_by_1(tr, c1, o4, s, e, col + F12.MULTIPLY_BY_014_T1_CALC_OFFSET)
    t2 = fp6_mul_by_nonresidue(t1)
    nonres6_rows(tr, t1, s, e, col + F12.MULTIPLY_BY_014_T2_CALC_OFFSET)
    add_red6_rows(tr, t2, t0, s, e, col + F12.MULTIPLY_BY_014_X_CALC_OFFSET)
    t3 = fp6_add(c0, c1)
    add_red6_rows(tr, c0, c1, s, e, col + F12.MULTIPLY_BY_014_T3_CALC_OFFSET)
    t4 = fp2_add(o1, o4)
    add_red_rows(tr, o1, o4, s, e, col + F12.MULTIPLY_BY_014_T4_CALC_OFFSET)
    t5 = fp6_multiply_by_01(t3, o0, t4)
    fill_trace_multiply_by_01(tr, t3, o0, t4, s, e, col + F12.MULTIPLY_BY_014_T5_CALC_OFFSET)
    t6 = fp6_sub(t5, t0)
    sub_red6_rows(tr, t5, t0, s, e, col + F12.MULTIPLY_BY_014_T6_CALC_OFFSET)
    sub_red6_rows(tr, t6, t1, s, e, col + F12.MULTIPLY_BY_014_Y_CALC_OFFSET)


def fill_trace_fp12_multiplication(tr, x, y, start_row, end_row, col):            # fp12.rs:186-232
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F12.FP12_MUL_X_INPUT_OFFSET:col + F12.FP12_MUL_X_INPUT_OFFSET + 144] = _flat(x)
    tr[rows, col + F12.FP12_MUL_Y_INPUT_OFFSET:col + F12.FP12_MUL_Y_INPUT_OFFSET + 144] = _flat(y)
    tr[rows, col + F12.FP12_MUL_SELECTOR_OFFSET] = 1
    tr[end_row, col + F12.FP12_MUL_SELECTOR_OFFSET] = 0
    c0, c1, r0, r1 = tuple(x[:6]), tuple(x[6:]), tuple(y[:6]), tuple(y[6:])
    s, e = start_row, end_row
    t0 = fp6_mul(c0, r0); fill_trace_fp6_multiplication(tr, c0, r0, s, e, col + F12.FP12_MUL_T0_CALC_OFFSET)
    t1 = fp6_mul(c1, r1); fill_trace_fp6_multiplication(tr, c1, r1, s, e, col + F12.FP12_MUL_T1_CALC_OFFSET)
    t2 = fp6_mul_by_nonresidue(t1); nonres6_rows(tr, t1, s, e, col + F12.FP12_MUL_T2_CALC_OFFSET)
    add_red6_rows(tr, t0, t2, s, e, col + F12.FP12_MUL_X_CALC_OFFSET)
    t3 = fp6_add(c0, c1); add_red6_rows(tr, c0, c1, s, e, col + F12.FP12_MUL_T3_CALC_OFFSET)
    t4 = fp6_add(r0, r1); add_red6_rows(tr, r0, r1, s, e, col + F12.FP12_MUL_T4_CALC_OFFSET)
    t5 = fp6_mul(t3, t4); fill_trace_fp6_multiplication(tr, t3, t4, s, e, col + F12.FP12_MUL_T5_CALC_OFFSET)
    t6 = fp6_sub(t5, t0); sub_red6_rows(tr, t5, t0, s, e, col + F12.FP12_MUL_T6_CALC_OFFSET)
    sub_red6_rows(tr, t6, t1, s, e, col + F12.FP12_MUL_Y_CALC_OFFSET)


def fill_trace_cyclotomic_sq(tr, x, start_row, end_row, col):                     # fp12.rs:234-332
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F12.CYCLOTOMIC_SQ_INPUT_OFFSET:col + F12.CYCLOTOMIC_SQ_INPUT_OFFSET + 144] = _flat(x)
    tr[rows, col + F12.CYCLOTOMIC_SQ_SELECTOR_OFFSET] = 1
    tr[end_row, col + F12.CYCLOTOMIC_SQ_SELECTOR_OFFSET] = 0
    c0c0, c0c1, c0c2, c1c0, c1c1, c1c2 = [(x[2 * i], x[2 * i + 1]) for i in range(6)]
    s, e = start_row, end_row
    t0 = fp4_square(c0c0, c1c1); fill_trace_fp4_sq(tr, c0c0, c1c1, s, e, col + F12.CYCLOTOMIC_SQ_T0_CALC_OFFSET)
    t1 = fp4_square(c1c0, c0c2); fill_trace_fp4_sq(tr, c1c0, c0c2, s, e, col + F12.CYCLOTOMIC_SQ_T1_CALC_OFFSET)
    t2 = fp4_square(c0c1, c1c2); fill_trace_fp4_sq(tr, c0c1, c1c2, s, e, col + F12.CYCLOTOMIC_SQ_T2_CALC_OFFSET)
    t3 = fp2_mul_by_nonresidue(t2[1]); nonres_rows(tr, t2[1], s, e, col + F12.CYCLOTOMIC_SQ_T3_CALC_OFFSET)

    def sub_branch(a, b, o_t, o_2, o_c):        # t = a - b ; t' = 2 t ; c = t' + a
        t = fp2_sub(a, b); sub_red_rows(tr, a, b, s, e, col + o_t)
        t_2 = fp2_mul_fp(t, 2); fill_trace_fp2_fp_mul(tr, t, 2, s, e, col + o_2)
        add_red_rows(tr, t_2, a, s, e, col + o_c)

    def add_branch(a, b, o_t, o_2, o_c):        # t = a + b ; t' = 2 t ; c = t' + a
        t = fp2_add(a, b); add_red_rows(tr, a, b, s, e, col + o_t)
        t_2 = fp2_mul_fp(t, 2); fill_trace_fp2_fp_mul(tr, t, 2, s, e, col + o_2)
        add_red_rows(tr, t_2, a, s, e, col + o_c)

    sub_branch(t0[0], c0c0, F12.CYCLOTOMIC_SQ_T4_CALC_OFFSET, F12.CYCLOTOMIC_SQ_T5_CALC_OFFSET, F12.CYCLOTOMIC_SQ_C0_CALC_OFFSET)
    sub_branch(t1[0], c0c1, F12.CYCLOTOMIC_SQ_T6_CALC_OFFSET, F12.CYCLOTOMIC_SQ_T7_CALC_OFFSET, F12.CYCLOTOMIC_SQ_C1_CALC_OFFSET)
    sub_branch(t2[0], c0c2, F12.CYCLOTOMIC_SQ_T8_CALC_OFFSET, F12.CYCLOTOMIC_SQ_T9_CALC_OFFSET, F12.CYCLOTOMIC_SQ_C2_CALC_OFFSET)
    add_branch(t3, c1c0, F12.CYCLOTOMIC_SQ_T10_CALC_OFFSET, F12.CYCLOTOMIC_SQ_T11_CALC_OFFSET, F12.CYCLOTOMIC_SQ_C3_CALC_OFFSET)
    add_branch(t0[1], c1c1, F12.CYCLOTOMIC_SQ_T12_CALC_OFFSET, F12.CYCLOTOMIC_SQ_T13_CALC_OFFSET, F12.CYCLOTOMIC_SQ_C4_CALC_OFFSET)
    add_branch(t1[1], c1c2, F12.CYCLOTOMIC_SQ_T14_CALC_OFFSET, F12.CYCLOTOMIC_SQ_T15_CALC_OFFSET, F12.CYCLOTOMIC_SQ_C5_CALC_OFFSET)


def fill_trace_cyclotomic_exp(tr, x, start_row, end_row, col):                    # fp12.rs:335-375
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F12.INPUT_OFFSET:col + F12.INPUT_OFFSET + 144] = _flat(x)
    tr[rows, col + F12.CYCLOTOMIC_EXP_SELECTOR_OFFSET] = 1
    tr[end_row, col + F12.CYCLOTOMIC_EXP_SELECTOR_OFFSET] = 0
    tr[start_row, col + F12.CYCLOTOMIC_EXP_START_ROW] = 1
    z = (1,) + (0,) * 11
    i = BLS_X.bit_length() - 1
    bitone = False
    assert end_row + 1 - start_row == 70 * 12 + 1
    for j in range(70):
        s_row = start_row + j * 12
        e_row = s_row + 11
        blk = slice(s_row, e_row + 1)
        if bitone:
            tr[blk, col + F12.BIT1_SELECTOR_OFFSET] = 1
        tr[blk, col + F12.Z_OFFSET:col + F12.Z_OFFSET + 144] = _flat(z)
        tr[s_row, col + F12.FIRST_ROW_SELECTOR_OFFSET] = 1
        if bitone:
            fill_trace_fp12_multiplication(tr, z, x, s_row, e_row, col + F12.Z_MUL_INPUT_OFFSET)
            z = fp12_mul(z, x)
        else:
            fill_trace_cyclotomic_sq(tr, z, s_row, e_row, col + F12.Z_CYCLOTOMIC_SQ_OFFSET)
            z = fp12_cyclotomic_square(z)
        if ((BLS_X >> i) & 1) and not bitone:
            bitone = True
        elif j < 69:
            i -= 1
            bitone = False
    tr[start_row + 70 * 12, col + F12.RES_ROW_SELECTOR_OFFSET] = 1
    tr[start_row + 70 * 12, col + F12.Z_OFFSET:col + F12.Z_OFFSET + 144] = _flat(z)
    return z


def fill_trace_fp12_forbenius_map(tr, x, pw, start_row, end_row, col):            # fp12.rs:378-411
    div, rem = pw // 12, pw % 12
    rows = slice(start_row, end_row + 1)
    tr[rows, col + F12.FP12_FORBENIUS_MAP_INPUT_OFFSET:col + F12.FP12_FORBENIUS_MAP_INPUT_OFFSET + 144] = _flat(x)
    tr[rows, col + F12.FP12_FORBENIUS_MAP_SELECTOR_OFFSET] = 1
    tr[rows, col + F12.FP12_FORBENIUS_MAP_POW_OFFSET] = pw
    tr[rows, col + F12.FP12_FORBENIUS_MAP_DIV_OFFSET] = div
    tr[rows, col + F12.FP12_FORBENIUS_MAP_REM_OFFSET] = rem
    tr[rows, col + F12.FP12_FORBENIUS_MAP_BIT0_OFFSET] = rem & 1
    tr[rows, col + F12.FP12_FORBENIUS_MAP_BIT1_OFFSET] = (rem >> 1) & 1
    tr[rows, col + F12.FP12_FORBENIUS_MAP_BIT2_OFFSET] = (rem >> 2) & 1
    tr[rows, col + F12.FP12_FORBENIUS_MAP_BIT3_OFFSET] = rem >> 3
    tr[end_row, col + F12.FP12_FORBENIUS_MAP_SELECTOR_OFFSET] = 0
    r0, r1 = tuple(x[:6]), tuple(x[6:])
    s, e = start_row, end_row
    fill_trace_fp6_forbenius_map(tr, r0, pw, s, e, col + F12.FP12_FORBENIUS_MAP_R0_CALC_OFFSET)
    c0, c1, c2 = fp6_parts(fp6_frobenius(r1, pw))
    fill_trace_fp6_forbenius_map(tr, r1, pw, s, e, col + F12.FP12_FORBENIUS_MAP_C0C1C2_CALC_OFFSET)
    k = FP12_FROB[pw % 12]
    generate_trace_fp2_mul(tr, c0, k, s, e, col + F12.FP12_FORBENIUS_MAP_C0_CALC_OFFSET)
    generate_trace_fp2_mul(tr, c1, k, s, e, col + F12.FP12_FORBENIUS_MAP_C1_CALC_OFFSET)
    generate_trace_fp2_mul(tr, c2, k, s, e, col + F12.FP12_FORBENIUS_MAP_C2_CALC_OFFSET)


def fill_trace_fp12_conjugate(tr, x, row, col):                                   # fp12.rs:414-426
    put(tr, row, col + F12.FP12_CONJUGATE_INPUT_OFFSET, _flat(x))
    conj = tuple(x[:6]) + tuple(fp_neg(a) for a in x[6:])
    put(tr, row, col + F12.FP12_CONJUGATE_OUTPUT_OFFSET, _flat(conj))
    fill_trace_addition_fp6(tr, tuple(x[6:]), tuple(conj[6:]), row, col + F12.FP12_CONJUGATE_ADDITIION_OFFSET)
    return conj
