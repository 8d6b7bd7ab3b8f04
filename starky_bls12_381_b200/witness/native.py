"""BLS12-381 tower arithmetic as the reference's witness generators use it (restated from
/root/reference/src/native.rs; every function cites the lines it follows).

Representation: Fp = Python int (the value of the reference's [u32; 12], little-endian limbs), Fp2 / Fp6 / Fp12 = tuples
of 2 / 6 / 12 ints in the reference's coefficient order.  The reference's quirks are kept because they show up in the
trace: `-x` is `p - x` (so `-0 == p`, native.rs:417-424), `add_fp` subtracts p at most once (native.rs:452-468).
"""
import json
import os

P = 4002409555221667393417789825735904156556882819939007885332058136124031650490837864442687629129015664037894272559787  # native.rs:12-14
BLS_X = 15132376222941642752                                                                                           # native.rs:20-22
M32 = 0xFFFFFFFF

_here = os.path.dirname(os.path.abspath(__file__))
_tab = json.load(open(os.path.join(_here, "offsets.json")))
OFFSETS = _tab["offsets"]


class NS:
    """Offsets of one reference file as attributes: NS('fp').X_INPUT_OFFSET."""

    def __init__(self, file):
        pre = file + "."
        for k, v in OFFSETS.items():
            if k.startswith(pre):
                setattr(self, k[len(pre):], v)


def limbs(x, n=12):
    """get_u32_vec_from_literal / _24 (native.rs:233-240, 261-267)."""
    assert 0 <= x < (1 << (32 * n)), "value does not fit %d limbs" % n
    return [(x >> (32 * i)) & M32 for i in range(n)]


# ---- Fp (native.rs:345-520) ----
def fp_add(x, y):
    s = x + y
    return s if s < P else s - P


def fp_sub(x, y):
    return (P + x - y) % P


def fp_mul(x, y):
    return (x * y) % P


def fp_neg(x):
    return P - x


def fp_inv(x):
    return pow(x, -1, P)


# ---- Fp2 (native.rs:522-716) ----
def fp2_add(x, y): return (fp_add(x[0], y[0]), fp_add(x[1], y[1]))
def fp2_sub(x, y): return (fp_sub(x[0], y[0]), fp_sub(x[1], y[1]))
def fp2_neg(x): return (fp_neg(x[0]), fp_neg(x[1]))


def fp2_mul(x, y):
    return (fp_sub(fp_mul(x[0], y[0]), fp_mul(x[1], y[1])), fp_add(fp_mul(x[0], y[1]), fp_mul(x[1], y[0])))


def fp2_mul_fp(x, k): return (fp_mul(x[0], k), fp_mul(x[1], k))
def fp2_mul_by_nonresidue(x): return (fp_sub(x[0], x[1]), fp_add(x[0], x[1]))


def fp2_multiply_by_b(x):
    t0, t1 = fp_mul(x[0], 4), fp_mul(x[1], 4)
    return (fp_sub(t0, t1), fp_add(t0, t1))


def fp2_inv(x):
    re, im = x
    f = fp_inv(fp_add(fp_mul(re, re), fp_mul(im, im)))
    return (fp_mul(f, re), fp_mul(f, fp_neg(im)))


FP2_FROB = [int(v) for v in _tab["tables"]["Fp2.forbenius_coefficients"]]


def _pairs(key):
    v = [int(x) for x in _tab["tables"][key]]
    return [(v[2 * i], v[2 * i + 1]) for i in range(len(v) // 2)]


FP6_FROB_1 = _pairs("Fp6.forbenius_coefficients_1")
FP6_FROB_2 = _pairs("Fp6.forbenius_coefficients_2")
FP12_FROB = _pairs("Fp12.forbenius_coefficients")


def fp2_frobenius(x, pw):        # native.rs:1058-1064
    return (x[0], fp_mul(x[1], FP2_FROB[pw % 2]))


# ---- Fp6 (native.rs:717-921) ----
def fp6_parts(x): return (x[0], x[1]), (x[2], x[3]), (x[4], x[5])
def fp6_add(x, y): return tuple(fp_add(a, b) for a, b in zip(x, y))
def fp6_sub(x, y): return tuple(fp_sub(a, b) for a, b in zip(x, y))
def fp6_neg(x): return tuple(fp_neg(a) for a in x)


def fp6_mul(x, y):
    c0, c1, c2 = fp6_parts(x)
    r0, r1, r2 = fp6_parts(y)
    t0, t1, t2 = fp2_mul(c0, r0), fp2_mul(c1, r1), fp2_mul(c2, r2)
    t5 = fp2_mul(fp2_add(c1, c2), fp2_add(r1, r2))
    t8 = fp2_mul_by_nonresidue(fp2_sub(fp2_sub(t5, t1), t2))
    xx = fp2_add(t8, t0)
    t11 = fp2_mul(fp2_add(c0, c1), fp2_add(r0, r1))
    yy = fp2_add(fp2_sub(fp2_sub(t11, t0), t1), fp2_mul_by_nonresidue(t2))
    t17 = fp2_mul(fp2_add(c0, c2), fp2_add(r0, r2))
    zz = fp2_add(fp2_sub(fp2_sub(t17, t0), t2), t1)
    return xx + yy + zz


def fp6_mul_by_nonresidue(x):    # native.rs:863-873
    c0 = fp2_mul_by_nonresidue((x[4], x[5]))
    return (c0[0], c0[1], x[0], x[1], x[2], x[3])


def fp6_multiply_by_01(x, b0, b1):
    c0, c1, c2 = fp6_parts(x)
    t0, t1 = fp2_mul(c0, b0), fp2_mul(c1, b1)
    xx = fp2_add(fp2_mul_by_nonresidue(fp2_mul(c2, b1)), t0)
    t6 = fp2_mul(fp2_add(b0, b1), fp2_add(c0, c1))
    yy = fp2_sub(fp2_sub(t6, t0), t1)
    zz = fp2_add(fp2_mul(c2, b0), t1)
    return xx + yy + zz


def fp6_multiply_by_1(x, b1):
    c0, c1, c2 = fp6_parts(x)
    return fp2_mul_by_nonresidue(fp2_mul(c2, b1)) + fp2_mul(c0, b1) + fp2_mul(c1, b1)


def fp6_inv(x):                  # native.rs:720-734
    c0, c1, c2 = fp6_parts(x)
    t0 = fp2_sub(fp2_mul(c0, c0), fp2_mul_by_nonresidue(fp2_mul(c2, c1)))
    t1 = fp2_sub(fp2_mul_by_nonresidue(fp2_mul(c2, c2)), fp2_mul(c0, c1))
    t2 = fp2_sub(fp2_mul(c1, c1), fp2_mul(c0, c2))
    t4 = fp2_inv(fp2_add(fp2_mul_by_nonresidue(fp2_add(fp2_mul(c2, t1), fp2_mul(c1, t2))), fp2_mul(c0, t0)))
    return fp2_mul(t4, t0) + fp2_mul(t4, t1) + fp2_mul(t4, t2)


def fp6_frobenius(x, pw):        # native.rs:1126-1145
    c0, c1, c2 = fp6_parts(x)
    return (fp2_frobenius(c0, pw) + fp2_mul(fp2_frobenius(c1, pw), FP6_FROB_1[pw % 6])
            + fp2_mul(fp2_frobenius(c2, pw), FP6_FROB_2[pw % 6]))


# ---- Fp12 (native.rs:921-1475) ----
FP12_ONE = (1,) + (0,) * 11


def fp12_mul(x, y):              # native.rs:1009-1027
    c0, c1, r0, r1 = x[:6], x[6:], y[:6], y[6:]
    t0, t1 = fp6_mul(c0, r0), fp6_mul(c1, r1)
    xx = fp6_add(t0, fp6_mul_by_nonresidue(t1))
    t5 = fp6_mul(fp6_add(c0, c1), fp6_add(r0, r1))
    return xx + fp6_sub(fp6_sub(t5, t0), t1)


def fp12_inv(x):                 # native.rs:932-940
    c0, c1 = x[:6], x[6:]
    t = fp6_inv(fp6_sub(fp6_mul(c0, c0), fp6_mul_by_nonresidue(fp6_mul(c1, c1))))
    return fp6_mul(c0, t) + fp6_neg(fp6_mul(c1, t))


def fp12_multiply_by_014(x, o0, o1, o4):     # native.rs:1228-1244
    c0, c1 = x[:6], x[6:]
    t0 = fp6_multiply_by_01(c0, o0, o1)
    t1 = fp6_multiply_by_1(c1, o4)
    xx = fp6_add(fp6_mul_by_nonresidue(t1), t0)
    t5 = fp6_multiply_by_01(fp6_add(c1, c0), o0, fp2_add(o1, o4))
    return xx + fp6_sub(fp6_sub(t5, t0), t1)


def fp12_conjugate(x):           # native.rs:1246-1252
    return tuple(x[:6]) + tuple(fp_neg(a) for a in x[6:])


def fp12_frobenius(x, pw):       # native.rs:1202-1224
    r0 = fp6_frobenius(x[:6], pw)
    c0, c1, c2 = fp6_parts(fp6_frobenius(x[6:], pw))
    k = FP12_FROB[pw % 12]
    return r0 + fp2_mul(c0, k) + fp2_mul(c1, k) + fp2_mul(c2, k)


def fp4_square(a, b):            # native.rs:224-231
    a2, b2 = fp2_mul(a, a), fp2_mul(b, b)
    s = fp2_add(a, b)
    return fp2_add(fp2_mul_by_nonresidue(b2), a2), fp2_sub(fp2_sub(fp2_mul(s, s), a2), b2)


def fp12_cyclotomic_square(x):   # native.rs:1254-1298
    c0c0, c0c1, c0c2, c1c0, c1c1, c1c2 = [(x[2 * i], x[2 * i + 1]) for i in range(6)]
    t0, t1, t2 = fp4_square(c0c0, c1c1), fp4_square(c1c0, c0c2), fp4_square(c0c1, c1c2)
    t3 = fp2_mul_by_nonresidue(t2[1])
    two = lambda v: fp2_mul_fp(v, 2)
    c0 = fp2_add(two(fp2_sub(t0[0], c0c0)), t0[0])
    c1 = fp2_add(two(fp2_sub(t1[0], c0c1)), t1[0])
    c2 = fp2_add(two(fp2_sub(t2[0], c0c2)), t2[0])
    c3 = fp2_add(two(fp2_add(t3, c1c0)), t3)
    c4 = fp2_add(two(fp2_add(t0[1], c1c1)), t0[1])
    c5 = fp2_add(two(fp2_add(t1[1], c1c2)), t1[1])
    return c0 + c1 + c2 + c3 + c4 + c5


def fp12_cyclotomic_exponent(x):  # native.rs:1300-1309
    z = FP12_ONE
    for i in reversed(range(BLS_X.bit_length())):
        z = fp12_cyclotomic_square(z)
        if (BLS_X >> i) & 1:
            z = fp12_mul(z, x)
    return z


def fp12_final_exponentiate(x):  # native.rs:1311-1345
    t0 = fp12_frobenius(x, 6)
    t1 = fp12_mul(t0, fp12_inv(x))
    t2 = fp12_frobenius(t1, 2)
    t3 = fp12_mul(t2, t1)
    t4 = fp12_cyclotomic_exponent(t3)
    t5 = fp12_conjugate(t4)
    t6 = fp12_cyclotomic_square(t3)
    t7 = fp12_conjugate(t6)
    t8 = fp12_mul(t7, t5)
    t9 = fp12_cyclotomic_exponent(t8)
    t10 = fp12_conjugate(t9)
    t11 = fp12_cyclotomic_exponent(t10)
    t12 = fp12_conjugate(t11)
    t13 = fp12_cyclotomic_exponent(t12)
    t14 = fp12_conjugate(t13)
    t15 = fp12_cyclotomic_square(t5)
    t16 = fp12_mul(t14, t15)
    t17 = fp12_cyclotomic_exponent(t16)
    t18 = fp12_conjugate(t17)
    t19 = fp12_mul(t5, t12)
    t20 = fp12_frobenius(t19, 2)
    t21 = fp12_mul(t10, t3)
    t22 = fp12_frobenius(t21, 3)
    t23 = fp12_conjugate(t3)
    t24 = fp12_mul(t16, t23)
    t25 = fp12_frobenius(t24, 1)
    t26 = fp12_conjugate(t8)
    t27 = fp12_mul(t18, t26)
    t28 = fp12_mul(t27, t3)
    t29 = fp12_mul(t20, t22)
    t30 = fp12_mul(t29, t25)
    return fp12_mul(t30, t28)


HALF = pow(2, -1, P)


def calc_precomp_stuff_loop0(rx, ry, rz):   # native.rs:291-327 (same values as one doubling step of calc_pairing_precomp)
    t0 = fp2_mul(ry, ry)
    t1 = fp2_mul(rz, rz)
    x0 = fp2_mul_fp(t1, 3)
    t2 = fp2_multiply_by_b(x0)
    t3 = fp2_mul_fp(t2, 3)
    x1 = fp2_mul(ry, rz)
    t4 = fp2_mul_fp(x1, 2)
    x2 = fp2_sub(t2, t0)
    x3 = fp2_mul(rx, rx)
    x4 = fp2_mul_fp(x3, 3)
    x5 = fp2_neg(t4)
    x6 = fp2_sub(t0, t3)
    x7 = fp2_mul(rx, ry)
    x8 = fp2_mul(x6, x7)
    x9 = fp2_add(t0, t3)
    x10 = fp2_mul_fp(x9, HALF)
    x11 = fp2_mul(x10, x10)
    x12 = fp2_mul(t2, t2)
    x13 = fp2_mul_fp(x12, 3)
    new_rx = fp2_mul_fp(x8, HALF)
    new_ry = fp2_sub(x11, x13)
    new_rz = fp2_mul(t0, t4)
    return [new_rx, new_ry, new_rz, t0, t1, x0, t2, t3, x1, t4, x3, x2, x4, x5, x6, x7, x8, x9, x10, x11, x12, x13]


def calc_precomp_stuff_loop1(rx, ry, rz, qx, qy):   # native.rs:329-372
    t0 = fp2_mul(qy, rz)
    t1 = fp2_sub(ry, t0)
    t2 = fp2_mul(qx, rz)
    t3 = fp2_sub(rx, t2)
    t4 = fp2_mul(t1, qx)
    t5 = fp2_mul(t3, qy)
    t6 = fp2_sub(t4, t5)
    t7 = fp2_neg(t1)
    t8 = fp2_mul(t3, t3)
    t9 = fp2_mul(t8, t3)
    t10 = fp2_mul(t8, rx)
    t11 = fp2_mul(t1, t1)
    t12 = fp2_mul(t11, rz)
    t13 = fp2_mul_fp(t10, 2)
    t14 = fp2_sub(t9, t13)
    t15 = fp2_add(t14, t12)
    t16 = fp2_sub(t10, t15)
    t17 = fp2_mul(t16, t1)
    t18 = fp2_mul(t9, ry)
    new_rx = fp2_mul(t3, t15)
    new_ry = fp2_sub(t17, t18)
    new_rz = fp2_mul(rz, t9)
    return [new_rx, new_ry, new_rz, t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18]


def calc_pairing_precomp(x, y, z):   # native.rs:1358-1437
    zi = fp2_inv(z)
    qx, qy = fp2_mul(x, zi), fp2_mul(y, zi)
    rx, ry, rz = qx, qy, (1, 0)
    ell = []
    for i in reversed(range(BLS_X.bit_length() - 1)):
        v = calc_precomp_stuff_loop0(rx, ry, rz)
        ell.append((v[11], v[12], v[13]))       # x2, x4, x5
        rx, ry, rz = v[0], v[1], v[2]
        if (BLS_X >> i) & 1:
            w = calc_precomp_stuff_loop1(rx, ry, rz, qx, qy)
            ell.append((w[9], w[10], w[6]))     # bit1_t6, bit1_t7, bit1_t3
            rx, ry, rz = w[0], w[1], w[2]
    return ell


def miller_loop(px, py, qx, qy, qz):   # native.rs:1440-1466
    pre = calc_pairing_precomp(qx, qy, qz)
    f12, j = FP12_ONE, 0
    for i in reversed(range(BLS_X.bit_length() - 1)):
        e = pre[j]
        f12 = fp12_multiply_by_014(f12, e[0], fp2_mul_fp(e[1], px), fp2_mul_fp(e[2], py))
        if (BLS_X >> i) & 1:
            j += 1
            e = pre[j]
            f12 = fp12_multiply_by_014(f12, e[0], fp2_mul_fp(e[1], px), fp2_mul_fp(e[2], py))
        if i != 0:
            f12 = fp12_mul(f12, f12)
        j += 1
    return fp12_conjugate(f12)
