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
