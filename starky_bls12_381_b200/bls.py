"""Input preparation for the seven proofs of one BLS signature verification, as the reference does it before it calls
its *_main functions (/root/reference/src/main.rs:8-55 and aggregate_proof.rs:244-340, which lean on
snowbridge-milagro-bls and eth-types): public keys and signature decompressed to affine coordinates, the public keys of
the set sync-committee bits aggregated, the signing root of the attested header, and the message hashed to G2
(BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_, RFC 9380 section 8.8.2 -- the same simple-SWU / 3-isogeny / psi cofactor
clearing the reference restates as a circuit in hash_to_curve.rs:84-350).  Host side, Python integers; the isogeny
coefficients are the reference's table (witness/iso_g2.json, written by tests/golden/make_bundled_fixture.py).
Everything here is checked end to end by the pairing equation: with these inputs the final exponentiation of the two
Miller-loop outputs is ONE (tests/test_bundled.py)."""
import hashlib
import json
import os

from .witness import native as N

P = N.P
BLS_X_ABS = N.BLS_X                       # |x| = 0xd201000000010000; the curve parameter itself is negative
DST = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"           # aggregate_proof.rs:234
# the constant second G1 point of aggregate_proof.rs:320-321: the negated generator, so that the product of the two Miller
# loops is e(apk, H(m)) / e(g1, sig)
NEG_G1 = (3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
          2662903010277190920397318445793982934971948944000658264905514399707520226534504357969962973775649129045502516118218)
_ISO = None


def iso_coefficients():
    global _ISO
    if _ISO is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "witness", "iso_g2.json")
        _ISO = [[(int(c[0]), int(c[1])) for c in poly] for poly in json.load(open(path))["ISOGENY_COEFFICIENTS_G2"]]
    return _ISO


# ---------------------------------------------------------------------------------------------------- Fp / Fp2 helpers
def fp_sqrt(a):
    """p = 3 (mod 4)."""
    r = pow(a, (P + 1) // 4, P)
    return r if r * r % P == a % P else None


def f2_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def f2_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def f2_neg(x): return ((-x[0]) % P, (-x[1]) % P)
def f2_mul(x, y): return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def f2_sqr(x): return f2_mul(x, x)
def f2_conj(x): return (x[0], (-x[1]) % P)


def f2_inv(x):
    d = pow((x[0] * x[0] + x[1] * x[1]) % P, -1, P)
    return (x[0] * d % P, (-x[1]) * d % P)


def f2_pow(x, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, x)
        x = f2_sqr(x)
        e >>= 1
    return r


def f2_sqrt(a):
    """Square root in Fp[i]/(i^2 + 1) by the complex method; None if `a` is not a square."""
    if a == (0, 0):
        return (0, 0)
    s = fp_sqrt((a[0] * a[0] + a[1] * a[1]) % P)
    if s is None:
        return None
    half = pow(2, -1, P)
    for cand in ((a[0] + s) * half % P, (a[0] - s) * half % P):
        x0 = fp_sqrt(cand)
        if x0 is None or x0 == 0:
            continue
        x1 = a[1] * pow(2 * x0, -1, P) % P
        if f2_sqr((x0, x1)) == (a[0] % P, a[1] % P):
            return (x0, x1)
    if a[1] % P == 0:                      # a = -c^2 with c in Fp: root is purely imaginary
        x1 = fp_sqrt((-a[0]) % P)
        if x1 is not None:
            return (0, x1)
    return None


def sgn0_f2(x):
    """RFC 9380 section 4.1 for m = 2."""
    s0, z0, s1 = x[0] & 1, x[0] == 0, x[1] & 1
    return s0 | (z0 & s1)


# ---------------------------------------------------------------------------------------------------- affine curve arithmetic
class Curve:
    """y^2 = x^3 + b over a field given by (add, sub, mul, inv); None = the point at infinity."""

    def __init__(self, add, sub, mul, inv, zero):
        self.fa, self.fs, self.fm, self.fi, self.zero = add, sub, mul, inv, zero

    def neg(self, p):
        return None if p is None else (p[0], self.fs(self.zero, p[1]))

    def add(self, p, q):
        if p is None:
            return q
        if q is None:
            return p
        fa, fs, fm, fi = self.fa, self.fs, self.fm, self.fi
        if p[0] == q[0]:
            if fa(p[1], q[1]) == self.zero:
                return None
            x2 = fm(p[0], p[0])
            lam = fm(fa(fa(x2, x2), x2), fi(fa(p[1], p[1])))
        else:
            lam = fm(fs(q[1], p[1]), fi(fs(q[0], p[0])))
        x3 = fs(fs(fm(lam, lam), p[0]), q[0])
        return (x3, fs(fm(lam, fs(p[0], x3)), p[1]))

    def mul(self, p, k):
        r, q = None, p
        while k:
            if k & 1:
                r = self.add(r, q)
            q = self.add(q, q)
            k >>= 1
        return r


G1 = Curve(lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a, b: a * b % P, lambda a: pow(a, -1, P), 0)
G2 = Curve(f2_add, f2_sub, f2_mul, f2_inv, (0, 0))


# ---------------------------------------------------------------------------------------------------- (de)compression
def g1_decompress(data):
    """48-byte compressed G1 point (ZCash format) -> affine (x, y) as ints; what PublicKey::from_bytes + getx/gety give
    the reference (aggregate_proof.rs:247-255)."""
    assert len(data) == 48 and data[0] & 0x80, "compressed form expected"
    if data[0] & 0x40:
        return None
    x = int.from_bytes(data, "big") & ((1 << 381) - 1)
    y = fp_sqrt((x * x * x + 4) % P)
    assert y is not None, "not on the curve"
    if (y > (P - 1) // 2) != bool(data[0] & 0x20):
        y = P - y
    return (x, y)


def g2_decompress(data):
    """96-byte compressed G2 point: x.c1 (with the flag bits) then x.c0 -> affine ((x0, x1), (y0, y1)); the reference's
    Signature::from_bytes + getx().geta()/getb() (aggregate_proof.rs:323-333)."""
    assert len(data) == 96 and data[0] & 0x80, "compressed form expected"
    x1 = int.from_bytes(data[:48], "big") & ((1 << 381) - 1)
    x0 = int.from_bytes(data[48:], "big")
    x = (x0, x1)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), (4, 4)))
    assert y is not None, "not on the curve"
    largest = y[1] > (P - 1) // 2 or (y[1] == 0 and y[0] > (P - 1) // 2)
    if largest != bool(data[0] & 0x20):
        y = f2_neg(y)
    return (x, y)


# ---------------------------------------------------------------------------------------------------- hash to G2
def expand_message_xmd(msg, dst, n):
    """RFC 9380 section 5.3.1 with SHA-256."""
    ell = (n + 31) // 32
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(64) + msg + n.to_bytes(2, "big") + b"\0" + dst_prime).digest()
    b = [hashlib.sha256(b0 + b"\x01" + dst_prime).digest()]
    for i in range(2, ell + 1):
        b.append(hashlib.sha256(bytes(x ^ y for x, y in zip(b0, b[-1])) + bytes([i]) + dst_prime).digest())
    return b"".join(b)[:n]


def hash_to_field_fp2(msg, dst, count=2):
    """RFC 9380 section 5.2: m = 2, L = 64 (hash_to_field.rs restates it as a circuit)."""
    u = expand_message_xmd(msg, dst, count * 2 * 64)
    return [tuple(int.from_bytes(u[64 * (j + 2 * i):64 * (j + 2 * i) + 64], "big") % P for j in range(2)) for i in range(count)]


ISO_A, ISO_B, ISO_Z = (0, 240), (1012, 1012), (P - 2, P - 1)      # hash_to_curve.rs:92-103


def map_to_curve_simple_swu(u):
    """RFC 9380 section 6.6.2 on E2': y^2 = x^3 + 240 i x + 1012 (1 + i)  (hash_to_curve.rs:84-202)."""
    tv1 = f2_mul(ISO_Z, f2_sqr(u))
    tv2 = f2_add(f2_sqr(tv1), tv1)
    if tv2 == (0, 0):
        x1 = f2_mul(ISO_B, f2_inv(f2_mul(ISO_Z, ISO_A)))
    else:
        x1 = f2_mul(f2_mul(f2_neg(ISO_B), f2_inv(ISO_A)), f2_add((1, 0), f2_inv(tv2)))
    g = lambda x: f2_add(f2_add(f2_mul(f2_sqr(x), x), f2_mul(ISO_A, x)), ISO_B)
    gx1 = g(x1)
    y = f2_sqrt(gx1)
    x = x1
    if y is None:
        x = f2_mul(tv1, x1)
        y = f2_sqrt(g(x))
        assert y is not None
    if sgn0_f2(u) != sgn0_f2(y):
        y = f2_neg(y)
    return (x, y)


def iso_map(pt):
    """The 3-isogeny E2' -> E2 (hash_to_curve.rs:203-249; coefficient table :9-82, highest power first)."""
    k = iso_coefficients()
    x, y = pt
    x2 = f2_sqr(x)
    x3 = f2_mul(x2, x)
    poly = lambda c, lead: f2_add(f2_add(f2_add(c[3], f2_mul(c[2], x)), f2_mul(c[1], x2)), f2_mul(lead, x3))
    x_num = poly(k[0], k[0][0])
    x_den = f2_add(f2_add(k[1][3], f2_mul(k[1][2], x)), x2)
    y_num = poly(k[2], k[2][0])
    y_den = f2_add(f2_add(f2_add(k[3][3], f2_mul(k[3][2], x)), f2_mul(k[3][1], x2)), x3)
    return (f2_mul(x_num, f2_inv(x_den)), f2_mul(y, f2_mul(y_num, f2_inv(y_den))))


PSI_CX = f2_inv(f2_pow((1, 1), (P - 1) // 3))
PSI_CY = f2_inv(f2_pow((1, 1), (P - 1) // 2))
PSI2_CX = pow(pow(2, (P - 1) // 3, P), -1, P)


def psi(pt):
    return None if pt is None else (f2_mul(PSI_CX, f2_conj(pt[0])), f2_mul(PSI_CY, f2_conj(pt[1])))


def psi2(pt):
    return None if pt is None else ((pt[0][0] * PSI2_CX % P, pt[0][1] * PSI2_CX % P), f2_neg(pt[1]))


def clear_cofactor_g2(pt):
    """RFC 9380 appendix G.4 (Budroni-Pintore), c1 = x = -|x|  (hash_to_curve.rs:290-319)."""
    mul_x = lambda q: G2.neg(G2.mul(q, BLS_X_ABS))
    t1 = mul_x(pt)
    t2 = psi(pt)
    t3 = psi2(G2.add(pt, pt))
    t3 = G2.add(t3, G2.neg(t2))
    t2 = mul_x(G2.add(t1, t2))
    t3 = G2.add(t3, t2)
    t3 = G2.add(t3, G2.neg(t1))
    return G2.add(t3, G2.neg(pt))


def hash_to_curve_g2(msg, dst=DST):
    """aggregate_proof.rs:290: hash_to_curve_g2(&signing_root, &dst) -> affine ((x0, x1), (y0, y1))."""
    u0, u1 = hash_to_field_fp2(msg, dst)
    q = G2.add(iso_map(map_to_curve_simple_swu(u0)), iso_map(map_to_curve_simple_swu(u1)))
    return clear_cofactor_g2(q)


# ---------------------------------------------------------------------------------------------------- SSZ signing root
def _merkleize(chunks):
    n = 1
    while n < len(chunks):
        n *= 2
    layer = list(chunks) + [bytes(32)] * (n - len(chunks))
    while len(layer) > 1:
        layer = [hashlib.sha256(layer[i] + layer[i + 1]).digest() for i in range(0, len(layer), 2)]
    return layer[0]


def signing_root(header, domain):
    """main.rs:33-42: SigningData { object_root: hash_tree_root(BeaconBlockHeader), domain }.tree_hash_root()."""
    hx = lambda s: bytes.fromhex(s[2:] if s.startswith("0x") else s)
    root = _merkleize([int(header["slot"]).to_bytes(8, "little") + bytes(24), int(header["proposer_index"]).to_bytes(8, "little") + bytes(24),
                       hx(header["parent_root"]), hx(header["state_root"]), hx(header["body_root"])])
    return _merkleize([root, domain])


def prepare(pubkeys_hex, bits_hex, signature_hex, root):
    """Everything generate_aggregate_proof derives before its seven *_main calls (aggregate_proof.rs:244-340):
    points (512 affine G1), bits, apk, the hashed message Q1 and the signature Q2 (affine G2)."""
    hx = lambda s: bytes.fromhex(s[2:] if s.startswith("0x") else s)
    points = [g1_decompress(hx(k)) for k in pubkeys_hex]
    raw = hx(bits_hex)
    bits = [bool((raw[i // 8] >> (i % 8)) & 1) for i in range(8 * len(raw))]
    apk = None
    for pt, b in zip(points, bits):
        if b:
            apk = G1.add(apk, pt)
    return dict(points=points, bits=bits, apk=apk, q1=hash_to_curve_g2(root), q2=g2_decompress(hx(signature_hex)), signing_root=root)


def pairing_product_is_one(inp):
    """The statement the seven proofs establish: FinalExp(MillerLoop(apk, H(m)) * MillerLoop(-g1, sig)) == 1."""
    one = (1, 0)
    ml1 = N.miller_loop(inp["apk"][0], inp["apk"][1], inp["q1"][0], inp["q1"][1], one)
    ml2 = N.miller_loop(NEG_G1[0], NEG_G1[1], inp["q2"][0], inp["q2"][1], one)
    return N.fp12_final_exponentiate(N.fp12_mul(ml1, ml2)) == N.FP12_ONE
