"""Constraint-program compiler: flat constraint DAG (reference emission order) -> grouped bytecode for the GPU
quotient evaluator (starky_bls12_381_b200/csrc/quotient.cu).

The reference's ConstraintConsumer folds constraints with Horner, acc <- acc*alpha + f_k*c_k; in a
field that equals sum_k alpha^(K-1-k) * f_k * c_k, so the evaluation order is free.  The compiler uses that freedom:

  * every constraint polynomial is split at its top-level product into selector factors and one body;
  * constraints with the same (class, selector factors) form a GROUP (first-appearance order): the body sums
    G_j = sum_k w_kj * body_k are accumulated unreduced and multiplied by the selector product and the class
    factor (1, z_last, L_first, L_last) once per group;
  * identical bodies inside a group share one evaluation: their weights are summed into one weight SLOT
    (the reference's range check emits the same two constraints 12 times, fp.rs:1346-1377);
  * bodies are normalised to sums of monomials and emitted as one 64-bit instruction per term.

Instruction word (64 bit):  op[0:4] | end[4] | neg0[5] | neg1[6] | v0[8:26] | v1[26:44] | v2[44:62]
  variables: 0..C-1 local column, C..2C-1 next column, 2C.. public input.
  ADD1  T +-= v0                    ADD2  T +-= v0 ; T +-= v1 (neg0, neg1)
  SHL1  T +-= v0 * 2^32             MULS  T +-= imm32 * v0            (imm32 in bits 26..58)
  MUL2  T +-= v0 * v1               MULC1 T +-= const[v2] * v0
  MULC2 T +-= const[v2] * v0 * v1   MUL3C T +-= const[next word] * v0 * v1 * v2   (two words)
  CONSTI T +-= imm32                CONSTC T +-= const[v2]
  GROUP cls = v0, n_factors = v1    (then n_factors selector-factor polynomials, each closed by `end`,
                                     then the bodies, each closed by `end` -> consumes the next weight slot)
"""
import struct

from .poly import Expander
from .rustsym import LOCAL, NEXT, P, PI

OP_NOP, OP_ADD1, OP_ADD2, OP_SHL1, OP_MULS, OP_MUL2, OP_MULC1, OP_MULC2, OP_MUL3C, OP_CONSTI, OP_CONSTC, OP_GROUP = range(12)
MAX_GROUP = 256           # constraints per group segment (giant groups are split so chunks can balance)
TWO32 = 1 << 32


class Program:
    def __init__(self):
        self.code = []          # u64 words
        self.consts = []        # u64
        self.const_index = {}
        self.slot_off = [0]     # CSR: slot -> constraint indices
        self.slot_ks = []
        self.group_pc = []      # pc of every GROUP instruction
        self.group_slot = []    # first weight slot of every group
        self.n_cols = self.n_pis = self.K = self.degree = 0

    def cidx(self, v):
        i = self.const_index.get(v)
        if i is None:
            i = len(self.consts)
            self.consts.append(v)
            self.const_index[v] = i
        return i


def ins(op, end=0, neg0=0, neg1=0, v0=0, v1=0, v2=0):
    assert v0 < (1 << 18) and v1 < (1 << 18) and v2 < (1 << 18)
    return op | (end << 4) | (neg0 << 5) | (neg1 << 6) | (v0 << 8) | (v1 << 26) | (v2 << 44)


def ins_imm(op, end, neg0, v0, imm):
    assert 0 <= imm < TWO32
    return op | (end << 4) | (neg0 << 5) | (v0 << 8) | (imm << 26)


def compile_program(dag, constraints, n_cols, n_pis, degree):
    ex = Expander(dag)
    prog = Program()
    prog.n_cols, prog.n_pis, prog.K, prog.degree = n_cols, n_pis, len(constraints), degree

    def var(nid):
        op, a, _ = dag.nodes[nid]
        return a if op == LOCAL else (n_cols + a if op == NEXT else 2 * n_cols + a)

    # ---- grouping ----
    groups, order = {}, []
    for k, (cls, nid) in enumerate(constraints):
        f = ex.top_factors(nid)
        if not f:
            body, sel = None, ()
        else:
            sizes = [ex.size(x) for x in f]
            bi = max(range(len(f)), key=lambda i: (sizes[i], i))
            body, sel = f[bi], tuple(sorted(f[:bi] + f[bi + 1:]))
        key = (cls, sel)
        g = groups.get(key)
        if g is None:
            g = groups[key] = {}
            order.append(key)
        poly = ex.poly(body) if body is not None else {(): 1}
        pkey = body if body is not None else -1
        # identical body node (hash-consed) -> same slot; fall back to polynomial identity
        ent = g.get(pkey)
        if ent is None:
            g[pkey] = ent = (poly, [])
        ent[1].append(k)

    def emit_poly(poly, words):
        """Emit the terms of one polynomial; the last word (always a one-word instruction) carries `end`."""
        two_word, single, lin = [], [], []      # lin: (neg, var) coefficient +-1 degree-1 terms, paired into ADD2
        for mono, c in sorted(poly.items(), key=lambda mc: (len(mc[0]), mc[0])):
            neg, mag = 0, c
            if P - c < c and (P - c) <= TWO32:      # prefer small magnitudes with a sign
                neg, mag = 1, P - c
            vs = [var(x) for x in mono]
            d = len(vs)
            if d == 0:
                if mag < TWO32: single.append(ins_imm(OP_CONSTI, 0, neg, 0, mag))
                else: single.append(ins(OP_CONSTC, neg0=neg, v2=prog.cidx(mag)))
            elif d == 1:
                if mag == 1: lin.append((neg, vs[0]))
                elif mag == TWO32: single.append(ins(OP_SHL1, neg0=neg, v0=vs[0]))
                elif mag < TWO32: single.append(ins_imm(OP_MULS, 0, neg, vs[0], mag))
                else: single.append(ins(OP_MULC1, neg0=neg, v0=vs[0], v2=prog.cidx(mag)))
            elif d == 2:
                if mag == 1: single.append(ins(OP_MUL2, neg0=neg, v0=vs[0], v1=vs[1]))
                else: single.append(ins(OP_MULC2, neg0=neg, v0=vs[0], v1=vs[1], v2=prog.cidx(mag)))
            elif d == 3:
                two_word += [ins(OP_MUL3C, neg0=neg, v0=vs[0], v1=vs[1], v2=vs[2]), mag]
            else:
                raise NotImplementedError("monomial of degree %d" % d)
        for i in range(0, len(lin) - 1, 2):
            single.append(ins(OP_ADD2, neg0=lin[i][0], neg1=lin[i + 1][0], v0=lin[i][1], v1=lin[i + 1][1]))
        if len(lin) % 2:
            single.append(ins(OP_ADD1, neg0=lin[-1][0], v0=lin[-1][1]))
        if not single:                               # zero polynomial, or only two-word terms
            single.append(ins(OP_NOP))
        single[-1] |= 1 << 4
        words.extend(two_word)
        words.extend(single)

    code = prog.code
    for key in order:
        cls, sel = key
        ents = list(groups[key].values())
        for s0 in range(0, len(ents), MAX_GROUP):
            seg = ents[s0:s0 + MAX_GROUP]
            prog.group_pc.append(len(code))
            prog.group_slot.append(len(prog.slot_off) - 1)
            code.append(ins(OP_GROUP, v0=cls, v1=len(sel)))
            for fnode in sel:
                emit_poly(ex.poly(fnode), code)
            for poly, ks in seg:
                emit_poly(poly, code)
                prog.slot_ks.extend(ks)
                prog.slot_off.append(len(prog.slot_ks))
    prog.group_pc.append(len(code))
    prog.group_slot.append(len(prog.slot_off) - 1)
    assert sorted(prog.slot_ks) == list(range(prog.K))
    return prog


MAGIC = b"SBAIRBN1"


def write_airbin(prog, path):
    hdr = struct.pack("<8s12I", MAGIC, prog.n_cols, prog.n_pis, prog.degree, prog.K, len(prog.code), len(prog.consts),
                      len(prog.slot_off) - 1, len(prog.group_pc) - 1, 0, 0, 0, 0)
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(struct.pack("<%dQ" % len(prog.code), *prog.code))
        f.write(struct.pack("<%dQ" % len(prog.consts), *prog.consts))
        f.write(struct.pack("<%dI" % len(prog.slot_off), *prog.slot_off))
        f.write(struct.pack("<%dI" % len(prog.slot_ks), *prog.slot_ks))
        f.write(struct.pack("<%dI" % len(prog.group_pc), *prog.group_pc))
        f.write(struct.pack("<%dI" % len(prog.group_slot), *prog.group_slot))


def read_airbin(path):
    data = open(path, "rb").read()
    magic, n_cols, n_pis, degree, K, n_code, n_consts, n_slots, n_groups, *_ = struct.unpack_from("<8s12I", data, 0)
    assert magic == MAGIC
    off = struct.calcsize("<8s12I")
    prog = Program()
    prog.n_cols, prog.n_pis, prog.degree, prog.K = n_cols, n_pis, degree, K

    def take(fmt, n):
        nonlocal off
        vals = struct.unpack_from("<%d%s" % (n, fmt), data, off)
        off += n * struct.calcsize(fmt)
        return list(vals)

    prog.code = take("Q", n_code)
    prog.consts = take("Q", n_consts)
    prog.slot_off = take("I", n_slots + 1)
    prog.slot_ks = take("I", K)
    prog.group_pc = take("I", n_groups + 1)
    prog.group_slot = take("I", n_groups + 1)
    return prog


def write_flat_air(dag, constraints, n_cols, n_pis, degree, path):
    """SBAIR001: the un-regrouped DAG in emission order, evaluated by the CPU oracle (oracle/air.h)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<8s8I", b"SBAIR001", n_cols, n_pis, degree, len(dag.consts), len(dag.nodes),
                            len(constraints), 0, 0))
        f.write(struct.pack("<%dQ" % len(dag.consts), *dag.consts))
        flat = []
        for op, a, b in dag.nodes:
            flat += (op, a, b)
        f.write(struct.pack("<%dI" % len(flat), *flat))
        flat = []
        for cls, nid in constraints:
            flat += (cls, nid)
        f.write(struct.pack("<%dI" % len(flat), *flat))


def emulate(prog, values, class_factors, weights):
    """Reference interpreter of the bytecode at ONE point (pure Python ints; used by the CPU tests).
    values: list indexed by variable; class_factors: {1:1, 2:z_last, 3:l_first, 4:l_last};
    weights[j][k] = alpha_j^(K-1-k).  Returns [sum_k w_jk f_k c_k for j]."""
    nj = len(weights)
    slot_w = [[sum(w[k] for k in prog.slot_ks[prog.slot_off[s]:prog.slot_off[s + 1]]) % P
               for s in range(len(prog.slot_off) - 1)] for w in weights]
    acc = [0] * nj
    G = [0] * nj
    S, sel_left, have_group, slot, T, cls = 1, 0, False, 0, 0, 1
    code, pc, n = prog.code, 0, len(prog.code)

    def flush():
        for j in range(nj):
            acc[j] = (acc[j] + S * G[j]) % P

    while pc < n:
        w = code[pc]; pc += 1
        op, end, n0, n1 = w & 15, (w >> 4) & 1, (w >> 5) & 1, (w >> 6) & 1
        v0, v1, v2 = (w >> 8) & 0x3FFFF, (w >> 26) & 0x3FFFF, (w >> 44) & 0x3FFFF
        imm = (w >> 26) & 0xFFFFFFFF
        sg = -1 if n0 else 1
        if op == OP_GROUP:
            if have_group: flush()
            have_group, G, cls, sel_left, S, T = True, [0] * nj, v0, v1, 1, 0
            if sel_left == 0: S = class_factors[cls]
            continue
        if op == OP_ADD1: T += sg * values[v0]
        elif op == OP_ADD2: T += sg * values[v0] + (-1 if n1 else 1) * values[v1]
        elif op == OP_SHL1: T += sg * values[v0] * TWO32
        elif op == OP_MULS: T += sg * imm * values[v0]
        elif op == OP_MUL2: T += sg * values[v0] * values[v1]
        elif op == OP_MULC1: T += sg * prog.consts[v2] * values[v0]
        elif op == OP_MULC2: T += sg * prog.consts[v2] * values[v0] * values[v1]
        elif op == OP_MUL3C:
            c = code[pc]; pc += 1
            T += sg * c * values[v0] * values[v1] * values[v2]
        elif op == OP_CONSTI: T += sg * imm
        elif op == OP_CONSTC: T += sg * prog.consts[v2]
        elif op == OP_NOP: pass
        else: raise ValueError("bad opcode %d" % op)
        if end:
            T %= P
            if sel_left:
                S = S * T % P
                sel_left -= 1
                if sel_left == 0: S = S * class_factors[cls] % P
            else:
                for j in range(nj):
                    G[j] = (G[j] + slot_w[j][slot] * T) % P
                slot += 1
            T = 0
    if have_group: flush()
    assert slot == len(prog.slot_off) - 1
    return acc
