"""Small hand-written AIRs for the tests: built as expression DAGs with the same builder tools/airgen uses, written in
both formats (flat .air for the oracle, .airbin for the GPU), with numpy witness generators for VALID traces."""
import os
import sys

import numpy as np

import oracle_lib as O

sys.path.insert(0, O.ROOT)
from tools.airgen.compile import compile_program, write_airbin, write_flat_air  # noqa: E402
from tools.airgen.rustsym import ColVec, Dag, LOCAL, NEXT, PI, Sym  # noqa: E402

P = O.P


class Builder:
    def __init__(self, n_cols, n_pis, degree):
        self.dag = Dag()
        self.n_cols, self.n_pis, self.degree = n_cols, n_pis, degree
        self.lv, self.nv, self.pi = (ColVec(self.dag, k, n) for k, n in ((LOCAL, n_cols), (NEXT, n_cols), (PI, max(n_pis, 1))))
        self.cons = []

    def c(self, v):
        return Sym(self.dag, self.dag.const(v))

    def constraint(self, e): self.cons.append((1, e.id))
    def transition(self, e): self.cons.append((2, e.id))
    def first_row(self, e): self.cons.append((3, e.id))
    def last_row(self, e): self.cons.append((4, e.id))

    def write(self, directory, name):
        flat = os.path.join(directory, name + ".air")
        binp = os.path.join(directory, name + ".airbin")
        write_flat_air(self.dag, self.cons, self.n_cols, self.n_pis, self.degree, flat)
        write_airbin(compile_program(self.dag, self.cons, self.n_cols, self.n_pis, self.degree), binp)
        return flat, binp


def fibonacci(directory):
    """2 columns; degree-3 declared (quotient factor 2, rate_bits 1).  PIs: a0, b0, b_last."""
    b = Builder(2, 3, 3)
    L, N, PI_ = b.lv, b.nv, b.pi
    b.first_row(L[0] - PI_[0])
    b.first_row(L[1] - PI_[1])
    b.transition(N[0] - L[1])
    b.transition(N[1] - L[0] - L[1])
    b.last_row(L[1] - PI_[2])
    flat, binp = b.write(directory, "fib")

    def witness(log_n, a0=1, b0=2):
        n = 1 << log_n
        t = np.zeros((2, n), np.uint64)
        x, y = a0, b0
        for i in range(n):
            t[0, i], t[1, i] = x, y
            x, y = y, (x + y) % P
        return t, np.array([a0, b0, int(t[1, n - 1])], np.uint64)
    return dict(flat=flat, airbin=binp, n_cols=2, n_pis=3, degree=3, rate_bits=1, witness=witness)


def limbs(directory, degree=4):
    """A 16-column gadget in the style of the reference's Fp limb arithmetic, touching every opcode of the bytecode:
      col 0 sel (boolean), 1 x, 2 y, 3 lo, 4 carry, 5 acc, 6..9 bits b0..b3, 10 cube, 11 mux, 12 bigmul, 13 z, 14 w, 15 cnt
      sel*(x*y - lo - carry*2^32)                       (MUL2, SHL1, ADD1/ADD2; selector group)
      sel*(1-sel)                                        (two selector-ish factors)
      (b0 + 2 b1 + 4 b2 + 8 b3) - z   twice              (MULS; duplicate body -> shared weight slot)
      b_i*(b_i - 1)                                      (MUL2)
      cube - 3*x*y*z                                     (MUL3C, degree 3)
      mux - ((1-b0)*C0 + b0*C1)                          (CONSTC / MULC1 with 64-bit constants)
      bigmul - BIG*x*y                                   (MULC2)
      w - 5                                              (CONSTI)
      transition: next.acc - acc - lo ; next.cnt - cnt - 1 ; sel*(next.x - x) wraps-free
      first row: acc - pi0, cnt ; last row: acc - pi1
    degree = 4 -> quotient factor 3, rate_bits 2."""
    C0, C1, BIG = 0xFEDCBA9876543210 % P, 0x123456789ABCDEF1 % P, 0xABCDEF0123456789 % P
    b = Builder(16, 2, degree)
    L, N, PI_ = b.lv, b.nv, b.pi
    one = b.c(1)
    sel = L[0]
    b.transition(sel * (L[1] * L[2] - L[3] - L[4] * b.c(1 << 32)))
    b.constraint(sel * (one - sel))
    rec = L[6] + L[7] * b.c(2) + L[8] * b.c(4) + L[9] * b.c(8)
    b.constraint(sel * (rec - L[13]))
    b.constraint(sel * (rec - L[13]))
    for i in range(6, 10):
        b.constraint(L[i] * (L[i] - one))
    b.constraint(L[10] - L[1] * L[2] * L[13] * b.c(3))
    b.constraint(L[11] - ((one - L[6]) * b.c(C0) + L[6] * b.c(C1)))
    b.constraint(L[12] - L[1] * L[2] * b.c(BIG))
    b.constraint(L[14] - b.c(5))
    b.transition(N[5] - L[5] - L[3])
    b.transition(N[15] - L[15] - one)
    b.transition(sel * (one - N[0]) * (N[1] - L[1]) * b.c(0))      # a vanishing constraint with three factors
    b.first_row(L[5] - PI_[0])
    b.first_row(L[15])
    b.last_row(L[5] - PI_[1])
    flat, binp = b.write(directory, "limbs%d" % degree)

    def witness(log_n, seed=1):
        n = 1 << log_n
        rng = np.random.default_rng(seed)
        t = np.zeros((16, n), dtype=object)
        acc = int(rng.integers(0, 1 << 32))
        pi0 = acc
        for i in range(n):
            s = int(rng.integers(0, 2))
            x, y = int(rng.integers(0, 1 << 32)), int(rng.integers(0, 1 << 32))
            prod = x * y
            lo, carry = prod & 0xFFFFFFFF, prod >> 32
            z = int(rng.integers(0, 16))
            bits = [(z >> k) & 1 for k in range(4)]
            t[0, i], t[1, i], t[2, i], t[3, i], t[4, i], t[5, i] = s, x, y, lo, carry, acc
            for k in range(4): t[6 + k, i] = bits[k]
            t[10, i] = 3 * x * y * z % P
            t[11, i] = C1 if bits[0] else C0
            t[12, i] = BIG * x * y % P
            t[13, i], t[14, i], t[15, i] = z, 5, i
            acc = (acc + lo) % P
        pi1 = int(t[5, n - 1])
        return t.astype(np.uint64), np.array([pi0, pi1], np.uint64)
    rate_bits = {3: 1, 4: 2, 5: 2}[degree]
    return dict(flat=flat, airbin=binp, n_cols=16, n_pis=2, degree=degree, rate_bits=rate_bits, witness=witness)
