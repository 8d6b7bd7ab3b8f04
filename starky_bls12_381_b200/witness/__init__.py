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
