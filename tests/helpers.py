"""Shared helpers for the parity tests."""
import numpy as np

import oracle_lib as O

P = O.P


def to_oracle_params(p):
    return O.Params.from_buffer_copy(bytes(p))


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def bitrev_perm(bits):
    return np.array([bitrev(i, bits) for i in range(1 << bits)], dtype=np.int64)


def pos_to_leaf(log_n, rate_bits):
    """Device LDE position J*n + k  ->  plonky2 leaf index J*n + bitrev_n(k) (ntt.cu layout)."""
    n = 1 << log_n
    rev = bitrev_perm(log_n)
    return np.concatenate([J * n + rev for J in range(1 << rate_bits)])


def pos_to_natural(log_n, rate_bits):
    """Device LDE position J*n + k -> natural LDE index bitrev_r(J) + 2^r k."""
    n = 1 << log_n
    k = np.arange(n, dtype=np.int64)
    return np.concatenate([bitrev(J, rate_bits) + (k << rate_bits) for J in range(1 << rate_bits)])


def random_trace(rng, n_cols, log_n, full_width=False):
    if full_width:
        return (rng.integers(0, 1 << 63, (n_cols, 1 << log_n), dtype=np.uint64) * np.uint64(2)
                + rng.integers(0, 2, (n_cols, 1 << log_n), dtype=np.uint64)) % np.uint64(P)
    return rng.integers(0, 1 << 32, (n_cols, 1 << log_n), dtype=np.uint64)
