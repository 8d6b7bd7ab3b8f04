"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

P = 0xFFFFFFFF00000001


class Params(C.Structure):
    """Binary-identical to sb_params (include/starky_b200.h)."""
    _fields_ = [(n, C.c_uint32) for n in (
        "stark_id", "log_n", "n_cols", "n_public_inputs", "constraint_degree", "rate_bits", "cap_height",
        "num_challenges", "pow_bits", "num_query_rounds", "fri_arity_bits", "fri_final_poly_bits", "flags",
        "reserved")] + [("fixed_pow_witness", C.c_uint64)]


class Layout(C.Structure):
    """Binary-identical to sb_proof_layout."""
    _fields_ = [(n, C.c_uint32) for n in (
        "log_n", "log_lde", "n_cols", "nq", "n_pis", "cap_len", "n_fri_rounds", "final_poly_len", "n_queries",
        "arity_bits", "trace_path_len", "reserved")] + [(n, C.c_uint64) for n in (
        "off_trace_cap", "off_quotient_cap", "off_local", "off_next", "off_quot_open", "off_fri_caps",
        "off_final_poly", "off_pow", "off_queries", "query_stride", "q_trace_leaf", "q_trace_path", "q_quot_leaf",
        "q_quot_path", "q_steps", "off_pis", "total_words")]


def make_params(stark_id=100, log_n=4, n_cols=8, n_pis=0, degree=3, rate_bits=1, cap_height=4, pow_bits=16,
                num_queries=84, flags=0, fixed_pow_witness=0):
    return Params(stark_id, log_n, n_cols, n_pis, degree, rate_bits, cap_height, 2, pow_bits, num_queries, 4, 5,
                  flags, 0, fixed_pow_witness)


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        L.orc_gl_mul.restype = C.c_uint64
        L.orc_gl_mul.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_gl_mul_slow.restype = C.c_uint64
        L.orc_gl_mul_slow.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_gl_root.restype = C.c_uint64
        L.orc_hash_no_pad.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_hash_or_noop.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_challenger_run.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_merkle.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_void_p, C.c_void_p]
        L.orc_prove.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_verify.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _LIB = L
    return _LIB


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def err():
    return lib().orc_last_error().decode()


def permute(state):
    s = u64(state).copy()
    lib().orc_poseidon_permute(ptr(s))
    return s


def hash_no_pad(x):
    x = u64(x); out = np.zeros(4, np.uint64)
    lib().orc_hash_no_pad(ptr(x), x.size, ptr(out))
    return out


def hash_or_noop(x):
    x = u64(x); out = np.zeros(4, np.uint64)
    lib().orc_hash_or_noop(ptr(x), x.size, ptr(out))
    return out


def two_to_one(l, r):
    l, r = u64(l), u64(r); out = np.zeros(4, np.uint64)
    lib().orc_two_to_one(ptr(l), ptr(r), ptr(out))
    return out


def ntt_batch(data, inverse=False):
    d = u64(data).copy()
    count, n = d.shape
    lib().orc_ntt_batch(ptr(d), int(n).bit_length() - 1, count, int(inverse))
    return d


def challenger_run(obs, n_out):
    obs = u64(obs); out = np.zeros(n_out, np.uint64)
    lib().orc_challenger_run(ptr(obs), obs.size, ptr(out), n_out)
    return out


def layout(p):
    l = Layout()
    if lib().orc_layout(C.byref(p), C.byref(l)):
        raise RuntimeError("oracle: " + err())
    return l


def lde_commit(p, trace_colmajor, want_leaves=True, want_coeffs=False):
    """-> dict(leaves [N][C], digests [N][4], cap [2^h][4], coeffs [C][n])"""
    t = u64(trace_colmajor)
    N = 1 << (p.log_n + p.rate_bits)
    leaves = np.zeros((N, p.n_cols), np.uint64) if want_leaves else None
    dig = np.zeros((N, 4), np.uint64)
    cap = np.zeros((1 << p.cap_height, 4), np.uint64)
    coeffs = np.zeros((p.n_cols, 1 << p.log_n), np.uint64) if want_coeffs else None
    rc = lib().orc_lde_commit(C.byref(p), ptr(t), ptr(leaves), ptr(dig), ptr(cap), ptr(coeffs))
    assert rc == 0, err()
    return dict(leaves=leaves, digests=dig, cap=cap, coeffs=coeffs)


def merkle(leaves, cap_height):
    lv = u64(leaves)
    dig = np.zeros((lv.shape[0], 4), np.uint64)
    cap = np.zeros((1 << cap_height, 4), np.uint64)
    rc = lib().orc_merkle(ptr(lv), lv.shape[0], lv.shape[1], cap_height, ptr(dig), ptr(cap))
    assert rc == 0
    return dig, cap


def quotient_values(air_path, p, trace_colmajor, pis, alphas):
    t, pi, al = u64(trace_colmajor), u64(pis), u64(alphas)
    qdf = max(1, p.constraint_degree - 1)
    qbits = (qdf - 1).bit_length()
    out = np.zeros((p.num_challenges, (1 << p.log_n) << qbits), np.uint64)
    rc = lib().orc_quotient_values(air_path.encode(), C.byref(p), ptr(t), ptr(pi), ptr(al), ptr(out))
    assert rc == 0, err()
    return out


def eval_constraints_row(air_path, local, nxt, pis):
    info = air_info(air_path)
    out = np.zeros(info["n_constraints"], np.uint64)
    rc = lib().orc_eval_constraints_row(air_path.encode(), ptr(u64(local)), ptr(u64(nxt)), ptr(u64(pis)), ptr(out))
    assert rc == 0, err()
    return out


def air_info(air_path):
    o = np.zeros(6, np.uint32)
    rc = lib().orc_air_info(air_path.encode(), ptr(o))
    assert rc == 0, err()
    return dict(zip(("n_cols", "n_pis", "degree", "n_consts", "n_nodes", "n_constraints"), map(int, o)))


def prove(air_path, p, trace_colmajor, pis):
    """-> (rc, words or None)"""
    l = layout(p)
    words = np.zeros(l.total_words, np.uint64)
    rc = lib().orc_prove(air_path.encode(), C.byref(p), ptr(u64(trace_colmajor)), ptr(u64(pis)), ptr(words),
                         words.size)
    return rc, (words if rc == 0 else None)


def verify(air_path, p, words):
    w = u64(words)
    return lib().orc_verify(air_path.encode(), C.byref(p), ptr(w), w.size)


def fri_commit(p, coeffs, betas):
    """Commit phase of FRI with injected folding challenges -> (caps [rounds][cap_len][4], final_poly [len][2])."""
    l = layout(p)
    c, b = u64(coeffs), u64(betas)
    caps = np.zeros((l.n_fri_rounds, l.cap_len, 4), np.uint64)
    fin = np.zeros((l.final_poly_len, 2), np.uint64)
    n = lib().orc_fri_commit(C.byref(p), ptr(c), ptr(b), ptr(caps), ptr(fin))
    assert n == l.final_poly_len, err()
    return caps, fin
