"""ctypes binding of libstarkyb200.so -- one Python name per C entry point of include/starky_b200.h."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libstarkyb200 error %d: %s" % (code, msg))
        self.code = code


class StarkId:
    FP12_MUL, PAIRING_PRECOMP, MILLER_LOOP, FINAL_EXP, ECC_AGG, CUSTOM = 0, 1, 2, 3, 4, 100


class TraceLayout:
    COLMAJOR_U64, COLS_U64_PTRS, ROWMAJOR_U64, ROWMAJOR_U32, DEVICE_COLMAJOR_U64 = 0, 1, 2, 3, 4


class Flags:
    ALLOW_INVALID_TRACE, FIXED_POW_WITNESS, OBSERVE_PUBLIC_INPUTS, FRI_MUL_BY_X = 1, 2, 4, 8


class Params(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "stark_id", "log_n", "n_cols", "n_public_inputs", "constraint_degree", "rate_bits", "cap_height",
        "num_challenges", "pow_bits", "num_query_rounds", "fri_arity_bits", "fri_final_poly_bits", "flags",
        "reserved")] + [("fixed_pow_witness", C.c_uint64)]

    def copy(self):
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


class ProofLayout(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "log_n", "log_lde", "n_cols", "n_quotient_polys", "n_public_inputs", "cap_len", "n_fri_rounds",
        "final_poly_len", "n_queries", "arity_bits", "trace_path_len", "reserved")] + [(n, C.c_uint64) for n in (
        "off_trace_cap", "off_quotient_cap", "off_local_values", "off_next_values", "off_quotient_polys",
        "off_fri_caps", "off_final_poly", "off_pow_witness", "off_queries", "query_stride", "q_off_trace_leaf",
        "q_off_trace_path", "q_off_quot_leaf", "q_off_quot_path", "q_off_steps", "off_public_inputs",
        "total_words")]


class _CProof(C.Structure):
    _fields_ = [("layout", ProofLayout), ("words", C.POINTER(C.c_uint64))] + [(n, C.c_float) for n in (
        "ms_h2d", "ms_trace_commit", "ms_quotient", "ms_quotient_commit", "ms_openings", "ms_fri", "ms_d2h",
        "ms_total")]


HOOK_COMMIT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64))
HOOK_QUOTIENT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
HOOK_OPENINGS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                            C.POINTER(C.c_uint64))
HOOK_COMBINE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
HOOK_QUERY_ROWS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p)


class Job(C.Structure):
    """Binary-identical to sb_job (include/starky_b200.h)."""
    _fields_ = [("params", Params), ("trace", C.c_void_p), ("layout", C.c_int), ("public_inputs", C.c_void_p),
                ("proof", C.POINTER(_CProof)), ("rc", C.c_int), ("ms", C.c_float)]


class ShardHooks(C.Structure):
    """Binary-identical to sb_shard_hooks (include/starky_b200.h)."""
    _fields_ = [("user", C.c_void_p), ("commit", HOOK_COMMIT), ("quotient", HOOK_QUOTIENT), ("openings", HOOK_OPENINGS),
                ("combine", HOOK_COMBINE), ("query_rows", HOOK_QUERY_ROWS)]


def lib_path():
    # SB_LIBRARY: another build of the same library (A/B runs of compile-time kernel variants); still no CPU fallback
    return os.environ.get("SB_LIBRARY") or os.path.join(_HERE, "libstarkyb200.so")


def lib():
    """Load the CUDA library.  There is no CPU fallback: a missing library is an error."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C starky_bls12_381_b200/csrc); there is no CPU fallback" % path)
        L = C.CDLL(path)
        L.sb_last_error.restype = C.c_char_p
        L.sb_last_error.argtypes = [C.c_void_p]
        L.sb_init.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.sb_destroy.argtypes = [C.c_void_p]
        L.sb_params_standard.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(Params)]
        L.sb_proof_layout_for.argtypes = [C.POINTER(Params), C.POINTER(ProofLayout)]
        L.sb_fri_step_path_len.argtypes = [C.POINTER(ProofLayout), C.c_uint32]
        L.sb_fri_step_path_len.restype = C.c_uint32
        L.sb_fri_step_offset.argtypes = [C.POINTER(ProofLayout), C.c_uint32]
        L.sb_fri_step_offset.restype = C.c_uint64
        L.sb_air_load.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p]
        L.sb_prove.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p,
                               C.POINTER(C.POINTER(_CProof))]
        L.sb_proof_free.argtypes = [C.POINTER(_CProof)]
        L.sb_lde_commit.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
        L.sb_coeffs_download.argtypes = [C.c_void_p, C.c_void_p]
        L.sb_quotient_values.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p]
        L.sb_ntt_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.sb_poseidon_permute_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.sb_hash_leaves.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.sb_trace_upload.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int]
        L.sb_kernel_launches.argtypes = [C.c_void_p]
        L.sb_kernel_launches.restype = C.c_uint64
        L.sb_stage_ms.argtypes = [C.c_void_p, C.c_char_p]
        L.sb_stage_ms.restype = C.c_float
        L.sb_measure_imad_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.sb_synchronize.argtypes = [C.c_void_p]
        L.sb_lde_cols_device.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.sb_hash_rows_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.sb_lde_cols_peer_device.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_void_p, C.c_void_p]
        L.sb_merkle_from_position_digests.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p]
        L.sb_quotient_rows_device.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        L.sb_transcript_alphas.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.sb_prove_sharded.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(ShardHooks), C.c_void_p,
                                       C.POINTER(C.POINTER(_CProof))]
        L.sb_openings_cols_device.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        L.sb_combine_cols_device.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                             C.c_void_p]
        L.sb_memcpy_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.sb_proof_serialize.argtypes = [C.POINTER(_CProof), C.POINTER(Params), C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.sb_proof_deserialize.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Params), C.POINTER(Params),
                                           C.POINTER(C.POINTER(_CProof))]
        L.sb_proof_from_words.argtypes = [C.POINTER(Params), C.c_void_p, C.c_size_t, C.POINTER(C.POINTER(_CProof))]
        L.sb_witness_fp12_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.sb_prove_fp12_mul.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(_CProof))]
        L.sb_witness_last_error.restype = C.c_char_p
        L.sb_openings.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p]
        L.sb_fri_commit.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sb_prove_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Job), C.c_int]
        L.sb_group_unique_id.argtypes = [C.c_void_p]
        L.sb_group_init_rank.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.sb_group_init_local.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
        L.sb_group_destroy.argtypes = [C.c_void_p]
        L.sb_group_rank.argtypes = [C.c_void_p]
        L.sb_group_size.argtypes = [C.c_void_p]
        L.sb_shard_columns.argtypes = [C.POINTER(Params), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_uint32)]
        L.sb_group_column_slice.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.sb_group_prove.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p, C.c_uint32,
                                     C.POINTER(C.POINTER(_CProof))]
        L.sb_group_phase_ms.argtypes = [C.c_void_p, C.c_char_p]
        L.sb_group_phase_ms.restype = C.c_float
        _LIB = L
    return _LIB


def standard_params(stark_id, log_n, flags=0):
    p = Params()
    rc = lib().sb_params_standard(stark_id, log_n, C.byref(p))
    if rc:
        raise SbError(rc, "sb_params_standard")
    p.flags = flags
    return p


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class WireFormat:
    POD, PLONKY2_BUFFER, SERDE_JSON = 0, 1, 2


def serialize_words(p, words, fmt):
    """sb_proof_serialize of the proof whose POD words are `words` (layout of `p`) -> bytes.  Host only."""
    w = np.ascontiguousarray(words, dtype=np.uint64)
    cp = C.POINTER(_CProof)()
    rc = lib().sb_proof_from_words(C.byref(p), _ptr(w), w.size, C.byref(cp))
    if rc:
        raise SbError(rc, lib().sb_last_error(None).decode())
    try:
        n = C.c_size_t()
        rc = lib().sb_proof_serialize(cp, C.byref(p), fmt, None, 0, C.byref(n))
        if rc:
            raise SbError(rc, lib().sb_last_error(None).decode())
        buf = (C.c_ubyte * n.value)()
        rc = lib().sb_proof_serialize(cp, C.byref(p), fmt, buf, n.value, C.byref(n))
        if rc:
            raise SbError(rc, lib().sb_last_error(None).decode())
        return bytes(buf)
    finally:
        lib().sb_proof_free(cp)


def deserialize_words(data, fmt, p=None):
    """sb_proof_deserialize -> (Params of the image, POD words)."""
    cp = C.POINTER(_CProof)()
    q = Params()
    rc = lib().sb_proof_deserialize(data, len(data), fmt, C.byref(p) if p is not None else None, C.byref(q), C.byref(cp))
    if rc:
        raise SbError(rc, lib().sb_last_error(None).decode())
    return q, Proof(cp).words


class _ProofMemory:
    """Keeps one sb_proof alive for as long as a numpy view of its words exists (the array's base is this object)."""

    def __init__(self, cproof, n_words):
        self._cp = cproof
        self.__array_interface__ = {"shape": (int(n_words),), "typestr": "<u8", "version": 3,
                                    "data": (C.cast(cproof.contents.words, C.c_void_p).value, False)}

    def __del__(self):
        cp, self._cp = getattr(self, "_cp", None), None
        if cp:
            try:
                lib().sb_proof_free(cp)
            except Exception:
                pass


class Proof:
    """Owns one sb_proof; `words` is a zero-copy numpy view of the flat POD in the library's pinned buffer.  The buffer goes
    back to the library's pool when the last view of it is dropped (no 50 MB copy per FinalExp proof, as in the Rust shim)."""

    def __init__(self, cproof):
        c = cproof.contents
        self.layout = ProofLayout()
        C.memmove(C.byref(self.layout), C.byref(c.layout), C.sizeof(ProofLayout))
        self.timings = {k: getattr(c, k) for k in ("ms_h2d", "ms_trace_commit", "ms_quotient", "ms_quotient_commit",
                                                   "ms_openings", "ms_fri", "ms_d2h", "ms_total")}
        self.words = np.asarray(_ProofMemory(cproof, self.layout.total_words))

    def field(self, name, count):
        off = getattr(self.layout, name)
        return self.words[off:off + count]


def fp12_limbs(x):
    """Fp12 as twelve Python ints (the reference's [Fp; 12]) -> uint32 [144] little-endian limbs."""
    return np.array([(int(v) >> (32 * i)) & 0xFFFFFFFF for v in x for i in range(12)], dtype=np.uint32)


def witness_fp12_mul(x, y, num_rows=16):
    """sb_witness_fp12_mul: (trace uint32 [num_rows][60285] row-major, public inputs uint64 [432])."""
    xl, yl = fp12_limbs(x), fp12_limbs(y)
    trace = np.empty((num_rows, 60285), np.uint32)
    pis = np.empty(432, np.uint64)
    rc = lib().sb_witness_fp12_mul(_ptr(xl), _ptr(yl), num_rows, _ptr(trace), _ptr(pis))
    if rc:
        raise SbError(rc, lib().sb_witness_last_error().decode())
    return trace, pis


def g1_limbs(points):
    """512 affine G1 points [(x, y)] as the [512][24] uint32 limb array of sb_witness_ecc_agg."""
    out = np.zeros((len(points), 24), np.uint32)
    for i, (x, y) in enumerate(points):
        for k in range(12):
            out[i, k] = (int(x) >> (32 * k)) & 0xFFFFFFFF
            out[i, 12 + k] = (int(y) >> (32 * k)) & 0xFFFFFFFF
    return out


def witness_ecc_agg(points, bits, num_rows=8192):
    """sb_witness_ecc_agg: (trace uint32 [num_rows][3339] row-major, public inputs uint64 [12824], aggregate point (x, y))."""
    pl = g1_limbs(points)
    bl = np.ascontiguousarray(np.array([1 if b else 0 for b in bits], np.uint8))
    trace = np.empty((num_rows, 3339), np.uint32)
    pis = np.empty(12824, np.uint64)
    res = np.zeros(24, np.uint32)
    L = lib()
    L.sb_witness_ecc_agg.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.sb_witness_ecc_agg(_ptr(pl), _ptr(bl), num_rows, _ptr(trace), _ptr(pis), _ptr(res))
    if rc:
        raise SbError(rc, L.sb_witness_last_error().decode())
    val = lambda l: sum(int(v) << (32 * k) for k, v in enumerate(l))
    return trace, pis, (val(res[:12]), val(res[12:]))


def fp2_limbs(vals):
    """Fp2 values [(c0, c1), ...] -> uint32 [len][24] little-endian limbs."""
    return np.array([[(int(c) >> (32 * i)) & 0xFFFFFFFF for c in v for i in range(12)] for v in vals], dtype=np.uint32)


def _witness_call(fn_name, args, num_rows, n_cols, n_pis):
    trace = np.empty((num_rows, n_cols), np.uint32)
    pis = np.empty(n_pis, np.uint64)
    L = lib()
    fn = getattr(L, fn_name)
    fn.argtypes = [C.c_void_p] * len(args) + [C.c_uint32, C.c_void_p, C.c_void_p]
    rc = fn(*[_ptr(a) for a in args], num_rows, _ptr(trace), _ptr(pis))
    if rc:
        raise SbError(rc, L.sb_witness_last_error().decode())
    return trace, pis


def witness_pairing_precomp(x, y, z, num_rows=1024):
    """sb_witness_pairing_precomp: (trace uint32 [num_rows][29376] row-major, public inputs uint64 [4968])."""
    return _witness_call("sb_witness_pairing_precomp", [fp2_limbs([x, y, z])], num_rows, 29376, 4968)


def witness_miller_loop(x, y, q, num_rows=1024):
    """sb_witness_miller_loop: (trace uint32 [num_rows][97330] row-major, public inputs uint64 [5064])."""
    g1 = np.array([(int(v) >> (32 * i)) & 0xFFFFFFFF for v in (x, y) for i in range(12)], dtype=np.uint32)
    return _witness_call("sb_witness_miller_loop", [g1, fp2_limbs(list(q))], num_rows, 97330, 5064)


def witness_final_exp(x, num_rows=8192):
    """sb_witness_final_exp: (trace uint32 [8192][73527] row-major, public inputs uint64 [288])."""
    return _witness_call("sb_witness_final_exp", [fp12_limbs(x)], num_rows, 73527, 288)


def prove_batch(contexts, jobs):
    """sb_prove_batch: jobs = [(Params, trace pointer or array, layout, public inputs)], returns [(Proof or SbError, ms)] in
    job order.  The proofs run on `contexts` with the library's scheduler (internal host threads)."""
    n = len(jobs)
    arr = (Job * n)()
    keep = []
    for i, (p, trace, layout, pis) in enumerate(jobs):
        pis = _u64(pis)
        keep.append((pis, trace))
        arr[i].params = p
        arr[i].trace = _ptr(trace)
        arr[i].layout = layout
        arr[i].public_inputs = _ptr(pis) if pis.size else None
    hs = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    lib().sb_prove_batch(hs, len(contexts), arr, n)
    out = []
    for i in range(n):
        if arr[i].rc:
            out.append((SbError(arr[i].rc, "job %d failed" % i), arr[i].ms))
        else:
            out.append((Proof(arr[i].proof), arr[i].ms))
    return out


class Context:
    """One sb_ctx: one GPU, one stream, resident device buffers reused across proofs.  Not re-entrant."""

    def __init__(self, device=None):
        """device: None (current device), an index, or a list of indices (one ctx over several GPUs: prove() shards)."""
        self._h = C.c_void_p()
        if device is None:
            rc = lib().sb_init(None, 0, C.byref(self._h))
        else:
            devs = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
            arr = (C.c_int * len(devs))(*devs)
            rc = lib().sb_init(arr, len(devs), C.byref(self._h))
        if rc:
            raise SbError(rc, lib().sb_last_error(None).decode())

    def close(self):
        if self._h:
            lib().sb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise SbError(rc, lib().sb_last_error(self._h).decode())

    def air_load(self, stark_id, path):
        self._check(lib().sb_air_load(self._h, stark_id, os.fsencode(path)))

    def trace_upload(self, p, trace, layout=TraceLayout.COLMAJOR_U64):
        self._check(lib().sb_trace_upload(self._h, C.byref(p), _ptr(trace), layout))

    def lde_commit(self, p, trace, layout=TraceLayout.COLMAJOR_U64, want_lde=True, want_digests=True):
        N = 1 << (p.log_n + p.rate_bits)
        lde = np.empty((p.n_cols, N), np.uint64) if want_lde else None
        dig = np.empty((N, 4), np.uint64) if want_digests else None
        cap = np.empty((1 << p.cap_height, 4), np.uint64)
        self._check(lib().sb_lde_commit(self._h, C.byref(p), _ptr(trace), layout, _ptr(lde), _ptr(dig), _ptr(cap)))
        return dict(lde=lde, digests=dig, cap=cap)

    def coeffs(self, p):
        out = np.empty((p.n_cols, 1 << p.log_n), np.uint64)
        self._check(lib().sb_coeffs_download(self._h, _ptr(out)))
        return out

    def quotient_values(self, p, public_inputs, alphas):
        N = 1 << (p.log_n + p.rate_bits)
        out = np.empty((p.num_challenges, N), np.uint64)
        pis, al = _u64(public_inputs), _u64(alphas)
        self._check(lib().sb_quotient_values(self._h, C.byref(p), _ptr(pis), _ptr(al), _ptr(out)))
        return out

    def openings(self, p, zeta):
        """sb_openings: (local_values, next_values) of the committed trace at zeta / g zeta, uint64 [n_cols][2] each."""
        z = _u64(zeta)
        loc, nxt = np.empty((p.n_cols, 2), np.uint64), np.empty((p.n_cols, 2), np.uint64)
        self._check(lib().sb_openings(self._h, C.byref(p), _ptr(z), _ptr(loc), _ptr(nxt)))
        return loc, nxt

    def fri_commit(self, p, coeffs, betas):
        """sb_fri_commit: (caps [rounds][2^cap_height][4], final_poly [len][2]) for injected folding challenges."""
        l = ProofLayout()
        self._check(lib().sb_proof_layout_for(C.byref(p), C.byref(l)))
        c, b = _u64(coeffs), _u64(betas)
        caps = np.zeros((l.n_fri_rounds, l.cap_len, 4), np.uint64)
        fin = np.zeros((l.final_poly_len, 2), np.uint64)
        self._check(lib().sb_fri_commit(self._h, C.byref(p), _ptr(c), _ptr(b) if b.size else None, _ptr(caps), _ptr(fin)))
        return caps, fin

    def ntt_batch(self, data, inverse=False):
        d = _u64(data).copy()
        count, n = d.shape
        self._check(lib().sb_ntt_batch(self._h, _ptr(d), int(n).bit_length() - 1, count, int(inverse)))
        return d

    def poseidon_permute_batch(self, states):
        s = _u64(states).copy()
        self._check(lib().sb_poseidon_permute_batch(self._h, _ptr(s), s.shape[0]))
        return s

    def hash_leaves(self, cols):
        c = _u64(cols)
        out = np.empty((c.shape[1], 4), np.uint64)
        self._check(lib().sb_hash_leaves(self._h, _ptr(c), c.shape[0], c.shape[1], _ptr(out)))
        return out

    def prove(self, p, trace, public_inputs, layout=TraceLayout.COLMAJOR_U64):
        out = C.POINTER(_CProof)()
        pis = _u64(public_inputs)
        self._check(lib().sb_prove(self._h, C.byref(p), _ptr(trace), layout, _ptr(pis), C.byref(out)))
        return Proof(out)

    def prove_fp12_mul(self, p, x, y):
        """sb_prove_fp12_mul: FP12MulStark proof from the two Fp12 operands (witness generated in C++ on the host)."""
        out = C.POINTER(_CProof)()
        xl, yl = fp12_limbs(x), fp12_limbs(y)
        self._check(lib().sb_prove_fp12_mul(self._h, C.byref(p), _ptr(xl), _ptr(yl), C.byref(out)))
        return Proof(out)

    def prove_pairing_precomp(self, p, x, y, z):
        """sb_prove_pairing_precomp: proof from the projective G2 point (witness generated in C++ on the host)."""
        out = C.POINTER(_CProof)()
        L = lib()
        L.sb_prove_pairing_precomp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        q = fp2_limbs([x, y, z])
        self._check(L.sb_prove_pairing_precomp(self._h, C.byref(p), _ptr(q), C.byref(out)))
        return Proof(out)

    def prove_miller_loop(self, p, x, y, q):
        """sb_prove_miller_loop: proof from the G1 point (x, y) and the projective G2 point q."""
        out = C.POINTER(_CProof)()
        L = lib()
        L.sb_prove_miller_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        g1 = np.array([(int(v) >> (32 * i)) & 0xFFFFFFFF for v in (x, y) for i in range(12)], dtype=np.uint32)
        ql = fp2_limbs(list(q))
        self._check(L.sb_prove_miller_loop(self._h, C.byref(p), _ptr(g1), _ptr(ql), C.byref(out)))
        return Proof(out)

    def prove_final_exp(self, p, x):
        """sb_prove_final_exp: proof from the Fp12 input."""
        out = C.POINTER(_CProof)()
        L = lib()
        L.sb_prove_final_exp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        xl = fp12_limbs(x)
        self._check(L.sb_prove_final_exp(self._h, C.byref(p), _ptr(xl), C.byref(out)))
        return Proof(out)

    def synchronize(self):
        self._check(lib().sb_synchronize(self._h))

    def kernel_launches(self):
        return int(lib().sb_kernel_launches(self._h))

    def stage_ms(self, name):
        return float(lib().sb_stage_ms(self._h, name.encode()))

    def measure_imad_peak(self):
        out = (C.c_double * 2)()
        self._check(lib().sb_measure_imad_peak(self._h, out))
        return {"mad_lo_u32_gops": out[0], "mad_wide_u32_gops": out[1]}
