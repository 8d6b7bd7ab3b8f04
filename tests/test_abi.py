"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/starky_b200.h declares,
and the pure-host helpers (standard params, proof layout) agree with the oracle's restatement."""
import ctypes as C
import os
import re

import oracle_lib as O
import starky_bls12_381_b200 as sb
from helpers import to_oracle_params

ROOT = O.ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "starky_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = sb.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(L, name), "libstarkyb200.so does not export %s" % name


def test_missing_library_fails_loudly(monkeypatch):
    import starky_bls12_381_b200.binding as B
    monkeypatch.setattr(B, "_LIB", None)
    monkeypatch.setattr(B, "lib_path", lambda: "/nonexistent/libstarkyb200.so")
    try:
        B.lib()
        assert False, "expected ImportError"
    except ImportError as e:
        assert "no CPU fallback" in str(e)


def test_standard_params_match_reference_configs():
    # columns / public inputs / degree / rate bits per stark
    want = {0: (60285, 432, 3, 1), 1: (29376, 4968, 4, 2), 2: (97330, 5064, 3, 1), 3: (73527, 288, 5, 2),
            4: (3339, 12824, 4, 2)}
    for sid, (c, pi, deg, r) in want.items():
        p = sb.standard_params(sid, 10)
        assert (p.n_cols, p.n_public_inputs, p.constraint_degree, p.rate_bits) == (c, pi, deg, r)
        assert (p.cap_height, p.num_challenges, p.pow_bits, p.num_query_rounds) == (4, 2, 16, 84)
    for info in sb.STARKS.values():
        p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1)
        assert (p.n_cols, p.n_public_inputs, p.constraint_degree, p.rate_bits) == (
            info.columns, info.public_inputs, info.constraint_degree, info.rate_bits)


def test_proof_layout_matches_oracle_and_survey():
    # FRI arity bits / final poly length per stark: FP12 [] / 16, PP [4,4] / 4, ML [4] / 64,
    # FE [4,4] / 32, ECC [4,4] / 32
    want = {"fp12_mul": (0, 16), "pairing_precomp": (2, 4), "miller_loop": (1, 64), "final_exp": (2, 32),
            "ecc_agg": (2, 32)}
    for name, info in sb.STARKS.items():
        p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1)
        l = sb.ProofLayout()
        assert sb.lib().sb_proof_layout_for(C.byref(p), C.byref(l)) == 0
        assert (l.n_fri_rounds, l.final_poly_len) == want[name]
        ol = O.layout(to_oracle_params(p))
        assert bytes(l) == bytes(ol)
        for r in range(l.n_fri_rounds):
            assert sb.lib().sb_fri_step_path_len(C.byref(l), r) == l.log_lde - 4 * (r + 1) - 4


def test_host_transcript_permutation_variants_match_oracle():
    """The Fiat-Shamir sponge runs on the host (scalar / AVX2 / AVX-512 paths): each must equal the oracle's permutation."""
    import numpy as np
    L = sb.lib()
    L.sb_host_poseidon_permute_variant.argtypes = [C.c_void_p, C.c_int]
    rng = np.random.default_rng(5)
    states = [np.zeros(12, np.uint64), np.arange(12, dtype=np.uint64), np.full(12, O.P - 1, np.uint64),
              np.full(12, 2 ** 64 - 1, np.uint64)] + [rng.integers(0, 2 ** 63, 12, dtype=np.uint64) * np.uint64(2) for _ in range(50)]
    ran = 0
    for variant in (0, 1, 2, 3, 4, 5):
        for st in states:
            got = st.copy()
            if not L.sb_host_poseidon_permute_variant(got.ctypes.data_as(C.c_void_p), variant):
                break
            want = O.permute(st % np.uint64(O.P))
            assert np.array_equal(got, want), (variant, st)
            ran += 1
    assert ran >= len(states)          # at least the scalar path


def test_bad_fri_parameters_are_rejected_not_looped_on():
    """plonky2 asserts degree_bits >= arity_bits inside ConstantArityBits and would underflow in usize; the C ABI answers
    SB_EINVAL from the one check_params every entry point shares (no 2^30-entry loop, no UB shift)."""
    L = sb.lib()
    l = sb.ProofLayout()
    base = sb.standard_params(sb.StarkId.CUSTOM, 5)
    base.n_cols, base.constraint_degree = 8, 3

    def rc(**kw):
        p = base.copy()
        for k, v in kw.items():
            setattr(p, k, v)
        return L.sb_proof_layout_for(C.byref(p), C.byref(l))
    assert rc() == 0
    assert rc(rate_bits=1, fri_arity_bits=4, fri_final_poly_bits=0, cap_height=2) == -1      # 5 -> 1 -> (1 < 4)
    assert b"arity_bits" in L.sb_last_error(None)
    assert rc(pow_bits=64) == -1
    assert rc(fri_arity_bits=0) == -1
    assert rc(fri_arity_bits=10) == -1
    assert rc(num_query_rounds=0) == -1
    assert rc(num_query_rounds=1 << 20) == -1
    assert rc(fri_final_poly_bits=40) == -1
    assert rc(cap_height=9) == -1
    assert rc(n_cols=0) == -1
    # the oracle applies the same rule instead of wrapping around
    op = to_oracle_params(base)
    op.rate_bits, op.fri_arity_bits, op.fri_final_poly_bits, op.cap_height = 1, 4, 0, 2
    import pytest
    with pytest.raises(RuntimeError):
        O.layout(op)


def test_standard_constraint_programs_are_linked_into_the_library():
    """sb_prove must not depend on files that only the Python harness unpacks: the five programs are .incbin'ed into
    libstarkyb200.so (csrc/air_blobs.S) and equal the committed air/<name>.airbin.xz byte for byte."""
    import lzma
    import os
    from starky_bls12_381_b200 import airfiles
    L = sb.lib()
    for name in airfiles.NAMES.values():
        b = C.c_ubyte.in_dll(L, "sb_airbin_" + name)
        e = C.c_ubyte.in_dll(L, "sb_airbin_%s_end" % name)
        n = C.addressof(e) - C.addressof(b)
        img = C.string_at(C.addressof(b), n)
        want = lzma.decompress(open(os.path.join(airfiles.AIR_DIR, name + ".airbin.xz"), "rb").read())
        assert img == want, name
