"""Host-side mirror of the reference's interface for the prove() path (Python here only because the image has no
Rust toolchain; the real shim is rust/starky_gpu).  Names follow starky: StarkConfig::standard_fast_config(),
the five Stark structs with `new(num_rows)`, and prove(stark, config, trace_poly_values, public_inputs).

Reference: /root/reference/src/aggregate_proof.rs:32-34,57-59 (PairingPrecomp), :76-77,104-105 (MillerLoop),
:122-123,137-138 (FP12Mul), :155-157,168-169 (FinalExp), :186-188,211-212 (ECCAggregate).
"""
import dataclasses

from . import binding as B


@dataclasses.dataclass
class FriConfig:
    rate_bits: int = 1
    cap_height: int = 4
    proof_of_work_bits: int = 16
    reduction_arity_bits: int = 4     # FriReductionStrategy::ConstantArityBits(4, 5)
    final_poly_bits: int = 5
    num_query_rounds: int = 84


@dataclasses.dataclass
class StarkConfig:
    security_bits: int = 100
    num_challenges: int = 2
    fri_config: FriConfig = dataclasses.field(default_factory=FriConfig)

    @staticmethod
    def standard_fast_config():
        return StarkConfig()


@dataclasses.dataclass(frozen=True)
class StarkInfo:
    """One of the reference's five `impl Stark` structs."""
    name: str
    stark_id: int
    columns: int
    public_inputs: int
    constraint_degree: int
    rate_bits: int          # the override the reference applies in aggregate_proof.rs
    num_rows: int           # the row count the reference instantiates it with


STARKS = {
    "fp12_mul": StarkInfo("FP12MulStark", B.StarkId.FP12_MUL, 60285, 432, 3, 1, 16),
    "pairing_precomp": StarkInfo("PairingPrecompStark", B.StarkId.PAIRING_PRECOMP, 29376, 4968, 4, 2, 1024),
    "miller_loop": StarkInfo("MillerLoopStark", B.StarkId.MILLER_LOOP, 97330, 5064, 3, 1, 1024),
    "final_exp": StarkInfo("FinalExponentiateStark", B.StarkId.FINAL_EXP, 73527, 288, 5, 2, 8192),
    "ecc_agg": StarkInfo("ECCAggStark", B.StarkId.ECC_AGG, 3339, 12824, 4, 2, 8192),
}


def params_for(stark, config, num_rows=None, flags=0):
    info = STARKS[stark] if isinstance(stark, str) else stark
    rows = num_rows or info.num_rows
    p = B.standard_params(info.stark_id, rows.bit_length() - 1, flags)
    f = config.fri_config
    p.rate_bits, p.cap_height, p.pow_bits = f.rate_bits, f.cap_height, f.proof_of_work_bits
    p.num_query_rounds, p.fri_arity_bits, p.fri_final_poly_bits = f.num_query_rounds, f.reduction_arity_bits, f.final_poly_bits
    p.num_challenges = config.num_challenges
    return p


def prove(ctx, stark, config, trace_poly_values, public_inputs, num_rows=None, flags=0,
          layout=B.TraceLayout.COLMAJOR_U64):
    """starky::prover::prove(stark, &config, trace_poly_values, &public_inputs, &mut timing) on the GPU of `ctx`."""
    p = params_for(stark, config, num_rows, flags)
    return ctx.prove(p, trace_poly_values, public_inputs, layout)


class _Stark:
    """`XStark::<F, D>::new(num_rows)` + `generate_trace(...)`.  generate_trace returns what
    `trace_rows_to_poly_values(stark.generate_trace(...))` yields in the reference (column-major uint64 [COLUMNS, rows],
    from the Python restatement `witness/` -- the independent checker) together with the public-input vector the reference's
    *_main functions build (aggregate_proof.rs:24-227).  generate_trace_rows returns the reference's own shape, the row-major
    `Vec<[F; COLUMNS]>`, as uint32 [rows, COLUMNS] for `TraceLayout.ROWMAJOR_U32`, from the library's C++ generators
    (csrc/witness.cpp: 50 x faster, half the bytes; tests/test_witness_cpp.py compares the two cell for cell)."""
    name = None

    def __init__(self, num_rows=None):
        self.info = STARKS[self.name]
        self.num_rows = num_rows or self.info.num_rows

    @classmethod
    def new(cls, num_rows):
        return cls(num_rows)

    def constraint_degree(self):
        return self.info.constraint_degree


class FP12MulStark(_Stark):
    name = "fp12_mul"

    def generate_trace(self, x, y):                 # fp12_mul.rs:44-48
        from . import witness
        return witness.fp12_mul_trace(x, y, self.num_rows)

    def generate_trace_rows(self, x, y):
        return B.witness_fp12_mul(x, y, self.num_rows)


class PairingPrecompStark(_Stark):
    name = "pairing_precomp"

    def generate_trace(self, x, y, z):              # calc_pairing_precomp.rs:150
        from . import witness
        return witness.pairing_precomp_trace(x, y, z, self.num_rows)

    def generate_trace_rows(self, x, y, z):
        return B.witness_pairing_precomp(x, y, z, self.num_rows)


class MillerLoopStark(_Stark):
    name = "miller_loop"

    def generate_trace(self, x, y, q):              # miller_loop.rs:157 (ell_coeffs derived from the G2 point q)
        from . import witness
        return witness.miller_loop_trace(x, y, q, self.num_rows)

    def generate_trace_rows(self, x, y, q):
        return B.witness_miller_loop(x, y, q, self.num_rows)


class FinalExponentiateStark(_Stark):
    name = "final_exp"

    def generate_trace(self, x):                    # final_exponentiate.rs:246
        from . import witness
        return witness.final_exp_trace(x, self.num_rows)

    def generate_trace_rows(self, x):
        return B.witness_final_exp(x, self.num_rows)


class ECCAggStark(_Stark):
    name = "ecc_agg"

    def generate_trace(self, points, bits):         # ecc_aggregate.rs:37
        from . import witness
        return witness.ecc_aggregate_trace(points, bits, self.num_rows)[:2]

    def generate_trace_rows(self, points, bits):
        return B.witness_ecc_agg(points, bits, self.num_rows)[:2]
