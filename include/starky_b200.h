/*
 * starky_b200.h -- C ABI of libstarkyb200.so, the B200-native (sm_100a) replacement for the
 * data-parallel core of starky::prover::prove as called by Electron-Labs/starky_bls12_381.
 *
 * Drop-in boundary.  The reference has no FFI today; the entry points below are
 * what a Rust `extern "C"` block in a `starky-gpu` crate would bind so that
 *     starky::prover::prove::<F, C, S, 2>(stark, &config, trace_poly_values, &public_inputs, &mut timing)
 * (/root/reference/src/aggregate_proof.rs:59,105,138,169,212 and ecc_aggregate.rs:545) can be served by
 * sb_prove() for F = GoldilocksField, C = PoseidonGoldilocksConfig, D = 2 and S one of the five starks.
 * The Rust shim is in rust/starky_gpu/ (source only: no cargo in the build image); INTEGRATION.md shows it.
 *
 * All field elements are canonical Goldilocks values (uint64_t in [0, 2^64 - 2^32 + 1)).
 * Extension elements F_p[X]/(X^2-7) are two consecutive uint64_t (c0, c1).
 * No function aborts across the boundary: every failure is a negative return code + sb_last_error().
 */
#ifndef STARKY_B200_H
#define STARKY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (reference behaviour: Err(anyhow) / panic, aggregate_proof.rs:65,67) ---- */
#define SB_OK 0
#define SB_EINVAL (-1)                  /* bad argument / shape mismatch with the constraint program      */
#define SB_ECUDA (-2)                   /* CUDA runtime failure (message in sb_last_error)                 */
#define SB_ENCCL (-3)                   /* collective failure: NCCL missing / error, or a peer rank of a group failed */
#define SB_EQUOTIENT_NOT_DIVISIBLE (-4) /* reference: panic "Quotient has failed, the vanishing polynomial is not divisible by Z_H" */
#define SB_EZETA_IN_SUBGROUP (-5)       /* reference: Err("Opening point is in the subgroup.")             */
#define SB_EPOW (-6)                    /* reference: expect("Proof of work failed...")                    */
#define SB_ENOMEM (-7)
#define SB_EAIR (-8)                    /* constraint program missing / malformed                          */

/* ---- stark identity: selects the constraint program (the C side cannot call eval_packed_generic) ---- */
enum sb_stark_id {
  SB_STARK_FP12_MUL = 0,        /* fp12_mul.rs:31            60285 cols, degree 3 */
  SB_STARK_PAIRING_PRECOMP = 1, /* calc_pairing_precomp.rs:135 29376 cols, degree 4 */
  SB_STARK_MILLER_LOOP = 2,     /* miller_loop.rs:81         97330 cols, degree 3 */
  SB_STARK_FINAL_EXP = 3,       /* final_exponentiate.rs:131 73527 cols, degree 5 */
  SB_STARK_ECC_AGG = 4,         /* ecc_aggregate.rs:23        3339 cols, degree 4 */
  SB_STARK_CUSTOM = 100         /* any AIR loaded with sb_air_load (tests)        */
};

/* ---- trace layouts accepted by sb_prove / sb_lde_commit ---- */
enum sb_trace_layout {
  SB_TRACE_COLMAJOR_U64 = 0, /* one [n_cols][n] block; column c at trace + c*n            (Vec<PolynomialValues<F>> flattened) */
  SB_TRACE_COLS_U64_PTRS = 1, /* const uint64_t* const* : n_cols pointers to n values each (Vec<PolynomialValues<F>> as is, aggregate_proof.rs:57) */
  SB_TRACE_ROWMAJOR_U64 = 2, /* [n][n_cols], the Vec<[F; COLUMNS]> that generate_trace returns (skips trace_rows_to_poly_values) */
  SB_TRACE_ROWMAJOR_U32 = 3, /* [n][n_cols] uint32_t: every cell the reference writes is < 2^32 (utils.rs:7-19) */
  SB_TRACE_DEVICE_COLMAJOR_U64 = 4 /* as COLMAJOR_U64 but `trace` is a device pointer on the ctx's GPU */
};

/* ---- flags ---- */
#define SB_FLAG_ALLOW_INVALID_TRACE 1u /* benchmarking on random traces: drop (instead of rejecting) non-zero high quotient coefficients */
#define SB_FLAG_FIXED_POW_WITNESS 2u   /* use params.fixed_pow_witness instead of grinding (transcript-exact replay of an oracle/reference run) */
#define SB_FLAG_OBSERVE_PUBLIC_INPUTS 4u /* observe PIs before the trace cap (off = the pinned plonky2 era) */
#define SB_FLAG_FRI_MUL_BY_X 8u        /* older plonky2 multiplied the FRI final polynomial by X (off by default) */

/* Mirrors starky::config::StarkConfig + the per-stark constants (aggregate_proof.rs:32-33,76,122,155-156,186-187). */
typedef struct sb_params {
  uint32_t stark_id;          /* enum sb_stark_id                                              */
  uint32_t log_n;             /* trace rows n = 2^log_n                                        */
  uint32_t n_cols;            /* S::COLUMNS                                                    */
  uint32_t n_public_inputs;   /* S::PUBLIC_INPUTS                                              */
  uint32_t constraint_degree; /* Stark::constraint_degree()                                    */
  uint32_t rate_bits;         /* config.fri_config.rate_bits (1, or 2 for PP/FE/ECC)           */
  uint32_t cap_height;        /* 4                                                             */
  uint32_t num_challenges;    /* 2                                                             */
  uint32_t pow_bits;          /* 16                                                            */
  uint32_t num_query_rounds;  /* 84                                                            */
  uint32_t fri_arity_bits;    /* ConstantArityBits(4, 5): arity bits                           */
  uint32_t fri_final_poly_bits; /* ... and final poly bits                                     */
  uint32_t flags;
  uint32_t reserved;
  uint64_t fixed_pow_witness;
} sb_params;

/* StarkConfig::standard_fast_config() with the per-stark rate_bits override; fills everything but log_n/flags. */
int sb_params_standard(uint32_t stark_id, uint32_t log_n, sb_params* out);

/* ---- proof object: a flat uint64_t buffer (POD) with the field order of
 *      starky::proof::StarkProofWithPublicInputs<F, C, 2> (SURVEY 8b "Proof object to fill") ---- */
typedef struct sb_proof_layout {
  uint32_t log_n, log_lde, n_cols, n_quotient_polys, n_public_inputs, cap_len, n_fri_rounds, final_poly_len;
  uint32_t n_queries, arity_bits, trace_path_len, reserved;
  uint64_t off_trace_cap;      /* [cap_len][4]                                                    */
  uint64_t off_quotient_cap;   /* [cap_len][4]                                                    */
  uint64_t off_local_values;   /* openings.local_values   [n_cols][2]                            */
  uint64_t off_next_values;    /* openings.next_values    [n_cols][2]                            */
  uint64_t off_quotient_polys; /* openings.quotient_polys [nq][2]                                */
  uint64_t off_fri_caps;       /* commit_phase_merkle_caps [n_fri_rounds][cap_len][4]            */
  uint64_t off_final_poly;     /* final_poly [final_poly_len][2]                                 */
  uint64_t off_pow_witness;    /* 1                                                               */
  uint64_t off_queries;        /* n_queries records of query_stride words                        */
  uint64_t query_stride;
  /* inside one query record: */
  uint64_t q_off_trace_leaf;   /* [n_cols]                                                        */
  uint64_t q_off_trace_path;   /* [trace_path_len][4]                                             */
  uint64_t q_off_quot_leaf;    /* [nq]                                                            */
  uint64_t q_off_quot_path;    /* [trace_path_len][4]                                             */
  uint64_t q_off_steps;        /* round r: evals [2^arity][2] then path [step_path_len(r)][4], packed */
  uint64_t off_public_inputs;  /* [n_public_inputs]                                               */
  uint64_t total_words;
} sb_proof_layout;

/* Computes the layout for the given parameters (pure host arithmetic). */
int sb_proof_layout_for(const sb_params* p, sb_proof_layout* out);
/* Merkle path length of FRI round r's tree in that layout. */
uint32_t sb_fri_step_path_len(const sb_proof_layout* l, uint32_t round);
/* Word offset, inside one query record, of FRI round r's step. */
uint64_t sb_fri_step_offset(const sb_proof_layout* l, uint32_t round);

typedef struct sb_ctx sb_ctx;
typedef struct sb_proof {
  sb_proof_layout layout;
  uint64_t* words; /* layout.total_words canonical u64, owned by the library */
  /* stage timings of this proof, milliseconds, CUDA events (TimingTree scopes of starky::prover::prove) */
  float ms_h2d, ms_trace_commit, ms_quotient, ms_quotient_commit, ms_openings, ms_fri, ms_d2h, ms_total;
} sb_proof;

/* ---- lifecycle ---- */
int sb_init(const int* devices, int n_devices, sb_ctx** out); /* devices == NULL: current device; n_devices > 1: one ctx over
                                                                 several GPUs, sb_prove shards the trace (see "multi-GPU groups") */
void sb_destroy(sb_ctx* ctx);
const char* sb_last_error(sb_ctx* ctx); /* ctx may be NULL: last error of this thread */

/* ---- constraint programs ("AIR") ---- */
/* Load a compiled constraint program (tools/airgen output) and bind it to a stark id on this ctx.
 * The five standard programs are linked into the library (csrc/air_blobs.S) and bound on first use; when $SB_AIR_DIR
 * is set, <dir>/<name>.airbin (fp12_mul, pairing_precomp, miller_loop, final_exp, ecc_agg) is read instead. */
int sb_air_load(sb_ctx* ctx, uint32_t stark_id, const char* path);
/* The run form the loader builds from a program image (groups reordered by column locality, bodies folded into run records,
 * weight slots permuted; csrc/quotient.cu: translate_runs), for inspection and for the CPU test that emulates it.  Host
 * only.  info_out[8] = {run-form words, weight slots, K, groups, runs, bodies in runs, bodies, columns}; the other outputs
 * are optional (code2_out == NULL: sizes only). */
int sb_air_run_form(const void* image, size_t image_len, uint64_t* code2_out, size_t code2_cap, uint32_t* slot_off2_out,
                    uint32_t* slot_ks2_out, uint32_t* group_pc2_out, uint32_t* group_slot2_out, uint32_t* info_out);

/* ---- the hot path: replaces starky::prover::prove ---- */
int sb_prove(sb_ctx* ctx, const sb_params* p, const void* trace, int layout,
             const uint64_t* public_inputs, sb_proof** out);
void sb_proof_free(sb_proof* proof);

/* ---- the proofs of one job list through one call.  The reference proves the seven starky proofs of one BLS
 *      signature verification one after the other (aggregate_proof.rs:279-370: 2 x PairingPrecomp, 2 x MillerLoop,
 *      FP12Mul, FinalExp, ECCAgg); they are independent.  sb_prove_batch runs the jobs on the given contexts (any mix of
 *      devices, several contexts per device allowed, multi-device contexts allowed), one internal host thread per context,
 *      longest job first: few-leaf (latency-bound) proofs share a GPU up to its number of contexts, many-leaf
 *      (throughput-bound) proofs get an idle GPU to themselves.  On return jobs[i].proof / rc / ms are filled in job order;
 *      the return value is the first failing job's code (0 if all proofs were produced). ---- */
typedef struct sb_job {
  sb_params params;
  const void* trace;             /* as for sb_prove */
  int layout;                    /* enum sb_trace_layout */
  const uint64_t* public_inputs;
  sb_proof* proof;               /* out: free with sb_proof_free */
  int rc;                        /* out: SB_OK or the error of this job */
  float ms;                      /* out: wall milliseconds of this job's sb_prove */
} sb_job;
int sb_prove_batch(sb_ctx* const* ctxs, int n_ctx, sb_job* jobs, int n_jobs);

/* ---- witness generation: the reference's generate_trace in C++, so that the caller hands over a few field
 *      elements instead of a multi-GB trace.  Cells are row-major uint32_t (SB_TRACE_ROWMAJOR_U32: every cell the reference
 *      writes is a u32 limb, a carry or a bit -- utils.rs:7-19), half the PCIe bytes of the u64 layouts.  Host code.
 *      sb_witness_fp12_mul: FP12MulStark::generate_trace (fp12_mul.rs:44-48; fill_trace_fp12_multiplication fp12.rs:186-232
 *      and the fp / fp2 / fp6 gadgets below it) + the public inputs of fp12_mul_main (aggregate_proof.rs:124-151).
 *      x, y: Fp12 operands as 12 x 12 little-endian u32 limbs ([Fp; 12], Fp = [u32; 12]), each coefficient < p.
 *      trace_out: [num_rows][60285] uint32_t; public_inputs_out: 432 values (x ++ y ++ x*y).
 *      sb_prove_fp12_mul: both steps in one call (generate on the host, prove on the ctx's GPU). ---- */
int sb_witness_fp12_mul(const uint32_t* x, const uint32_t* y, uint32_t num_rows, uint32_t* trace_out, uint64_t* public_inputs_out);
int sb_prove_fp12_mul(sb_ctx* ctx, const sb_params* p, const uint32_t* x, const uint32_t* y, sb_proof** out);
/*      sb_witness_ecc_agg: ECCAggStark::generate_trace (ecc_aggregate.rs:37-82; fill_trace_g1_addition g1.rs:26-255) + the
 *      public inputs of ec_aggregate_main (aggregate_proof.rs:186-227).  points: 512 affine G1 points, x ++ y as 12
 *      little-endian u32 limbs each ([512][24]); bits: 512 participation flags (0 / 1).  trace_out: [num_rows][3339]
 *      uint32_t; public_inputs_out: 12 824 values (points ++ bits ++ aggregate); result_out (may be NULL): the aggregate
 *      point, 24 limbs.  sb_prove_ecc_agg: both steps in one call. */
int sb_witness_ecc_agg(const uint32_t* points, const uint8_t* bits, uint32_t num_rows, uint32_t* trace_out,
                       uint64_t* public_inputs_out, uint32_t* result_out);
int sb_prove_ecc_agg(sb_ctx* ctx, const sb_params* p, const uint32_t* points, const uint8_t* bits, sb_proof** out);
/*      sb_witness_pairing_precomp: PairingPrecompStark::generate_trace (calc_pairing_precomp.rs:150-366) + the public inputs of
 *      calc_pairing_precomp_main (aggregate_proof.rs:24-69).  q: the projective G2 point x ++ y ++ z, each an Fp2 of 2 x 12
 *      little-endian u32 limbs ([3][24]).  trace_out: [num_rows][29376]; public_inputs_out: 4968 values.
 *      sb_witness_miller_loop: MillerLoopStark::generate_trace (miller_loop.rs:87-160) + miller_loop_main's public inputs
 *      (aggregate_proof.rs:71-121).  g1: the affine G1 point x ++ y ([24]); q as above.  trace_out: [num_rows][97330];
 *      public_inputs_out: 5064 values (point, 68 x 3 line coefficients, result).
 *      sb_witness_final_exp: FinalExponentiateStark::generate_trace (final_exponentiate.rs:137-281) + final_exponentiate_main's
 *      public inputs (aggregate_proof.rs:153-184).  x: the Fp12 input ([12][12]).  num_rows = 8192; trace_out:
 *      [8192][73527] (2.4 GB); public_inputs_out: 288 values (x, result).  sb_prove_*: both steps in one call. */
int sb_witness_pairing_precomp(const uint32_t* q, uint32_t num_rows, uint32_t* trace_out, uint64_t* public_inputs_out);
int sb_witness_miller_loop(const uint32_t* g1, const uint32_t* q, uint32_t num_rows, uint32_t* trace_out, uint64_t* public_inputs_out);
int sb_witness_final_exp(const uint32_t* x, uint32_t num_rows, uint32_t* trace_out, uint64_t* public_inputs_out);
int sb_prove_pairing_precomp(sb_ctx* ctx, const sb_params* p, const uint32_t* q, sb_proof** out);
int sb_prove_miller_loop(sb_ctx* ctx, const sb_params* p, const uint32_t* g1, const uint32_t* q, sb_proof** out);
int sb_prove_final_exp(sb_ctx* ctx, const sb_params* p, const uint32_t* x, sb_proof** out);
const char* sb_witness_last_error(void);

/* ---- proof wire formats: the proof as bytes for a consumer that does not link this library -- the
 *      reference's verify_stark_proof / recursive verifier take a starky::proof::StarkProofWithPublicInputs<F, C, 2>
 *      (aggregate_proof.rs:67,113,146,177,220 and :435-439).  Host code only: works in a process without a GPU. ---- */
enum sb_wire_format {
  SB_WIRE_POD = 0,            /* "SBPROOF1" | sb_params | total_words u64 | words (LE u64): self-describing flat POD          */
  SB_WIRE_PLONKY2_BUFFER = 1, /* the fields of StarkProofWithPublicInputs in declaration order through the primitives of
                                 plonky2::util::serialization::Write (write_merkle_cap, write_field_ext_vec, write_fri_proof
                                 with u8-prefixed Merkle proofs, write_field_vec); no length prefixes: read with the params */
  SB_WIRE_SERDE_JSON = 2      /* serde_json of the same struct (FriProof / MerkleCap / MerkleProof exactly as their
                                 #[derive(Serialize)] print them); write-only                                              */
};
/* buf == NULL: only computes *len_out.  p may be NULL for the two plonky2 formats (the JSON then omits "config"). */
int sb_proof_serialize(const sb_proof* proof, const sb_params* p, int format, void* buf, size_t cap, size_t* len_out);
/* SB_WIRE_POD: p optional (checked against the image when given), the image's params are returned in params_out (optional);
 * SB_WIRE_PLONKY2_BUFFER: p required.  Rejects truncated / oversized images and non-canonical field elements. */
int sb_proof_deserialize(const void* buf, size_t len, int format, const sb_params* p, sb_params* params_out, sb_proof** out);
/* A proof object from n_words raw POD words of the layout `p` defines (copied). */
int sb_proof_from_words(const sb_params* p, const uint64_t* words, size_t n_words, sb_proof** out);

/* ---- stage-level entry points (parity tests and benchmarks; SURVEY 8b "Stage-level exports") ----
 * Device buffers stay resident in the ctx between stage calls of one proof. */

/* PolynomialBatch::from_values: iNTT, coset LDE (shift 7), Poseidon Merkle tree to the cap.
 * lde_out (optional, host): [n_cols][N] values in LEAF order (index = bitrev(natural LDE index)),
 * digests_out (optional, host): [N][4] leaf digests, cap_out (optional, host): [2^cap_height][4]. */
int sb_lde_commit(sb_ctx* ctx, const sb_params* p, const void* trace, int layout,
                  uint64_t* lde_out, uint64_t* digests_out, uint64_t* cap_out);
/* compute_quotient_polys up to (excluding) coset_ifft: q_j(x_i) for both challenges.
 * Needs a preceding sb_lde_commit on this ctx.  out (host): [num_challenges][N] in NATURAL LDE index order. */
int sb_quotient_values(sb_ctx* ctx, const sb_params* p, const uint64_t* public_inputs,
                       const uint64_t* alphas, uint64_t* out);
/* StarkOpeningSet::new on the trace committed by the preceding sb_lde_commit on this ctx: local_out[c] = P_c(zeta),
 * next_out[c] = P_c(g zeta), g = primitive_root_of_unity(log_n); extension elements as (c0, c1), host [n_cols][2] each. */
int sb_openings(sb_ctx* ctx, const sb_params* p, const uint64_t* zeta, uint64_t* local_out, uint64_t* next_out);
/* fri_committed_trees on a given polynomial: coeffs = host [n][2] extension coefficients in natural order (the final_poly
 * of PolynomialBatch::prove_openings), betas = [n_fri_rounds][2] folding challenges in place of the transcript.
 * caps_out: [n_fri_rounds][2^cap_height][4], final_poly_out: [final_poly_len][2] (sb_proof_layout_for gives both counts). */
int sb_fri_commit(sb_ctx* ctx, const sb_params* p, const uint64_t* coeffs, const uint64_t* betas, uint64_t* caps_out,
                  uint64_t* final_poly_out);
/* Batched forward/inverse NTT of `count` length-2^log_n vectors (natural order in and out). */
int sb_ntt_batch(sb_ctx* ctx, uint64_t* data, uint32_t log_n, uint32_t count, int inverse);
/* `count` Poseidon-12 permutations on host states [count][12]. */
int sb_poseidon_permute_batch(sb_ctx* ctx, uint64_t* states, uint32_t count);
/* hash_or_noop of `count` leaves of `leaf_len` elements: leaves given column-major [leaf_len][count]. */
int sb_hash_leaves(sb_ctx* ctx, const uint64_t* cols, uint32_t leaf_len, uint32_t count, uint64_t* digests_out);

/* Coefficients of the last committed trace, [n_cols][n] in bit-reversed coefficient order (parity tests). */
int sb_coeffs_download(sb_ctx* ctx, uint64_t* coeffs_out);

/* ---- multi-GPU stage entry points.  Device pointers in and out, work is queued on the ctx
 *      stream; the exchange between them -- an all-to-all of LDE slabs, an all-gather of 32-byte digests -- is done by
 *      the host with NCCL (torch.distributed in the harness, ncclSend/ncclRecv in a Rust/C++ host).
 *      Phase 1, column-sharded: rank g runs K1 on its column slice and writes the LDE as n_row_blocks slabs
 *        d_lde_out[b][c][N / n_row_blocks], slab b = the positions rank b will hash (replaces the per-column
 *        ifft/lde/coset_fft of PolynomialBatch::from_values; aggregate_proof.rs:59,105,138,169,212).
 *      Phase 2, row-sharded: after the all-to-all rank g holds [n_cols][N / G] and hashes its leaves (K2); the
 *        position-ordered digests of all ranks are gathered and the tree is built to the cap (K3). ---- */
int sb_lde_cols_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_trace, uint32_t n_cols_local,
                       uint32_t n_row_blocks, uint64_t* d_coeffs_out /* optional */, uint64_t* d_lde_out);
/*      K1 fused with the exchange ("P2P stores fused into the K1 epilogue"): the same transform, but every LDE
 *        value is stored straight into the row buffer of the rank that owns its row block -- peer_rows[b] (host array of
 *        n_row_blocks device pointers, peer-mapped over NVLink) = rank b's [n_cols total][N / n_row_blocks] buffer; this
 *        rank's columns are first_col .. first_col + n_cols_local - 1.  There is no all-to-all afterwards. */
int sb_lde_cols_peer_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_trace, uint32_t n_cols_local,
                            uint32_t n_row_blocks, uint32_t first_col, uint64_t* d_coeffs_out /* optional */,
                            const uint64_t* peer_rows);
int sb_hash_rows_device(sb_ctx* ctx, const uint64_t* d_cols, uint32_t leaf_len, uint32_t n_leaves, uint64_t* d_digests);
int sb_merkle_from_position_digests(sb_ctx* ctx, const sb_params* p, const uint64_t* d_digests_pos, uint64_t* cap_out);
/*      Phase 2 continued: the quotient values q_j(x) (j < 2) of this rank's row block: d_rows = [n_cols][rows_per_block]
 *        holding LDE positions block_index * rows_per_block ..., d_halo_next_row = the n_cols values of the row after
 *        the block's last position when that row lives on another rank (rows_per_block < n; NULL otherwise),
 *        d_out = [2][rows_per_block] on the device (this rank's share of starky::prover::compute_quotient_polys).
 *        sb_transcript_alphas: observe the gathered trace cap, draw the alphas (every rank gets the same values). */
int sb_quotient_rows_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_rows, uint32_t rows_per_block,
                            uint32_t block_index, const uint64_t* d_halo_next_row, const uint64_t* public_inputs,
                            const uint64_t* alphas, uint64_t* d_out);
int sb_transcript_alphas(const uint64_t* trace_cap, uint32_t cap_len, uint32_t num_challenges, uint64_t* alphas_out);
/*      The tail of a sharded proof (SURVEY 8e "small collectives").  sb_prove_sharded runs the SAME orchestration as
 *      sb_prove on every rank (same transcript, same proof on every rank) but asks the host, through five hooks, for the
 *      five things that are distributed: the hooks are collective (every rank calls them in the same order) and are
 *      implemented with the per-rank stage functions above and below plus NCCL (starky_bls12_381_b200/sharded.py).
 *        commit     : the sharded trace commitment; on return the ctx holds the trace Merkle tree
 *                     (sb_merkle_from_position_digests) and cap_out the cap
 *        quotient   : q_j(x) at all N positions, device [2][N], position order
 *        openings   : P_c(zeta), P_c(g zeta) of all n_cols trace columns, host [n_cols][2] each (column-sharded
 *                     coefficient slices, all-gathered)
 *        combine    : sum_c alpha^c coeffs_c over all trace columns, device [n][2], bit-reversed coefficient order
 *                     (per-rank partial sums with the rank's alpha-power offset, all-gathered and added)
 *        query_rows : the full trace rows at `count` device LDE positions, device [count][n_cols] (from the row owners)
 *      Every hook returns 0 or an SB_E* code.  Everything else (quotient commitment, transcript, FRI rounds, proof of
 *      work, Merkle paths) is small and is computed redundantly on every rank. */
typedef struct sb_shard_hooks {
  void* user;
  int (*commit)(void* user, uint64_t* cap_out);
  int (*quotient)(void* user, const uint64_t* alphas, uint64_t* d_q_out);
  int (*openings)(void* user, const uint64_t* zeta, const uint64_t* zeta_next, uint64_t* local_out, uint64_t* next_out);
  int (*combine)(void* user, const uint64_t* alpha, uint64_t* d_out);
  int (*query_rows)(void* user, const uint32_t* positions, uint32_t count, uint64_t* d_rows_out);
} sb_shard_hooks;
int sb_prove_sharded(sb_ctx* ctx, const sb_params* p, const sb_shard_hooks* hooks, const uint64_t* public_inputs,
                     sb_proof** out);
/*      Per-rank pieces for the hooks: openings and alpha-weighted sum of a column slice of coefficients
 *      (d_coeffs = [n_cols_local][n], bit-reversed coefficient order, as sb_lde_cols_device leaves them). */
int sb_openings_cols_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_coeffs, uint32_t n_cols_local,
                            const uint64_t* zeta, const uint64_t* zeta_next, uint64_t* local_out, uint64_t* next_out);
int sb_combine_cols_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_coeffs, uint32_t n_cols_local,
                           const uint64_t* alpha, uint32_t first_col, uint64_t* d_out);
int sb_memcpy_device(sb_ctx* ctx, void* d_dst, const void* d_src, uint64_t bytes);
int sb_synchronize(sb_ctx* ctx);

/* ---- multi-GPU groups: one proof with the trace sharded over the GPUs of one box, entirely inside the
 *      library.  Replaces the same starky::prover::prove call (aggregate_proof.rs:59,105,138,169,212); the host only
 *      decides which GPUs form a group.
 *      Column-sharded K1 stores the LDE straight into the row buffers of the ranks that own the row blocks (peer memory
 *      over NVLink: CUDA IPC mappings between processes, peer access inside one process); leaf hashing and the quotient
 *      run row-local; digests, halo rows, quotient values, openings, FRI combine partials (added mod p on the device)
 *      and the 84 query rows are exchanged with NCCL (one process per GPU) or with peer copies (one process, several
 *      GPUs).  Every rank ends with the same proof.
 *
 *      One process per GPU:   rank 0 calls sb_group_unique_id and hands the 128 bytes to the other ranks over any channel
 *                             (MPI, a file, torch.distributed); every rank then calls sb_group_init_rank (collective).
 *                             NCCL is loaded with dlopen("libnccl.so.2") at that point; SB_ENCCL if it is missing.
 *      One process, n GPUs:   sb_init(devices, n > 1, &ctx) -- sb_prove(ctx, ...) then cuts a host trace
 *                             (SB_TRACE_COLMAJOR_U64 or SB_TRACE_COLS_U64_PTRS) into column slices and proves on all n
 *                             devices, one internal host thread per device; or sb_group_init_local over contexts the
 *                             caller made, with sb_group_prove called from one thread per rank. ---- */
typedef struct sb_group sb_group;
#define SB_GROUP_UNIQUE_ID_BYTES 128
#define SB_GROUP_NO_FUSED 1u   /* sb_group_prove flag: all-to-all after K1 instead of K1 storing into peer memory */
int sb_group_unique_id(uint8_t id[SB_GROUP_UNIQUE_ID_BYTES]);
int sb_group_init_rank(sb_ctx* ctx, int rank, int world, const uint8_t id[SB_GROUP_UNIQUE_ID_BYTES], sb_group** out);
int sb_group_init_local(sb_ctx* const* ctxs, int world, sb_group** out /* [world] */);
void sb_group_destroy(sb_group* g);
int sb_group_rank(const sb_group* g);
int sb_group_size(const sb_group* g);
/* The columns rank `rank` of a `world`-rank group owns ([first_col, first_col + n_cols_local), ragged: 97330 = 2 x 12167 +
 * 6 x 12166) and the LDE positions per rank; pure host arithmetic.  sb_group_column_slice: the same for this rank. */
int sb_shard_columns(const sb_params* p, uint32_t world, uint32_t rank, uint32_t* first_col, uint32_t* n_cols_local,
                     uint32_t* rows_per_rank);
int sb_group_column_slice(const sb_group* g, const sb_params* p, uint32_t* first_col, uint32_t* n_cols_local);
/* Collective: every rank of the group calls it with ITS column slice, local_trace = [n_cols_local][n] u64 column-major in
 * host memory (on_device = 0; copied inside the call) or on this rank's GPU (on_device = 1), and the same public inputs.
 * Every rank receives the same proof (free with sb_proof_free). */
int sb_group_prove(sb_group* g, const sb_params* p, const void* local_trace, int on_device, const uint64_t* public_inputs,
                   uint32_t flags, sb_proof** out);
/* Wall milliseconds this rank spent in the phases of the last sb_group_prove ("commit", "quotient", "openings",
 * "combine", "query_rows"; includes waiting for peers); "fused" = 1 if K1 stored into peer memory.  <0 if unknown. */
float sb_group_phase_ms(const sb_group* g, const char* phase);

/* The host-side transcript permutation (plonky2 Challenger, run between kernels): variant 0 = portable scalar,
 * 1 = AVX2, 2 = AVX-512, 3 = AVX-512 full rounds + sparse partial rounds, 4 / 5 = AVX-512 IFMA + BMI2 hybrid (words 0..7
 * in a vector, 8..11 on mulx; look-ahead of one / two partial rounds); the library picks the best one the CPU supports.
 * Returns 1 if that variant ran, 0 if the CPU lacks the extension.  Parity-test hook. */
int sb_host_poseidon_permute_variant(uint64_t state[12], int variant);

/* ---- device-resident benchmarking hooks (bench.py `value` leg: inputs already in HBM) ---- */
/* Upload a trace once; subsequent sb_prove(..., trace=NULL, layout=SB_TRACE_DEVICE_COLMAJOR_U64) re-uses it. */
int sb_trace_upload(sb_ctx* ctx, const sb_params* p, const void* trace, int layout);
/* Number of kernels this ctx has launched so far (bench.py gpu_launches). */
uint64_t sb_kernel_launches(sb_ctx* ctx);
/* Milliseconds (CUDA events, on the ctx stream) of named stages of the last sb_prove / stage call:
 * "lde", "leaf_hash", "merkle", "quotient", ... returns <0 if unknown. */
float sb_stage_ms(sb_ctx* ctx, const char* stage);
/* dependent-free u32 multiply-add throughput microbenchmark (Gop/s), the IMAD roofline denominator. */
int sb_measure_imad_peak(sb_ctx* ctx, double* gops_out);

#ifdef __cplusplus
}
#endif
#endif /* STARKY_B200_H */
