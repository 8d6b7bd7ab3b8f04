//! `extern "C"` block for include/starky_b200.h (field-for-field; keep in sync with the header).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const SB_OK: c_int = 0;
pub const SB_EQUOTIENT_NOT_DIVISIBLE: c_int = -4;
pub const SB_EZETA_IN_SUBGROUP: c_int = -5;

pub const SB_TRACE_COLS_U64_PTRS: c_int = 1;
pub const SB_TRACE_ROWMAJOR_U64: c_int = 2;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct sb_params {
    pub stark_id: u32, pub log_n: u32, pub n_cols: u32, pub n_public_inputs: u32, pub constraint_degree: u32,
    pub rate_bits: u32, pub cap_height: u32, pub num_challenges: u32, pub pow_bits: u32, pub num_query_rounds: u32,
    pub fri_arity_bits: u32, pub fri_final_poly_bits: u32, pub flags: u32, pub reserved: u32,
    pub fixed_pow_witness: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct sb_proof_layout {
    pub log_n: u32, pub log_lde: u32, pub n_cols: u32, pub n_quotient_polys: u32, pub n_public_inputs: u32,
    pub cap_len: u32, pub n_fri_rounds: u32, pub final_poly_len: u32, pub n_queries: u32, pub arity_bits: u32,
    pub trace_path_len: u32, pub reserved: u32,
    pub off_trace_cap: u64, pub off_quotient_cap: u64, pub off_local_values: u64, pub off_next_values: u64,
    pub off_quotient_polys: u64, pub off_fri_caps: u64, pub off_final_poly: u64, pub off_pow_witness: u64,
    pub off_queries: u64, pub query_stride: u64, pub q_off_trace_leaf: u64, pub q_off_trace_path: u64,
    pub q_off_quot_leaf: u64, pub q_off_quot_path: u64, pub q_off_steps: u64, pub off_public_inputs: u64,
    pub total_words: u64,
}

#[repr(C)]
pub struct sb_proof {
    pub layout: sb_proof_layout,
    pub words: *mut u64,
    pub ms_h2d: f32, pub ms_trace_commit: f32, pub ms_quotient: f32, pub ms_quotient_commit: f32,
    pub ms_openings: f32, pub ms_fri: f32, pub ms_d2h: f32, pub ms_total: f32,
}

#[repr(C)]
pub struct sb_ctx { _private: [u8; 0] }

/// Binary-identical to `sb_shard_hooks`: five collective callbacks of a sharded proof.
#[repr(C)]
pub struct sb_shard_hooks {
    pub user: *mut c_void,
    pub commit: Option<unsafe extern "C" fn(user: *mut c_void, cap_out: *mut u64) -> c_int>,
    pub quotient: Option<unsafe extern "C" fn(user: *mut c_void, alphas: *const u64, d_q_out: *mut u64) -> c_int>,
    pub openings: Option<unsafe extern "C" fn(user: *mut c_void, zeta: *const u64, zeta_next: *const u64,
                                              local_out: *mut u64, next_out: *mut u64) -> c_int>,
    pub combine: Option<unsafe extern "C" fn(user: *mut c_void, alpha: *const u64, d_out: *mut u64) -> c_int>,
    pub query_rows: Option<unsafe extern "C" fn(user: *mut c_void, positions: *const u32, count: u32, d_rows_out: *mut u64) -> c_int>,
}

#[cfg(feature = "gpu")]
extern "C" {
    pub fn sb_init(devices: *const c_int, n_devices: c_int, out: *mut *mut sb_ctx) -> c_int;
    pub fn sb_destroy(ctx: *mut sb_ctx);
    pub fn sb_last_error(ctx: *mut sb_ctx) -> *const c_char;
    pub fn sb_params_standard(stark_id: u32, log_n: u32, out: *mut sb_params) -> c_int;
    pub fn sb_prove(ctx: *mut sb_ctx, p: *const sb_params, trace: *const c_void, layout: c_int,
                    public_inputs: *const u64, out: *mut *mut sb_proof) -> c_int;
    pub fn sb_proof_free(proof: *mut sb_proof);
    // one proof sharded over the GPUs of a box (include/starky_b200.h: sb_shard_hooks); the hooks are collective and
    // are implemented by the host with ncclSend/ncclRecv (the Python harness does them with torch.distributed)
    pub fn sb_prove_sharded(ctx: *mut sb_ctx, p: *const sb_params, hooks: *const sb_shard_hooks,
                            public_inputs: *const u64, out: *mut *mut sb_proof) -> c_int;
    pub fn sb_lde_cols_device(ctx: *mut sb_ctx, p: *const sb_params, d_trace: *const u64, n_cols_local: u32,
                              n_row_blocks: u32, d_coeffs_out: *mut u64, d_lde_out: *mut u64) -> c_int;
    pub fn sb_hash_rows_device(ctx: *mut sb_ctx, d_cols: *const u64, leaf_len: u32, n_leaves: u32, d_digests: *mut u64) -> c_int;
    pub fn sb_merkle_from_position_digests(ctx: *mut sb_ctx, p: *const sb_params, d_digests_pos: *const u64, cap_out: *mut u64) -> c_int;
    pub fn sb_transcript_alphas(trace_cap: *const u64, cap_len: u32, num_challenges: u32, alphas_out: *mut u64) -> c_int;
    pub fn sb_quotient_rows_device(ctx: *mut sb_ctx, p: *const sb_params, d_rows: *const u64, rows_per_block: u32,
                                   block_index: u32, d_halo_next_row: *const u64, public_inputs: *const u64,
                                   alphas: *const u64, d_out: *mut u64) -> c_int;
    pub fn sb_openings_cols_device(ctx: *mut sb_ctx, p: *const sb_params, d_coeffs: *const u64, n_cols_local: u32,
                                   zeta: *const u64, zeta_next: *const u64, local_out: *mut u64, next_out: *mut u64) -> c_int;
    pub fn sb_combine_cols_device(ctx: *mut sb_ctx, p: *const sb_params, d_coeffs: *const u64, n_cols_local: u32,
                                  alpha: *const u64, first_col: u32, d_out: *mut u64) -> c_int;
    pub fn sb_fri_step_path_len(l: *const sb_proof_layout, round: u32) -> u32;
    pub fn sb_fri_step_offset(l: *const sb_proof_layout, round: u32) -> u64;
    pub fn sb_proof_layout_for(p: *const sb_params, out: *mut sb_proof_layout) -> c_int;
    // proof wire formats (include/starky_b200.h: enum sb_wire_format); host code, no GPU needed
    pub fn sb_proof_serialize(proof: *const sb_proof, p: *const sb_params, format: c_int, buf: *mut c_void, cap: usize,
                              len_out: *mut usize) -> c_int;
    pub fn sb_proof_deserialize(buf: *const c_void, len: usize, format: c_int, p: *const sb_params,
                                params_out: *mut sb_params, out: *mut *mut sb_proof) -> c_int;
    // multi-GPU groups, one process per GPU (a single process uses sb_init with several devices: GpuProver::new_multi)
    pub fn sb_group_unique_id(id: *mut u8) -> c_int;
    pub fn sb_group_init_rank(ctx: *mut sb_ctx, rank: c_int, world: c_int, id: *const u8, out: *mut *mut sb_group) -> c_int;
    pub fn sb_group_destroy(g: *mut sb_group);
    pub fn sb_group_column_slice(g: *const sb_group, p: *const sb_params, first_col: *mut u32, n_cols_local: *mut u32) -> c_int;
    pub fn sb_group_prove(g: *mut sb_group, p: *const sb_params, local_trace: *const c_void, on_device: c_int,
                          public_inputs: *const u64, flags: u32, out: *mut *mut sb_proof) -> c_int;
    // the seven proofs of one BLS verification (aggregate_proof.rs:279-370) through one call
    pub fn sb_prove_batch(ctxs: *const *mut sb_ctx, n_ctx: c_int, jobs: *mut sb_job, n_jobs: c_int) -> c_int;
    // generate_trace in C++ + prove, from the operands of the *_main functions (aggregate_proof.rs:24-227); limbs are the
    // reference's Fp = [u32; 12], little-endian
    pub fn sb_prove_fp12_mul(ctx: *mut sb_ctx, p: *const sb_params, x: *const u32, y: *const u32, out: *mut *mut sb_proof) -> c_int;
    pub fn sb_prove_ecc_agg(ctx: *mut sb_ctx, p: *const sb_params, points: *const u32, bits: *const u8, out: *mut *mut sb_proof) -> c_int;
    pub fn sb_prove_pairing_precomp(ctx: *mut sb_ctx, p: *const sb_params, q: *const u32, out: *mut *mut sb_proof) -> c_int;
    pub fn sb_prove_miller_loop(ctx: *mut sb_ctx, p: *const sb_params, g1: *const u32, q: *const u32, out: *mut *mut sb_proof) -> c_int;
    pub fn sb_prove_final_exp(ctx: *mut sb_ctx, p: *const sb_params, x: *const u32, out: *mut *mut sb_proof) -> c_int;
}

#[repr(C)]
pub struct sb_group { _private: [u8; 0] }

/// Binary-identical to `sb_job` (include/starky_b200.h).
#[repr(C)]
pub struct sb_job {
    pub params: sb_params,
    pub trace: *const c_void,
    pub layout: c_int,
    pub public_inputs: *const u64,
    pub proof: *mut sb_proof,
    pub rc: c_int,
    pub ms: f32,
}
