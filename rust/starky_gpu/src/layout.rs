//! `sb_proof_layout` of include/starky_b200.h computed in Rust (the same arithmetic as csrc/prover.cu: proof_layout),
//! so that a serialized proof can be read without linking libstarkyb200.  Field order = the order in which
//! starky::proof::StarkProofWithPublicInputs<F, C, 2> declares its fields.
use crate::ffi::{sb_params, sb_proof_layout};

/// FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits).reduction_arity_bits(...)
pub fn fri_arities(p: &sb_params) -> Vec<u32> {
    let mut r = Vec::new();
    let mut db = p.log_n;
    while db > p.fri_final_poly_bits && db + p.rate_bits - p.fri_arity_bits >= p.cap_height {
        assert!(db >= p.fri_arity_bits);
        r.push(p.fri_arity_bits);
        db -= p.fri_arity_bits;
    }
    r
}

pub fn fri_step_path_len(l: &sb_proof_layout, round: u32) -> u32 {
    l.log_lde - l.arity_bits * (round + 1) - l.cap_len.trailing_zeros()
}

pub fn fri_step_offset(l: &sb_proof_layout, round: u32) -> u64 {
    let mut o = l.q_off_steps;
    for r in 0..round {
        o += (2u64 << l.arity_bits) + 4 * fri_step_path_len(l, r) as u64;
    }
    o
}

pub fn layout_for(p: &sb_params) -> sb_proof_layout {
    let ar = fri_arities(p);
    let qdf = if p.constraint_degree > 1 { p.constraint_degree - 1 } else { 1 };
    let mut l = sb_proof_layout::default();
    l.log_n = p.log_n;
    l.log_lde = p.log_n + p.rate_bits;
    l.n_cols = p.n_cols;
    l.n_quotient_polys = p.num_challenges * qdf;
    l.n_public_inputs = p.n_public_inputs;
    l.cap_len = 1 << p.cap_height;
    l.n_fri_rounds = ar.len() as u32;
    l.final_poly_len = 1 << (p.log_n - p.fri_arity_bits * ar.len() as u32);
    l.n_queries = p.num_query_rounds;
    l.arity_bits = p.fri_arity_bits;
    l.trace_path_len = l.log_lde - p.cap_height;
    let mut o = 0u64;
    l.off_trace_cap = o; o += 4 * l.cap_len as u64;
    l.off_quotient_cap = o; o += 4 * l.cap_len as u64;
    l.off_local_values = o; o += 2 * l.n_cols as u64;
    l.off_next_values = o; o += 2 * l.n_cols as u64;
    l.off_quotient_polys = o; o += 2 * l.n_quotient_polys as u64;
    l.off_fri_caps = o; o += 4 * l.cap_len as u64 * l.n_fri_rounds as u64;
    l.off_final_poly = o; o += 2 * l.final_poly_len as u64;
    l.off_pow_witness = o; o += 1;
    l.off_queries = o;
    let mut q = 0u64;
    l.q_off_trace_leaf = q; q += l.n_cols as u64;
    l.q_off_trace_path = q; q += 4 * l.trace_path_len as u64;
    l.q_off_quot_leaf = q; q += l.n_quotient_polys as u64;
    l.q_off_quot_path = q; q += 4 * l.trace_path_len as u64;
    l.q_off_steps = q;
    for r in 0..l.n_fri_rounds {
        q += (2u64 << l.arity_bits) + 4 * fri_step_path_len(&l, r) as u64;
    }
    l.query_stride = q;
    o += q * l.n_queries as u64;
    l.off_public_inputs = o; o += l.n_public_inputs as u64;
    l.total_words = o;
    l
}
