//! Reading a serialized proof (include/starky_b200.h: enum sb_wire_format) in pure Rust.
//!
//! `SB_WIRE_POD` image: "SBPROOF1" | sb_params (64 bytes, little-endian u32 fields + one u64) | total_words u64 |
//! words (little-endian u64) -- what `sb_proof_serialize(.., SB_WIRE_POD, ..)` writes and what
//! tests/golden/ecc_agg_proof.sbproof holds.
use anyhow::{anyhow, ensure, Result};
use plonky2::plonk::config::PoseidonGoldilocksConfig as C;
use starky::config::StarkConfig;
use starky::proof::StarkProofWithPublicInputs;

use crate::ffi::sb_params;
use crate::layout::layout_for;

type F = plonky2::field::goldilocks_field::GoldilocksField;
const GOLDILOCKS_ORDER: u64 = 0xFFFF_FFFF_0000_0001;

fn u32_at(b: &[u8], o: usize) -> u32 { u32::from_le_bytes(b[o..o + 4].try_into().unwrap()) }
fn u64_at(b: &[u8], o: usize) -> u64 { u64::from_le_bytes(b[o..o + 8].try_into().unwrap()) }

/// (parameters of the image, the proof as the struct `verify_stark_proof` takes)
pub fn proof_from_pod_image(bytes: &[u8]) -> Result<(sb_params, StarkProofWithPublicInputs<F, C, 2>)> {
    ensure!(bytes.len() >= 8 + 64 + 8 && &bytes[..8] == b"SBPROOF1", "not an SBPROOF1 image");
    let h = &bytes[8..72];
    let p = sb_params {
        stark_id: u32_at(h, 0), log_n: u32_at(h, 4), n_cols: u32_at(h, 8), n_public_inputs: u32_at(h, 12),
        constraint_degree: u32_at(h, 16), rate_bits: u32_at(h, 20), cap_height: u32_at(h, 24), num_challenges: u32_at(h, 28),
        pow_bits: u32_at(h, 32), num_query_rounds: u32_at(h, 36), fri_arity_bits: u32_at(h, 40),
        fri_final_poly_bits: u32_at(h, 44), flags: u32_at(h, 48), reserved: u32_at(h, 52), fixed_pow_witness: u64_at(h, 56),
    };
    ensure!(p.fri_arity_bits >= 1 && p.log_n >= 1 && p.log_n <= 13 && p.cap_height <= p.log_n + p.rate_bits, "bad parameters in the image");
    let l = layout_for(&p);
    let n = u64_at(bytes, 72);
    ensure!(n == l.total_words, "image holds {n} words, the layout has {}", l.total_words);
    ensure!(bytes.len() as u64 == 80 + 8 * n, "image is {} bytes, expected {}", bytes.len(), 80 + 8 * n);
    let words: Vec<u64> = (0..n as usize).map(|i| u64_at(bytes, 80 + 8 * i)).collect();
    if let Some(i) = words.iter().position(|&w| w >= GOLDILOCKS_ORDER) {
        return Err(anyhow!("non-canonical field element at word {i}"));
    }
    Ok((p, crate::unpack_words(&l, &words)))
}

/// The StarkConfig the image was proved with (standard_fast_config() + the per-stark overrides of aggregate_proof.rs).
pub fn config_of(p: &sb_params) -> StarkConfig {
    let mut config = StarkConfig::standard_fast_config();
    config.num_challenges = p.num_challenges as usize;
    config.fri_config.rate_bits = p.rate_bits as usize;
    config.fri_config.cap_height = p.cap_height as usize;
    config.fri_config.proof_of_work_bits = p.pow_bits;
    config.fri_config.num_query_rounds = p.num_query_rounds as usize;
    config.fri_config.reduction_strategy =
        plonky2::fri::reduction_strategies::FriReductionStrategy::ConstantArityBits(p.fri_arity_bits as usize, p.fri_final_poly_bits as usize);
    config
}
