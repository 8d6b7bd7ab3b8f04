//! P3 of the parity ladder: the UNMODIFIED reference verifier accepts a GPU proof.
//!
//! The image tests/golden/ecc_agg_proof.sbproof is a VALID ECCAggStark proof (3339 columns x 8192 rows,
//! StarkConfig::standard_fast_config() with rate_bits = 2, the configuration of aggregate_proof.rs:186-188); the
//! repository's GPU test tests/test_gpu_wire.py asserts that `sb_prove` + `sb_proof_serialize` reproduce that file byte for
//! byte.  This test reads it in pure Rust and hands it to the reference's own acceptance path:
//!   * `starky::verifier::verify_stark_proof`            (aggregate_proof.rs:220, ecc_aggregate.rs:553)
//!   * the recursive verifier circuit, restating the private `recursive_proof` (aggregate_proof.rs:417-451) line by line
//!
//! No CUDA needed:   cargo test --release --no-default-features --test verify
//! (Proving a fresh trace on a GPU box is the reference's own `ecc_aggregate.rs::tests::test_stark` after the
//! INTEGRATION.md patch: `impl GpuStark` must live in the reference crate, so that test does too.)
use plonky2::field::extension::quadratic::QuadraticExtension;
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::types::Field;
use plonky2::iop::witness::PartialWitness;
use plonky2::plonk::circuit_builder::CircuitBuilder;
use plonky2::plonk::circuit_data::CircuitConfig;
use plonky2::plonk::config::PoseidonGoldilocksConfig;
use starky::recursive_verifier::{add_virtual_stark_proof_with_pis, set_stark_proof_with_pis_target, verify_stark_proof_circuit};
use starky::verifier::verify_stark_proof;
use starky_bls12_381::ecc_aggregate::ECCAggStark;

type F = GoldilocksField;
type C = PoseidonGoldilocksConfig;
const D: usize = 2;

fn golden() -> Vec<u8> {
    std::fs::read(concat!(env!("CARGO_MANIFEST_DIR"), "/../../tests/golden/ecc_agg_proof.sbproof")).expect("golden image")
}

#[test]
fn reference_verifier_accepts_the_serialized_gpu_proof() {
    let (p, proof) = starky_gpu::wire::proof_from_pod_image(&golden()).unwrap();
    assert_eq!((p.stark_id, p.n_cols, p.log_n, p.rate_bits), (4, 3339, 13, 2));
    let config = starky_gpu::wire::config_of(&p);
    let stark = ECCAggStark::<F, D>::new(1 << p.log_n);
    verify_stark_proof(stark, proof.clone(), &config).unwrap();

    // a flipped opening must be rejected
    let mut bad = proof.clone();
    bad.proof.openings.local_values[7] += QuadraticExtension::<F>::ONE;
    assert!(verify_stark_proof(stark, bad, &config).is_err());
}

#[test]
fn recursive_verifier_circuit_accepts_the_serialized_gpu_proof() {
    let (p, proof) = starky_gpu::wire::proof_from_pod_image(&golden()).unwrap();
    let config = starky_gpu::wire::config_of(&p);
    let stark = ECCAggStark::<F, D>::new(1 << p.log_n);
    // aggregate_proof.rs:417-451 (`recursive_proof` is private there)
    let mut builder = CircuitBuilder::<F, D>::new(CircuitConfig::standard_recursion_config());
    let mut pw = PartialWitness::new();
    let degree_bits = proof.proof.recover_degree_bits(&config);
    assert_eq!(degree_bits, p.log_n as usize);
    let pt = add_virtual_stark_proof_with_pis(&mut builder, stark, &config, degree_bits);
    builder.register_public_inputs(&pt.public_inputs);
    set_stark_proof_with_pis_target(&mut pw, &pt, &proof);
    verify_stark_proof_circuit::<F, C, ECCAggStark<F, D>, D>(&mut builder, stark, pt, &config);
    let data = builder.build::<C>();
    let rec = data.prove(pw).unwrap();
    data.verify(rec).unwrap();
}
