//! P3 of the parity ladder: the UNMODIFIED reference verifier accepts GPU proofs.
//! Not run in the build image (no cargo); run with `cargo test --release` inside the reference crate after adding
//! `starky_gpu` as a dependency and the five `impl GpuStark` lines of INTEGRATION.md.
//!
//!     let (stark, trace, pis) = /* exactly as aggregate_proof.rs:44-57 builds them */;
//!     let mut gpu = starky_gpu::GpuProver::new(0)?;
//!     let proof = starky_gpu::prove(&mut gpu, stark, &config, trace, &pis, &mut TimingTree::default())?;
//!     starky::verifier::verify_stark_proof(stark, proof.clone(), &config)?;      // aggregate_proof.rs:67
//!     recursive_proof::<F, C, S, C, D>(stark, proof, &config, true)?;             // aggregate_proof.rs:284
