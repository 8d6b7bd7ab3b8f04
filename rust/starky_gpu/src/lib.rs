//! Drop-in for `starky::prover::prove` on the five starks of Electron-Labs/starky_bls12_381
//! (call sites: src/aggregate_proof.rs:59,105,138,169,212 and src/ecc_aggregate.rs:545 of the reference).
//!
//! Same signature, specialised to F = GoldilocksField, C = PoseidonGoldilocksConfig, D = 2.  The stark is identified to
//! the C side by `GpuStark::STARK_ID` (the library holds a pre-compiled constraint program per stark; it cannot call
//! back into `eval_packed_generic`).  SOURCE ONLY in this repository: no Rust toolchain in the build image.
//!
//! Cargo feature `gpu` (default): links libstarkyb200.so and provides `GpuProver` / `prove`.  Without it
//! (`--no-default-features`) the crate is pure Rust -- `layout`, `wire`, `unpack_words` -- so that
//! `cargo test --no-default-features` can feed a serialized GPU proof to the reference's verifier on a box without CUDA.
pub mod ffi;
pub mod layout;
pub mod wire;

use anyhow::{anyhow, Result};
use plonky2::field::extension::quadratic::QuadraticExtension;
use plonky2::field::goldilocks_field::GoldilocksField as F;
use plonky2::field::polynomial::{PolynomialCoeffs, PolynomialValues};
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::proof::{FriInitialTreeProof, FriProof, FriQueryRound, FriQueryStep};
use plonky2::hash::hash_types::HashOut;
use plonky2::hash::merkle_proofs::MerkleProof;
use plonky2::hash::merkle_tree::MerkleCap;
use plonky2::hash::poseidon::PoseidonHash;
use plonky2::plonk::config::PoseidonGoldilocksConfig as C;
use plonky2::util::timing::TimingTree;
use starky::config::StarkConfig;
use starky::proof::{StarkOpeningSet, StarkProof, StarkProofWithPublicInputs};
use starky::stark::Stark;

type FE = QuadraticExtension<F>;
const D: usize = 2;

/// Implemented (in the reference crate, one line each) for FP12MulStark = 0, PairingPrecompStark = 1,
/// MillerLoopStark = 2, FinalExponentiateStark = 3, ECCAggStark = 4 (enum sb_stark_id).
pub trait GpuStark: Stark<F, D> {
    const STARK_ID: u32;
    fn num_rows(&self) -> usize;
}

/// One GPU context; reuse it across the proofs of a signature verification (device buffers are grow-only).
#[cfg(feature = "gpu")]
pub struct GpuProver { ctx: *mut ffi::sb_ctx }
#[cfg(feature = "gpu")]
unsafe impl Send for GpuProver {}

#[cfg(feature = "gpu")]
impl GpuProver {
    pub fn new(device: i32) -> Result<Self> {
        Self::new_multi(&[device])
    }
    /// One context over several GPUs of the box: `prove` then shards every trace over them inside the library
    /// (column-sharded LDE storing into the owners' row buffers over NVLink, row-sharded leaf hashing and quotient;
    /// include/starky_b200.h "multi-GPU groups").  The number of devices must be a power of two.
    pub fn new_multi(devices: &[i32]) -> Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { ffi::sb_init(devices.as_ptr(), devices.len() as i32, &mut ctx) };
        if rc != ffi::SB_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::sb_last_error(std::ptr::null_mut())).to_string_lossy().into_owned() };
            return Err(anyhow!("sb_init failed ({rc}): {msg}"));
        }
        Ok(Self { ctx })
    }
    pub fn last_error(&self) -> String {
        unsafe { std::ffi::CStr::from_ptr(ffi::sb_last_error(self.ctx)).to_string_lossy().into_owned() }
    }
}
#[cfg(feature = "gpu")]
impl Drop for GpuProver { fn drop(&mut self) { unsafe { ffi::sb_destroy(self.ctx) } } }

#[cfg(feature = "gpu")]
fn params_for<S: GpuStark>(stark: &S, config: &StarkConfig, n_pis: usize) -> ffi::sb_params {
    let mut p = ffi::sb_params::default();
    let log_n = stark.num_rows().trailing_zeros();
    unsafe { ffi::sb_params_standard(S::STARK_ID, log_n, &mut p) };
    let f = &config.fri_config;
    p.n_cols = S::COLUMNS as u32;
    p.n_public_inputs = n_pis as u32;
    p.constraint_degree = stark.constraint_degree() as u32;
    p.rate_bits = f.rate_bits as u32;
    p.cap_height = f.cap_height as u32;
    p.num_challenges = config.num_challenges as u32;
    p.pow_bits = f.proof_of_work_bits;
    p.num_query_rounds = f.num_query_rounds as u32;
    p
}

/// `starky::prover::prove` on the GPU.  `_timing` is accepted for signature compatibility.
#[cfg(feature = "gpu")]
pub fn prove<S: GpuStark>(
    gpu: &mut GpuProver, stark: S, config: &StarkConfig, trace_poly_values: Vec<PolynomialValues<F>>,
    public_inputs: &[F], _timing: &mut TimingTree,
) -> Result<StarkProofWithPublicInputs<F, C, D>> {
    let p = params_for(&stark, config, public_inputs.len());
    // Vec<PolynomialValues<F>>: C separately allocated columns of canonical u64 (GoldilocksField is repr(transparent))
    let cols: Vec<*const u64> = trace_poly_values.iter().map(|v| v.values.as_ptr() as *const u64).collect();
    let pis: Vec<u64> = public_inputs.iter().map(|x| x.to_canonical_u64()).collect();
    let mut out = std::ptr::null_mut();
    let rc = unsafe {
        ffi::sb_prove(gpu.ctx, &p, cols.as_ptr() as *const _, ffi::SB_TRACE_COLS_U64_PTRS, pis.as_ptr(), &mut out)
    };
    match rc {
        ffi::SB_OK => {}
        // reference behaviour: panic!("Quotient has failed, the vanishing polynomial is not divisible by Z_H")
        ffi::SB_EQUOTIENT_NOT_DIVISIBLE => panic!("{}", gpu.last_error()),
        _ => return Err(anyhow!("{}", gpu.last_error())),
    }
    let proof = unsafe { unpack(&*out) };
    unsafe { ffi::sb_proof_free(out) };
    Ok(proof)
}

/// `prove` from the row-major `Vec<[F; COLUMNS]>` that `generate_trace` returns: skips
/// `trace_rows_to_poly_values` (aggregate_proof.rs:57,104,137,168,211), the transpose runs on the device.
#[cfg(feature = "gpu")]
pub fn prove_from_rows<S: GpuStark, const COLUMNS: usize>(
    gpu: &mut GpuProver, stark: S, config: &StarkConfig, rows: &[[F; COLUMNS]], public_inputs: &[F],
) -> Result<StarkProofWithPublicInputs<F, C, D>> {
    let p = params_for(&stark, config, public_inputs.len());
    let pis: Vec<u64> = public_inputs.iter().map(|x| x.to_canonical_u64()).collect();
    let mut out = std::ptr::null_mut();
    let rc = unsafe {
        ffi::sb_prove(gpu.ctx, &p, rows.as_ptr() as *const _, ffi::SB_TRACE_ROWMAJOR_U64, pis.as_ptr(), &mut out)
    };
    if rc != ffi::SB_OK { return Err(anyhow!("{}", gpu.last_error())); }
    let proof = unsafe { unpack(&*out) };
    unsafe { ffi::sb_proof_free(out) };
    Ok(proof)
}

/// The operands of one `*_main` function of aggregate_proof.rs, as the reference's limb arrays (`Fp = [u32; 12]`): the
/// library generates the trace in C++ (csrc/witness.cpp, the restated `generate_trace`) and proves it, so neither
/// `generate_trace` nor `trace_rows_to_poly_values` runs on the Rust side and the trace crosses PCIe once, as u32.
pub enum Operands<'a> {
    /// `fp12_mul_main(x, y)` (aggregate_proof.rs:122-151)
    Fp12Mul { x: &'a [[u32; 12]; 12], y: &'a [[u32; 12]; 12] },
    /// `calc_pairing_precomp_main(x, y, z)` (aggregate_proof.rs:24-69): Fp2 = [[u32; 12]; 2]
    PairingPrecomp { q: &'a [[[u32; 12]; 2]; 3] },
    /// `miller_loop_main(x, y, q)` (aggregate_proof.rs:71-121)
    MillerLoop { g1: &'a [[u32; 12]; 2], q: &'a [[[u32; 12]; 2]; 3] },
    /// `final_exponentiate_main(x)` (aggregate_proof.rs:153-184)
    FinalExp { x: &'a [[u32; 12]; 12] },
    /// `ec_aggregate_main(points, bits)` (aggregate_proof.rs:186-227): 512 affine G1 points and participation flags
    EccAgg { points: &'a [[[u32; 12]; 2]; 512], bits: &'a [u8; 512] },
}

/// `generate_trace` + `prove` inside the library.  `p` = `params_for(&stark, &config, S::PUBLIC_INPUTS)`.
#[cfg(feature = "gpu")]
pub fn prove_from_operands(gpu: &mut GpuProver, p: &ffi::sb_params, ops: Operands) -> Result<StarkProofWithPublicInputs<F, C, D>> {
    let mut out = std::ptr::null_mut();
    let rc = unsafe {
        match ops {
            Operands::Fp12Mul { x, y } => ffi::sb_prove_fp12_mul(gpu.ctx, p, x.as_ptr() as *const u32, y.as_ptr() as *const u32, &mut out),
            Operands::PairingPrecomp { q } => ffi::sb_prove_pairing_precomp(gpu.ctx, p, q.as_ptr() as *const u32, &mut out),
            Operands::MillerLoop { g1, q } => ffi::sb_prove_miller_loop(gpu.ctx, p, g1.as_ptr() as *const u32, q.as_ptr() as *const u32, &mut out),
            Operands::FinalExp { x } => ffi::sb_prove_final_exp(gpu.ctx, p, x.as_ptr() as *const u32, &mut out),
            Operands::EccAgg { points, bits } => ffi::sb_prove_ecc_agg(gpu.ctx, p, points.as_ptr() as *const u32, bits.as_ptr(), &mut out),
        }
    };
    if rc != ffi::SB_OK { return Err(anyhow!("{}", gpu.last_error())); }
    let proof = unsafe { unpack(&*out) };
    unsafe { ffi::sb_proof_free(out) };
    Ok(proof)
}

#[cfg(feature = "gpu")]
unsafe fn unpack(pr: &ffi::sb_proof) -> StarkProofWithPublicInputs<F, C, D> {
    unpack_words(&pr.layout, std::slice::from_raw_parts(pr.words, pr.layout.total_words as usize))
}

/// The flat POD of include/starky_b200.h (layout `l`, words `w`) as the struct the reference's verifier and recursion
/// circuit take.  Pure Rust: also used on proofs read from a file (wire.rs).
pub fn unpack_words(l: &ffi::sb_proof_layout, w: &[u64]) -> StarkProofWithPublicInputs<F, C, D> {
    assert_eq!(w.len() as u64, l.total_words);
    let f = |i: usize| F::from_canonical_u64(w[i]);
    let fe = |i: usize| FE::from([f(i), f(i + 1)]);
    let cap = |off: usize| MerkleCap::<F, PoseidonHash>(
        (0..l.cap_len as usize).map(|i| HashOut { elements: [f(off + 4 * i), f(off + 4 * i + 1), f(off + 4 * i + 2), f(off + 4 * i + 3)] }).collect());
    let path = |off: usize, len: usize| MerkleProof::<F, PoseidonHash> {
        siblings: (0..len).map(|i| HashOut { elements: [f(off + 4 * i), f(off + 4 * i + 1), f(off + 4 * i + 2), f(off + 4 * i + 3)] }).collect() };
    let (c, nq) = (l.n_cols as usize, l.n_quotient_polys as usize);
    let openings = StarkOpeningSet {
        local_values: (0..c).map(|i| fe(l.off_local_values as usize + 2 * i)).collect(),
        next_values: (0..c).map(|i| fe(l.off_next_values as usize + 2 * i)).collect(),
        permutation_zs: None,
        permutation_zs_next: None,
        quotient_polys: (0..nq).map(|i| fe(l.off_quotient_polys as usize + 2 * i)).collect(),
    };
    let arity = 1usize << l.arity_bits;
    let query_round_proofs = (0..l.n_queries as usize).map(|q| {
        let b = (l.off_queries + q as u64 * l.query_stride) as usize;
        let tl = l.trace_path_len as usize;
        let evals_proofs = vec![
            ((0..c).map(|i| f(b + l.q_off_trace_leaf as usize + i)).collect(), path(b + l.q_off_trace_path as usize, tl)),
            ((0..nq).map(|i| f(b + l.q_off_quot_leaf as usize + i)).collect(), path(b + l.q_off_quot_path as usize, tl)),
        ];
        let steps = (0..l.n_fri_rounds).map(|r| {
            let o = b + layout::fri_step_offset(l, r) as usize;
            FriQueryStep { evals: (0..arity).map(|i| fe(o + 2 * i)).collect(),
                           merkle_proof: path(o + 2 * arity, layout::fri_step_path_len(l, r) as usize) }
        }).collect();
        FriQueryRound { initial_trees_proof: FriInitialTreeProof { evals_proofs }, steps }
    }).collect();
    let opening_proof = FriProof {
        commit_phase_merkle_caps: (0..l.n_fri_rounds as usize).map(|r| cap(l.off_fri_caps as usize + 4 * l.cap_len as usize * r)).collect(),
        query_round_proofs,
        final_poly: PolynomialCoeffs::new((0..l.final_poly_len as usize).map(|i| fe(l.off_final_poly as usize + 2 * i)).collect()),
        pow_witness: f(l.off_pow_witness as usize),
    };
    StarkProofWithPublicInputs {
        proof: StarkProof { trace_cap: cap(l.off_trace_cap as usize), permutation_zs_cap: None,
                            quotient_polys_cap: cap(l.off_quotient_cap as usize), openings, opening_proof },
        public_inputs: (0..l.n_public_inputs as usize).map(|i| f(l.off_public_inputs as usize + i)).collect(),
    }
}
