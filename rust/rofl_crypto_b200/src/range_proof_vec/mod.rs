//! Replaces rofl_crypto/src/range_proof_vec/mod.rs (keep errors.rs next to it).  Same public functions and signatures
//! (create_rangeproof :16-102, clip_f32_to_range_vec :104-111, verify_rangeproof :149-191); the work runs on the GPU.
use bulletproofs::{ProofError, RangeProof};
use curve25519_dalek_ng::ristretto::RistrettoPoint;
use curve25519_dalek_ng::scalar::Scalar;

pub mod errors;
use self::errors::RangeProofError;
use crate::b200::{self, ffi};

pub fn create_rangeproof(value_vec_clipped: &Vec<f32>, blinding_vec: &Vec<Scalar>, prove_range: usize, n_partition: usize)
    -> Result<(Vec<RangeProof>, Vec<RistrettoPoint>), RangeProofError> {
    if value_vec_clipped.len() != blinding_vec.len() {
        return Err(ProofError::WrongNumBlindingFactors.into());
    }
    let d = value_vec_clipped.len();
    let (mut np, mut pl) = (0usize, 0usize);
    unsafe { ffi::rofl_range_proof_shape(d, prove_range as i32, n_partition, &mut np, &mut pl) };
    let (mut proofs, mut commits) = (vec![0u8; np * pl], vec![0u8; 32 * d]);
    let blind = b200::scs(blinding_vec);
    let seed = b200::seed();
    let rc = unsafe {
        ffi::rofl_range_prove(b200::ctx(), value_vec_clipped.as_ptr(), blind.as_ptr(), d, prove_range as i32, n_partition, b200::n_bits(), b200::frac(),
                              seed.as_ptr(), proofs.as_mut_ptr(), &mut pl, &mut np, commits.as_mut_ptr())
    };
    match rc {
        0 => Ok((proofs.chunks_exact(pl).map(|p| RangeProof::from_bytes(p).expect("library returned a malformed proof")).collect(), b200::unpts(&commits))),
        ffi::ROFL_ERR_VALUE_OUT_OF_RANGE => Err(RangeProofError::ValueOutOfRangeError),
        ffi::ROFL_ERR_BITSIZE => Err(ProofError::InvalidBitsize.into()),
        // the reference panics on the same inputs: NaN (Fix::saturating_from_float), non power-of-two chunking ("Should not get here", :137-140)
        _ => panic!("create_rangeproof: rofl_b200 error {}: {}", rc, b200::last_error()),
    }
}

pub fn clip_f32_to_range_vec(value_vec: &Vec<f32>, prove_range: usize) -> Vec<f32> {
    let mut out = vec![0f32; value_vec.len()];
    unsafe { ffi::rofl_clip_f32_to_range_vec(value_vec.as_ptr(), value_vec.len(), prove_range as i32, b200::n_bits(), b200::frac(), out.as_mut_ptr()) };
    out
}

pub fn verify_rangeproof(range_proof_vec: &Vec<RangeProof>, commit_vec: &Vec<RistrettoPoint>, prove_range: usize) -> Result<bool, ProofError> {
    let mut bytes = Vec::new();
    for p in range_proof_vec { bytes.extend_from_slice(&p.to_bytes()); }
    let pl = if range_proof_vec.is_empty() { 0 } else { bytes.len() / range_proof_vec.len() };
    let commits = b200::pts(commit_vec);
    let seed = b200::seed();
    let rc = unsafe { ffi::rofl_range_verify(b200::ctx(), bytes.as_ptr(), pl, range_proof_vec.len(), commits.as_ptr(), commit_vec.len(), prove_range as i32, seed.as_ptr()) };
    match rc {
        1 => Ok(true),
        0 => Ok(false),
        ffi::ROFL_ERR_BITSIZE => Err(ProofError::InvalidBitsize),
        ffi::ROFL_ERR_GENS => Err(ProofError::InvalidGeneratorsLength),
        ffi::ROFL_ERR_FORMAT => Err(ProofError::FormatError),
        _ => panic!("verify_rangeproof: rofl_b200 error {}: {}", rc, b200::last_error()),
    }
}

/// Server side, all clients of a round in one batched check (one random linear combination over clients x chunks; the reference calls
/// verify_rangeproof once per client from a rayon pool, rofl_service/src/flserver/server.rs:516-522,666-667).  Every update must have the same
/// length, range and number of proofs.  One verdict per client, identical to the per-client calls'.
pub fn verify_rangeproof_batch(proofs: &[&Vec<RangeProof>], commits: &[&Vec<RistrettoPoint>], prove_range: usize) -> Vec<Result<bool, ProofError>> {
    let k = proofs.len();
    assert!(k > 0 && commits.len() == k);
    let (np, d) = (proofs[0].len(), commits[0].len());
    let mut pb = Vec::new(); let mut cb = Vec::with_capacity(32 * d * k);
    for (p, c) in proofs.iter().zip(commits) { assert!(p.len() == np && c.len() == d); for x in p.iter() { pb.extend_from_slice(&x.to_bytes()); } cb.extend_from_slice(&b200::pts(c)); }
    let pl = pb.len() / (k * np);
    let mut ok = vec![0i32; k];
    let seed = b200::seed();
    let rc = unsafe { ffi::rofl_range_verify_batch(b200::ctx(), pb.as_ptr(), pl, np, cb.as_ptr(), d, k, prove_range as i32, seed.as_ptr(), ok.as_mut_ptr()) };
    assert!(rc == 0, "verify_rangeproof_batch: rofl_b200 error {}: {}", rc, b200::last_error());
    ok.into_iter().map(|r| match r { 1 => Ok(true), 0 => Ok(false), ffi::ROFL_ERR_BITSIZE => Err(ProofError::InvalidBitsize), ffi::ROFL_ERR_GENS => Err(ProofError::InvalidGeneratorsLength), _ => Err(ProofError::FormatError) }).collect()
}
