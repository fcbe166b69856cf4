//! Replaces rofl_crypto/src/square_proof_vec/mod.rs (keep errors.rs): create_l2rangeproof_vec_existing :19-75, create_l2rangeproof_vec :77-128,
//! verify_l2rangeproof_vec :130-160.  SquareProof / SquareProofCommitments stay the reference's types (square_proof/mod.rs, pedersen.rs); the
//! library speaks their to_bytes() forms (160 B, 64 B).
use curve25519_dalek_ng::ristretto::RistrettoPoint;
use curve25519_dalek_ng::scalar::Scalar;

pub mod errors;
pub use self::errors::L2RangeProofError;
use crate::b200::{self, ffi};
use crate::square_proof::pedersen::SquareProofCommitments;
use crate::square_proof::{ProofError, SquareProof};

fn prove(value_vec: &Vec<f32>, value_com: &[u8], r1: &Vec<Scalar>, r2: &Vec<Scalar>) -> Result<(Vec<SquareProof>, Vec<SquareProofCommitments>), L2RangeProofError> {
    let d = value_vec.len();
    let (mut proofs, mut commits) = (vec![0u8; 160 * d], vec![0u8; 64 * d]);
    let (b1, b2) = (b200::scs(r1), b200::scs(r2));
    let seed = b200::seed();
    let rc = unsafe {
        ffi::rofl_square_prove(b200::ctx(), value_vec.as_ptr(), value_com.as_ptr(), b1.as_ptr(), b2.as_ptr(), d, b200::n_bits(), b200::frac(), seed.as_ptr(),
                               proofs.as_mut_ptr(), commits.as_mut_ptr())
    };
    if rc != 0 { panic!("square proofs: rofl_b200 error {}: {}", rc, b200::last_error()); }
    Ok((proofs.chunks_exact(160).map(|p| SquareProof::from_bytes(p).expect("malformed proof")).collect(),
        commits.chunks_exact(64).map(|c| SquareProofCommitments::from_bytes(c).expect("malformed commitments")).collect()))
}

pub fn create_l2rangeproof_vec_existing(value_vec: &Vec<f32>, value_com_vec: Vec<RistrettoPoint>, random_vec: &Vec<Scalar>, random_vec_2: &Vec<Scalar>)
    -> Result<(Vec<SquareProof>, Vec<SquareProofCommitments>), L2RangeProofError> {
    if value_vec.len() != random_vec.len() { return Err(L2RangeProofError::WrongNumBlindingFactors); }
    prove(value_vec, &b200::pts(&value_com_vec), random_vec, random_vec_2)
}

pub fn create_l2rangeproof_vec(value_vec: &Vec<f32>, random_vec: &Vec<Scalar>, random_vec_2: &Vec<Scalar>)
    -> Result<(Vec<SquareProof>, Vec<SquareProofCommitments>), L2RangeProofError> {
    if value_vec.len() != random_vec.len() { return Err(L2RangeProofError::WrongNumBlindingFactors); }
    let d = value_vec.len();
    let mut c = vec![0u8; 32 * d];
    let b1 = b200::scs(random_vec);
    let rc = unsafe { ffi::rofl_commit(b200::ctx(), value_vec.as_ptr(), b1.as_ptr(), d, b200::n_bits(), b200::frac(), c.as_mut_ptr(), std::ptr::null_mut()) };
    if rc != 0 { panic!("commit: rofl_b200 error {}: {}", rc, b200::last_error()); }
    prove(value_vec, &c, random_vec, random_vec_2)
}

pub fn verify_l2rangeproof_vec(randproof_vec: &Vec<SquareProof>, commit_vec: &Vec<SquareProofCommitments>) -> Result<bool, L2RangeProofError> {
    if randproof_vec.len() != commit_vec.len() { return Err(L2RangeProofError::WrongNumberOfElGamalPairs); }
    let (mut p, mut c) = (Vec::with_capacity(160 * randproof_vec.len()), Vec::with_capacity(64 * commit_vec.len()));
    for x in randproof_vec { p.extend_from_slice(&x.to_bytes()); }
    for x in commit_vec { c.extend_from_slice(&x.to_bytes()); }
    match unsafe { ffi::rofl_square_verify(b200::ctx(), p.as_ptr(), c.as_ptr(), randproof_vec.len()) } {
        1 => Ok(true),
        0 => Ok(false),
        _ => Err(ProofError::FormatError.into()),
    }
}
