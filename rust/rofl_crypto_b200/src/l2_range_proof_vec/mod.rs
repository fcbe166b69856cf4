//! Replaces rofl_crypto/src/l2_range_proof_vec/mod.rs (keep errors.rs): create_rangeproof_l2 :15-140, verify_rangeproof_l2 :185-228.
use bulletproofs::{ProofError, RangeProof};
use curve25519_dalek_ng::ristretto::{CompressedRistretto, RistrettoPoint};
use curve25519_dalek_ng::scalar::Scalar;

pub mod errors;
use self::errors::L2RangeProofError;
use crate::b200::{self, ffi};

pub fn create_rangeproof_l2(value_vec_clipped: &Vec<f32>, blinding_vec: &Vec<Scalar>, prove_range: usize, _n_partition: usize)
    -> Result<(RangeProof, RistrettoPoint), L2RangeProofError> {
    if value_vec_clipped.len() != blinding_vec.len() {
        return Err(ProofError::WrongNumBlindingFactors.into());
    }
    let mut proof = vec![0u8; unsafe { ffi::rofl_range_proof_len(prove_range.max(1)) }];
    let (mut pl, mut commit) = (0usize, [0u8; 32]);
    let blind = b200::scs(blinding_vec);
    let seed = b200::seed();
    let rc = unsafe {
        ffi::rofl_l2_prove(b200::ctx(), value_vec_clipped.as_ptr(), blind.as_ptr(), value_vec_clipped.len(), prove_range as i32, b200::n_bits(), b200::frac(), seed.as_ptr(),
                           proof.as_mut_ptr(), &mut pl, commit.as_mut_ptr())
    };
    match rc {
        0 => Ok((RangeProof::from_bytes(&proof[..pl]).expect("library returned a malformed proof"), CompressedRistretto(commit).decompress().expect("invalid point"))),
        ffi::ROFL_ERR_VALUE_OUT_OF_RANGE => Err(L2RangeProofError::ValueOutOfRangeError),
        ffi::ROFL_ERR_OVERFLOW => Err(L2RangeProofError::OverflowError("scalar sum".into(), "f32 sum".into())),
        ffi::ROFL_ERR_NORM_OUT_OF_RANGE => Err(L2RangeProofError::NormOutOfRangeError("sum of squares".into())),
        ffi::ROFL_ERR_BITSIZE => Err(ProofError::InvalidBitsize.into()),
        _ => panic!("create_rangeproof_l2: rofl_b200 error {}: {}", rc, b200::last_error()),
    }
}

pub fn verify_rangeproof_l2(range_proof: &RangeProof, commit: &RistrettoPoint, prove_range: usize) -> Result<bool, ProofError> {
    let p = range_proof.to_bytes();
    let seed = b200::seed();
    let rc = unsafe { ffi::rofl_l2_verify(b200::ctx(), p.as_ptr(), p.len(), commit.compress().as_bytes().as_ptr(), prove_range as i32, seed.as_ptr()) };
    match rc {
        1 => Ok(true),
        0 => Ok(false),
        ffi::ROFL_ERR_BITSIZE => Err(ProofError::InvalidBitsize),
        ffi::ROFL_ERR_GENS => Err(ProofError::InvalidGeneratorsLength),
        _ => Err(ProofError::FormatError),
    }
}
