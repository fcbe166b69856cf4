//! Replaces the three `helper_*` functions at the end of `impl CompressedRandProof` in rofl_crypto/src/compressed_rand_proof/mod.rs:134-158 (the
//! struct, to_bytes / from_bytes and the single-proof prove / verify stay).  These are the calls rofl_service makes (params.rs:729-735,820-823):
//! they also materialise the ElGamal right halves R_i = r_i B of every optimised encoding.
use curve25519_dalek_ng::ristretto::RistrettoPoint;
use curve25519_dalek_ng::scalar::Scalar;

use super::types::CompressedRandProofCommitments;
use super::{CompressedRandProof, ProofError};
use crate::b200::{self, ffi};
use crate::rand_proof::ElGamalPair;

impl CompressedRandProof {
    pub fn helper_prove(m_vec: &Vec<f32>, r_vec: Vec<Scalar>) -> Result<(CompressedRandProof, CompressedRandProofCommitments), ProofError> { Self::b200_prove(m_vec, None, r_vec) }
    pub fn helper_prove_existing(m_vec: &Vec<f32>, m_com: Vec<RistrettoPoint>, r_vec: Vec<Scalar>) -> Result<(CompressedRandProof, CompressedRandProofCommitments), ProofError> {
        Self::b200_prove(m_vec, Some(b200::pts(&m_com)), r_vec)
    }
    fn b200_prove(m_vec: &Vec<f32>, m_com: Option<Vec<u8>>, r_vec: Vec<Scalar>) -> Result<(CompressedRandProof, CompressedRandProofCommitments), ProofError> {
        let d = m_vec.len();
        let (mut proof, mut pairs) = ([0u8; 128], vec![0u8; 64 * d.max(1)]);
        let bl = b200::scs(&r_vec);
        let seed = b200::seed();
        let rc = unsafe {
            ffi::rofl_crp_prove(b200::ctx(), m_vec.as_ptr(), m_com.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()), bl.as_ptr(), d, b200::n_bits(), b200::frac(), seed.as_ptr(),
                                proof.as_mut_ptr(), pairs.as_mut_ptr())
        };
        if rc != 0 { panic!("compressed rand proof: rofl_b200 error {} (more than 900 000 pairs exhaust the reference's label table too): {}", rc, b200::last_error()); }
        let c_vec = pairs[..64 * d].chunks_exact(64).map(|c| ElGamalPair::from_bytes(c).expect("malformed pair")).collect();
        Ok((CompressedRandProof::from_bytes(&proof)?, CompressedRandProofCommitments { c_vec }))
    }
    pub fn helper_verify(&self, c_vec: Vec<ElGamalPair>) -> Result<(), ProofError> {
        let mut pairs = Vec::with_capacity(64 * c_vec.len());
        for c in &c_vec { pairs.extend_from_slice(&c.to_bytes()); }
        let proof = self.to_bytes();
        match unsafe { ffi::rofl_crp_verify(b200::ctx(), proof.as_ptr(), pairs.as_ptr(), c_vec.len()) } {
            1 => Ok(()),
            0 => Err(ProofError::VerificationError),
            _ => Err(ProofError::FormatError),
        }
    }
}
