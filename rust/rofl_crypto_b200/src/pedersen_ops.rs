//! Replaces the vector functions of rofl_crypto/src/pedersen_ops.rs that sit on the hot path (commit_no_blinding_vec :9-16, commit_vec :18-25,
//! default_discrete_log_vec / discrete_log_vec / discrete_log_vec_table :27-53, add_rp_vec_vec :61-69); the small helpers (zero vectors, shifts,
//! cancelling / random scalars :71-127) stay as they are in the reference file.
use curve25519_dalek_ng::ristretto::RistrettoPoint;
use curve25519_dalek_ng::scalar::Scalar;

use crate::b200::{self, ffi};
use crate::bsgs32::BSGSTable;

/// the library commits f32 values (conversion fused into the kernel); scalars that are already converted go through their f32 form only when they
/// came from f32_to_scalar_vec -- these two functions take SCALARS like the reference, so they use the scalar entry of the ABI: v*B + r*H is
/// computed as commit(0.0, r) + v*B would cost a second pass, hence the dedicated call below.
pub fn commit_no_blinding_vec(scalar_vec: &Vec<Scalar>) -> Vec<RistrettoPoint> { commit_scalars(scalar_vec, None) }
pub fn commit_vec(scalar_vec: &Vec<Scalar>, blinding_vec: &Vec<Scalar>) -> Vec<RistrettoPoint> { commit_scalars(scalar_vec, Some(blinding_vec)) }
fn commit_scalars(values: &Vec<Scalar>, blind: Option<&Vec<Scalar>>) -> Vec<RistrettoPoint> {
    // values are fixed-point scalars +-raw (conversion32::f32_to_scalar): hand them over as f32 = scalar_to_f32 (exact for |raw| < 2^24, the range
    // every experiment of the reference uses); larger magnitudes take the reference's own CPU path
    let f = crate::conversion32::scalar_to_f32_vec(values);
    if crate::conversion32::f32_to_scalar_vec(&f) != *values {
        let pc = bulletproofs::PedersenGens::default();
        return values.iter().enumerate().map(|(i, v)| pc.commit(*v, blind.map_or(Scalar::zero(), |b| b[i]))).collect();
    }
    let d = values.len();
    let mut out = vec![0u8; 32 * d];
    let bl = blind.map(|b| b200::scs(b));
    let rc = unsafe { ffi::rofl_commit(b200::ctx(), f.as_ptr(), bl.as_ref().map_or(std::ptr::null(), |b| b.as_ptr()), d, b200::n_bits(), b200::frac(), out.as_mut_ptr(), std::ptr::null_mut()) };
    assert!(rc == 0, "commit: rofl_b200 error {}: {}", rc, b200::last_error());
    b200::unpts(&out)
}

pub fn default_discrete_log_vec(rp_vec: &Vec<RistrettoPoint>) -> Vec<Scalar> { discrete_log_vec_table(rp_vec, &BSGSTable::default()) }
pub fn discrete_log_vec(rp_vec: &Vec<RistrettoPoint>, table_size: usize) -> Vec<Scalar> { discrete_log_vec_table(rp_vec, &BSGSTable::new(table_size)) }
pub fn discrete_log_vec_table(rp_vec: &Vec<RistrettoPoint>, bsgs: &BSGSTable) -> Vec<Scalar> {
    let d = rp_vec.len();
    let p = b200::pts(rp_vec);
    let mut out = vec![0u8; 32 * d];
    let rc = unsafe { ffi::rofl_dlog(b200::ctx(), p.as_ptr(), d, bsgs.table_size() as u64, b200::bsgs_bits(), b200::n_bits(), b200::frac(), out.as_mut_ptr(), std::ptr::null_mut()) };
    assert!(rc == 0, "discrete log: no logarithm in range (the reference unwrap()s here too, bsgs32.rs:69-70): {}", rc);
    b200::unscs(&out)
}

pub fn add_rp_vec_vec(rp_vec_vec: &Vec<Vec<RistrettoPoint>>) -> Vec<RistrettoPoint> {
    let (k, d) = (rp_vec_vec.len(), rp_vec_vec[0].len());
    let mut all = Vec::with_capacity(32 * k * d);
    for v in rp_vec_vec { assert_eq!(v.len(), d); all.extend_from_slice(&b200::pts(v)); }
    let mut out = vec![0u8; 32 * d];
    let rc = unsafe { ffi::rofl_aggregate(b200::ctx(), all.as_ptr(), k, d, 0, out.as_mut_ptr()) };
    assert!(rc == 0, "aggregate: rofl_b200 error {}", rc);
    b200::unpts(&out)
}
