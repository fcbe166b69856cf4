//! Glue between the reference's Rust types and the C ABI of librofl_b200.so: one lazily created context per process (thread-safe: concurrent
//! callers run on separate lanes inside the library), byte conversions at the edge, a fresh OS-random seed per call (the reference draws every
//! nonce from `rand::thread_rng()`; the library derives them from this seed).
//! Add to rofl_crypto/src/lib.rs:  `pub mod b200;`
use curve25519_dalek_ng::ristretto::{CompressedRistretto, RistrettoPoint};
use curve25519_dalek_ng::scalar::Scalar;
use once_cell::sync::Lazy;
use rand::RngCore;
pub use rofl_b200_sys as ffi;

use crate::fp::{BSGS_N_BITS, N_BITS};

pub struct Ctx(pub *mut ffi::rofl_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

pub static CTX: Lazy<Ctx> = Lazy::new(|| {
    let mut p = std::ptr::null_mut();
    let dev = std::env::var("ROFL_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
    let rc = unsafe { ffi::rofl_ctx_create(&mut p, dev) };
    assert_eq!(rc, 0, "rofl_b200 needs a CUDA device (there is no CPU fallback): {}", last_error());
    Ctx(p)
});

pub fn ctx() -> *mut ffi::rofl_ctx { CTX.0 }
pub fn last_error() -> String { unsafe { std::ffi::CStr::from_ptr(ffi::rofl_last_error()) }.to_string_lossy().into_owned() }
/// cargo features fpN / fracK of the reference (fp.rs:35-137) -> the runtime pair of the library
pub fn n_bits() -> i32 { N_BITS as i32 }
pub fn frac() -> i32 { crate::fp::Fix::frac_nbits() as i32 }
pub fn bsgs_bits() -> i32 { BSGS_N_BITS as i32 }
pub fn seed() -> [u8; 32] { let mut s = [0u8; 32]; rand::thread_rng().fill_bytes(&mut s); s }

pub fn pts(v: &[RistrettoPoint]) -> Vec<u8> { let mut o = Vec::with_capacity(32 * v.len()); for p in v { o.extend_from_slice(p.compress().as_bytes()); } o }
pub fn scs(v: &[Scalar]) -> Vec<u8> { let mut o = Vec::with_capacity(32 * v.len()); for s in v { o.extend_from_slice(s.as_bytes()); } o }
/// the library only returns encodings it produced itself: they always decode
pub fn unpts(b: &[u8]) -> Vec<RistrettoPoint> { b.chunks_exact(32).map(|c| CompressedRistretto::from_slice(c).decompress().expect("library returned an invalid point")).collect() }
pub fn unscs(b: &[u8]) -> Vec<Scalar> { b.chunks_exact(32).map(|c| { let mut a = [0u8; 32]; a.copy_from_slice(c); Scalar::from_bytes_mod_order(a) }).collect() }
