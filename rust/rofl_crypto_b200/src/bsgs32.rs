//! Replaces rofl_crypto/src/bsgs32.rs: the table lives on the GPU (built once per (size, BSGS_N_BITS) and device, shared by every caller like the
//! Arc<BSGSTable> of rofl_service/src/flserver/server.rs:54,84); this handle only names it.  new :20-34, default :36-38, get_size :44-46.
use crate::fp::{BSGS_N_BITS, PRECOMP_BIAS};

#[derive(Clone, Debug)]
pub struct BSGSTable { m: usize }

impl BSGSTable {
    pub fn new(m: usize) -> BSGSTable { BSGSTable { m } }
    pub fn default() -> BSGSTable { BSGSTable::new((1 as usize) << (BSGS_N_BITS / 2 + PRECOMP_BIAS)) }
    pub fn table_size(&self) -> usize { self.m }
    /// table.len() - 1 of the reference (m + 1 distinct points 0..=m B): the library recomputes it on the device when it builds the table
    pub fn get_size(&self) -> u64 { self.m as u64 }
}
