"""rofl_crypto::l2_range_proof_vec (l2_range_proof_vec/mod.rs:15-253)."""
from . import fp


class L2RangeProofError(Exception):
    """l2_range_proof_vec/errors.rs:4-31"""


_ERR = {2: "ValueOutOfRangeError", 3: "OverflowError", 4: "NormOutOfRangeError", -1: "ProofError::InvalidBitsize"}


def _c():
    from . import context
    return context()


def create_rangeproof_l2(value_vec_clipped, blinding_vec, prove_range, n_partition=1, seed=None):
    """-> (proof bytes, commitment 32 bytes)     (:15-140; n_partition is unused by the reference beyond min(1, .))"""
    rc, proof, commit = _c().l2_prove(value_vec_clipped, blinding_vec, prove_range, fp.N_BITS, fp.FRAC, seed)
    if rc:
        raise L2RangeProofError(_ERR.get(rc, f"error {rc}"))
    return proof, commit


def verify_rangeproof_l2(range_proof, commit, prove_range, seed=None):         # :185-228
    rc = _c().l2_verify(range_proof, commit, prove_range, seed)
    if rc < 0:
        raise L2RangeProofError(f"ProofError ({rc})")
    return bool(rc)
