"""rofl_crypto::range_proof_vec (range_proof_vec/mod.rs:16-246)."""
from . import fp


class RangeProofError(Exception):
    """range_proof_vec/errors.rs:5-19"""


def _c():
    from . import context
    return context()


def clip_f32_to_range_vec(value_vec, prove_range):                     # :104-111
    return _c().clip_f32_to_range_vec(value_vec, prove_range, fp.N_BITS, fp.FRAC)


def create_rangeproof(value_vec_clipped, blinding_vec, prove_range, n_partition, seed=None):
    """-> (proofs [n_chunks, proof_len] uint8, commitments [D, 32] uint8)     (:16-102)"""
    rc, proofs, commits = _c().range_prove(value_vec_clipped, blinding_vec, prove_range, n_partition, fp.N_BITS, fp.FRAC, seed)
    if rc == 2:
        raise RangeProofError("ValueOutOfRangeError")
    if rc == -1:
        raise RangeProofError("ProofError::InvalidBitsize")
    if rc:
        raise RangeProofError(f"create_rangeproof failed ({rc}); the reference panics here")
    return proofs, commits


def verify_rangeproof(range_proof_vec, commit_vec, prove_range, seed=None):
    """-> bool; raises for malformed input like the reference's Err(..)    (:149-191)"""
    rc = _c().range_verify(range_proof_vec, commit_vec, prove_range, seed)
    if rc < 0:
        raise RangeProofError(f"ProofError ({rc})")
    return bool(rc)
