"""rofl_crypto::square_proof_vec (square_proof_vec/mod.rs:19-160).  SquareProof = 160 bytes
(C'_l | C'_sq | z_m | z_r1 | z_r2, square_proof/mod.rs:118-125), SquareProofCommitments = 64 bytes (c_l | c_sq)."""
from . import fp


class L2RangeProofError(Exception):
    """square_proof_vec/errors.rs"""


def _c():
    from . import context
    return context()


def create_l2rangeproof_vec_existing(value_vec, value_com_vec, random_vec, random_vec_2, seed=None):      # :19-75
    if len(value_vec) != len(random_vec):
        raise L2RangeProofError("WrongNumBlindingFactors")
    rc, proofs, commits = _c().square_prove(value_vec, value_com_vec, random_vec, random_vec_2, fp.N_BITS, fp.FRAC, seed)
    if rc:
        raise L2RangeProofError(f"error {rc}")
    return proofs, commits


def verify_l2rangeproof_vec(randproof_vec, commit_vec):                                                   # :130-160
    if len(randproof_vec) != len(commit_vec):
        raise L2RangeProofError("WrongNumberOfElGamalPairs")
    rc = _c().square_verify(randproof_vec, commit_vec)
    if rc < 0:
        raise L2RangeProofError("ProofError::FormatError")
    return bool(rc)
