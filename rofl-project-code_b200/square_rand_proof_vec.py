"""rofl_crypto::square_rand_proof_vec (square_rand_proof_vec/mod.rs:18-160): per-element square + randomness proofs of the un-optimised L2
encoding (enc type 3).  SquareRandProof = 192 bytes (C'.L | C'.R | C'_sq | z_m | z_r1 | z_r2, square_rand_proof/mod.rs:118-125);
SquareRandProofCommitments = 96 bytes (c.L | c.R | c_sq, square_rand_proof/pedersen.rs:21-30)."""
from . import fp


class L2RangeProofError(Exception):
    """square_rand_proof_vec/errors.rs"""


def _c():
    from . import context
    return context()


def create_l2rangeproof_vec(value_vec, random_vec, random_vec_2, seed=None):                              # :72-127
    if len(value_vec) != len(random_vec):
        raise L2RangeProofError("WrongNumBlindingFactors")
    rc, proofs, commits = _c().square_rand_prove(value_vec, None, random_vec, random_vec_2, fp.N_BITS, fp.FRAC, seed)
    if rc:
        raise L2RangeProofError(f"error {rc}")
    return proofs, commits


def create_l2rangeproof_vec_existing(value_vec, value_com_vec, random_vec, random_vec_2, seed=None):      # :18-70
    if len(value_vec) != len(random_vec):
        raise L2RangeProofError("WrongNumBlindingFactors")
    rc, proofs, commits = _c().square_rand_prove(value_vec, value_com_vec, random_vec, random_vec_2, fp.N_BITS, fp.FRAC, seed)
    if rc:
        raise L2RangeProofError(f"error {rc}")
    return proofs, commits


def verify_l2rangeproof_vec(randproof_vec, commit_vec):                                                    # :129-160
    if len(randproof_vec) != len(commit_vec):
        raise L2RangeProofError("WrongNumberOfElGamalPairs")
    rc = _c().square_rand_verify(randproof_vec, commit_vec)
    if rc < 0:
        raise L2RangeProofError("ProofError::FormatError")
    return bool(rc)
