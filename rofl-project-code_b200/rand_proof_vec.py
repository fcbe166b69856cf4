"""rofl_crypto::rand_proof_vec (rand_proof_vec/mod.rs:14-118): per-element ElGamal randomness proofs of the un-optimised range encoding
(enc type 2).  RandProof = 128 bytes (C'_L | C'_R | z_m | z_r, rand_proof/mod.rs:87-97); ElGamalPair = 64 bytes (L | R)."""
from . import fp


class RandProofError(Exception):
    """rand_proof_vec/errors.rs"""


def _c():
    from . import context
    return context()


def create_randproof_vec(value_vec, random_vec, seed=None):                                   # :14-49
    if len(value_vec) != len(random_vec):
        raise RandProofError("WrongNumBlindingFactors")
    rc, proofs, pairs = _c().rand_prove(value_vec, None, random_vec, fp.N_BITS, fp.FRAC, seed)
    if rc:
        raise RandProofError(f"error {rc}")
    return proofs, pairs


def create_randproof_vec_existing(value_vec, existing_value_com_vec, random_vec, seed=None):  # :51-89
    if len(value_vec) != len(random_vec):
        raise RandProofError("WrongNumBlindingFactors")
    rc, proofs, pairs = _c().rand_prove(value_vec, existing_value_com_vec, random_vec, fp.N_BITS, fp.FRAC, seed)
    if rc:
        raise RandProofError(f"error {rc}")
    return proofs, pairs


def verify_randproof_vec(randproof_vec, commit_vec):                                           # :91-118
    if len(randproof_vec) != len(commit_vec):
        raise RandProofError("WrongNumberOfElGamalPairs")
    rc = _c().rand_verify(randproof_vec, commit_vec)
    if rc < 0:
        raise RandProofError("ProofError::FormatError")
    return bool(rc)
