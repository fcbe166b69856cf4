"""rofl_crypto::compressed_rand_proof (compressed_rand_proof/mod.rs:43-158).  CompressedRandProof = 128 bytes
(C'_L | C'_R | z_m | z_r, mod.rs:108-114); the commitments are D ElGamal pairs of 64 bytes (L | R, rand_proof/el_gamal.rs:105-111)."""
from . import fp

MAX_PAIRS = 900000            # rows of the reference's per-index label table (generate_unique_u8_triplets.py:2)


class ProofError(Exception):
    """compressed_rand_proof/errors.rs"""


def _c():
    from . import context
    return context()


def helper_prove(m_vec, r_vec, seed=None):                                   # :134-140
    """-> (proof[128], c_vec[D, 64]); L_i = commit(m_i, r_i), R_i = r_i B"""
    rc, proof, pairs = _c().crp_prove(m_vec, None, r_vec, fp.N_BITS, fp.FRAC, seed)
    if rc:
        raise ProofError(f"error {rc}")
    return proof, pairs


def helper_prove_existing(m_vec, m_com, r_vec, seed=None):                   # :141-148
    """-> (proof[128], c_vec[D, 64]); L_i = m_com[i] (already published), R_i = r_i B"""
    rc, proof, pairs = _c().crp_prove(m_vec, m_com, r_vec, fp.N_BITS, fp.FRAC, seed)
    if rc:
        raise ProofError(f"error {rc}")
    return proof, pairs


def helper_verify(proof, c_vec):                                              # :149-158
    """-> bool; raises ProofError('FormatError') for malformed input like CompressedRandProof::from_bytes"""
    rc = _c().crp_verify(proof, c_vec)
    if rc < 0:
        raise ProofError("FormatError" if rc == -1 else f"error {rc}")
    return bool(rc)
