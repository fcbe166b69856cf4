"""rofl_crypto::pedersen_ops (pedersen_ops.rs:9-127) and the ElGamal right halves (rand_proof/el_gamal.rs:57-69).
Points are (D, 32) uint8 arrays of compressed ristretto255 encodings, scalars (D, 32) little-endian."""
import numpy as np
from . import fp


def _c():
    from . import context
    return context()


def commit_vec_f32(values, blinding_vec):              # commit_vec(f32_to_scalar_vec(values), blindings), pedersen_ops.rs:18-25
    return _c().commit(values, blinding_vec, fp.N_BITS, fp.FRAC)


def commit_no_blinding_vec_f32(values):                # pedersen_ops.rs:9-16
    return _c().commit(values, None, fp.N_BITS, fp.FRAC)


def elgamal_commit_vec_f32(values, blinding_vec):      # ElGamalGens::commit per element: (L, R), el_gamal.rs:57-62
    return _c().commit(values, blinding_vec, fp.N_BITS, fp.FRAC, want_R=True)


def add_rp_vec_vec(rp_vec_vec):                        # pedersen_ops.rs:61-69
    return _c().aggregate(np.asarray(rp_vec_vec, dtype=np.uint8), 0)


def accumulate_unity(rp_vec_vec):                      # EncModelParamsAccumulator starting at ElGamalPair::unity (params.rs:81-124,165-179)
    return _c().aggregate(np.asarray(rp_vec_vec, dtype=np.uint8), 1)


def discrete_log_vec_table(rp_vec, table):             # pedersen_ops.rs:47-53 -> scalars
    rc, s, _ = _c().dlog(rp_vec, table.size, fp.BSGS_N_BITS, fp.N_BITS, fp.FRAC)
    if rc:
        raise RuntimeError("discrete log not found (the reference panics here, bsgs32.rs:69-70)")
    return s


def default_discrete_log_vec(rp_vec):                  # pedersen_ops.rs:27-35
    from .bsgs32 import BSGSTable
    return discrete_log_vec_table(rp_vec, BSGSTable.default())


def rnd_scalar_vec(length, seed=None):                # pedersen_ops.rs:124-127 (seeded ChaCha20 instead of thread_rng)
    return _c().rnd_scalar_vec(seed, length)


def zero_scalar_vec(length):                           # pedersen_ops.rs:102-104
    return np.zeros((length, 32), np.uint8)
