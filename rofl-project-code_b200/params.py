"""rofl_service::flserver::params -- the two optimised encodings end to end (params.rs:699-743 EncParamsRangeCompressed,
:787-885 EncParamsL2Compressed, :181-291 EncModelParams::verify).  A message is a dict of the wire fields of flservice.proto
(`enc_values`, `rand_proof`, `range_proof`, `square_proof`, `square_range_proof`, `range_bits`, `l2_range_bits`) as uint8 arrays in the
fixed-width layouts of SURVEY Appendix C; the protobuf framing itself is the service's business."""
from . import fp


def _c():
    from . import context
    return context()


class EncParamsRangeCompressed:
    @staticmethod
    def encrypt(plaintext_vec, blinding_vec, prove_range, n_partition, seed=None):
        rc, msg = _c().enc_range_compressed_encrypt(plaintext_vec, blinding_vec, prove_range, n_partition, fp.N_BITS, fp.FRAC, seed)
        if rc:
            raise RuntimeError(f"encrypt failed ({rc}); the reference unwrap()s here")
        return msg

    @staticmethod
    def verify(msg, check_percentage=1.0, seed=None):
        return bool(_c().enc_range_compressed_verify(msg, check_percentage, seed))


class EncParamsL2Compressed:
    @staticmethod
    def encrypt(plaintext_vec, blinding_vec, prove_range, n_partition, l2_range, seed=None):
        rc, msg = _c().enc_l2_compressed_encrypt(plaintext_vec, blinding_vec, prove_range, n_partition, l2_range, fp.N_BITS, fp.FRAC, seed)
        if rc:
            raise RuntimeError(f"encrypt failed ({rc}); the reference unwrap()s here")
        return msg

    @staticmethod
    def verify(msg, seed=None):
        return bool(_c().enc_l2_compressed_verify(msg, seed))


class EncParamsRange:
    """params.rs:467-510 / :186-203 -- un-optimised L-inf encoding: one RandProof per element"""
    @staticmethod
    def encrypt(plaintext_vec, blinding_vec, prove_range, n_partition, check_percentage=1.0, seed=None):
        rc, msg = _c().enc_range_encrypt(plaintext_vec, blinding_vec, prove_range, n_partition, check_percentage, fp.N_BITS, fp.FRAC, seed)
        if rc:
            raise RuntimeError(f"encrypt failed ({rc}); the reference unwrap()s here")
        return msg

    @staticmethod
    def verify(msg, check_percentage=1.0, seed=None):
        return bool(_c().enc_range_verify(msg, check_percentage, seed))


class EncParamsL2:
    """params.rs:607-646 / :205-233 -- un-optimised L2 encoding: one SquareRandProof per element"""
    @staticmethod
    def encrypt(plaintext_vec, blinding_vec, prove_range, n_partition, l2_range, seed=None):
        rc, msg = _c().enc_l2_encrypt(plaintext_vec, blinding_vec, prove_range, n_partition, l2_range, fp.N_BITS, fp.FRAC, seed)
        if rc:
            raise RuntimeError(f"encrypt failed ({rc}); the reference unwrap()s here")
        return msg

    @staticmethod
    def verify(msg, seed=None):
        return bool(_c().enc_l2_verify(msg, seed))
