"""Fixed-point configuration: the reference selects it at compile time with cargo features (rofl_crypto/src/fp.rs:35-137);
here it is a process-wide runtime setting with the same defaults (N_BITS = 16, frac = 7, PRECOMP_BIAS = 8)."""
N_BITS = 16
FRAC = 7
PRECOMP_BIAS = 8
BSGS_N_BITS = 16
_BIAS = {8: 3, 16: 7, 32: 7, 64: 0}


def configure(n_bits=None, frac=None):
    """configure() -> reference defaults; configure(8|16|32|64, frac) -> the `fpN` + `fracK` feature pair."""
    global N_BITS, FRAC, PRECOMP_BIAS, BSGS_N_BITS
    if n_bits is None:
        N_BITS, FRAC, PRECOMP_BIAS, BSGS_N_BITS = 16, 7 if frac is None else frac, 8, 16
        return
    if n_bits not in _BIAS:
        raise ValueError("n_bits must be 8, 16, 32 or 64")
    N_BITS, FRAC = n_bits, 7 if frac is None else frac
    PRECOMP_BIAS = _BIAS[n_bits]
    BSGS_N_BITS = n_bits if n_bits <= 16 else 16            # fp.rs:84-85,100-101
