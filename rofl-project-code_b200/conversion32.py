"""rofl_crypto::conversion32 (conversion32.rs:11-64)."""
from . import fp


def _c():
    from . import context
    return context()


def f32_to_scalar_vec(values):                 # conversion32.rs:21-23
    return _c().f32_to_scalar_vec(values, fp.N_BITS, fp.FRAC)


def scalar_to_f32_vec(scalars):                # conversion32.rs:36-38
    return _c().scalar_to_f32_vec(scalars, fp.N_BITS, fp.FRAC)


def get_clip_bounds(range_bits):               # conversion32.rs:56-60
    return _c().clip_bounds(range_bits, fp.N_BITS, fp.FRAC)


def get_l2_clip_bounds(range_bits):            # conversion32.rs:62-64
    return _c().l2_clip_bound(range_bits, fp.N_BITS, fp.FRAC)
