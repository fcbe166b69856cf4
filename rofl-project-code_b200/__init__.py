"""rofl_b200 -- B200-native (sm_100a CUDA) implementation of RoFL's rofl_crypto client-prove / server-verify hot path.

The directory is named `rofl-project-code_b200`; import it through `__graft_entry__.load_package()` (registers the
package as `rofl_b200`).  Sub-modules mirror the reference's Rust module paths:
    rofl_b200.fp, conversion32, pedersen_ops, range_proof_vec, l2_range_proof_vec, square_proof_vec, bsgs32
All work is done by `librofl_b200.so` (C ABI in include/rofl_b200.h).  There is NO CPU fallback: importing works
anywhere, but creating a context raises unless the library is built and a CUDA device is present."""
import ctypes as _C
import os as _os
import subprocess as _sp

from ._ffi import Api, RoflError, EXPORTED_SYMBOLS, bind  # noqa: F401

_DIR = _os.path.dirname(_os.path.abspath(__file__))
LIB_PATH = _os.path.join(_DIR, "librofl_b200.so")
_lib = None
_ctx = {}


def build(jobs=8):
    """Compile every CUDA translation unit for sm_100a (nvcc cross-compiles without a GPU)."""
    _sp.check_call(["make", "-C", _os.path.join(_DIR, "csrc"), "-s", "-j", str(jobs)])
    return LIB_PATH


def load_library():
    global _lib
    if _lib is None:
        if not _os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                              "rofl_b200 has no CPU fallback.")
        _lib = bind(_C.CDLL(LIB_PATH))
    return _lib


def context(device=None):
    """Process-wide context for `device` (default: LOCAL_RANK or 0).  Raises RoflError without a CUDA device."""
    if device is None:
        device = int(_os.environ.get("LOCAL_RANK", "0"))
    if device not in _ctx:
        _ctx[device] = Api(load_library(), device)
    return _ctx[device]


from . import fp, conversion32, pedersen_ops, range_proof_vec, l2_range_proof_vec, square_proof_vec, compressed_rand_proof, rand_proof_vec, square_rand_proof_vec, bsgs32, sharding, params  # noqa: E402,F401
