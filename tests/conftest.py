import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def locked_make(directory, *targets, env=None, jobs=None):
    """`make` under an exclusive file lock: pytest-xdist workers would otherwise rebuild the same library at the same time
    (one of them then loads a half-written .so)."""
    import fcntl, subprocess
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            cmd = ["make", "-C", directory, "-s"] + (["-j", str(jobs)] if jobs else []) + list(targets)
            subprocess.check_call(cmd, env={**os.environ, **(env or {})})
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def sodium():
    """libsodium (ristretto255) shipped inside pyzmq's wheel: an independent anchor for the oracle."""
    import ctypes, glob
    import zmq  # noqa: F401  (locates site-packages)
    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(zmq.__file__)), "pyzmq.libs", "libsodium*.so*"))
    if not cands:
        pytest.skip("libsodium not found")
    lib = ctypes.CDLL(cands[0])
    if lib.sodium_init() < 0:
        pytest.skip("sodium_init failed")
    return lib
