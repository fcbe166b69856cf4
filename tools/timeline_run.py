#!/usr/bin/env python3
"""ROFL_TIMELINE=1 python tools/timeline_run.py OUT [repetitions] : per-stream kernel timeline (CUDA events) of ONE configs[1] proof + verification."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); api = pkg.context(0)
out = sys.argv[1]
D = 62006
rng = np.random.default_rng(1)
v = rng.uniform(-255.9, 255.9, D).astype(np.float32); bl = api.rnd_scalar_vec(b"\x01" * 32, D)
dump = api.lib.rofl_timeline_dump; dump.argtypes = [ctypes.c_char_p]; dump.restype = None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # one file per timed repetition: OUT, OUT.1, OUT.2, ..
for it in range(2 + reps):
    rc, p, c = api.range_prove(v, bl, 16, 64, 16, 7, bytes([it + 1] * 32)); assert rc == 0
    assert api.range_verify(p, c, 16, bytes(32)) == 1
    dump(b"/dev/null" if it < 2 else (out if it == 2 else "%s.%d" % (out, it - 2)).encode())
