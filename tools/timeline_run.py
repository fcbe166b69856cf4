#!/usr/bin/env python3
"""ROFL_TIMELINE=1 python tools/timeline_run.py OUT [groups] : per-stream kernel timeline (CUDA events) of ONE configs[1] proof + verification."""
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
for it in range(3):
    rc, p, c = api.range_prove(v, bl, 16, 64, 16, 7, bytes([it + 1] * 32)); assert rc == 0
    assert api.range_verify(p, c, 16, bytes(32)) == 1
    if it < 2:
        dump(b"/dev/null")
dump(out.encode())
