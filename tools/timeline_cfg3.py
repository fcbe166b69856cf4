#!/usr/bin/env python3
"""ROFL_TIMELINE=1 python tools/timeline_cfg3.py OUT [n_chunks] : kernel timeline of one rank's share of configs[3] (chunks of 2^18 values, no generator tables)."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); api = pkg.context(0)
out = sys.argv[1]; nch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m = (1 << 24) // 64; n = nch * m
rng = np.random.default_rng(1)
v = rng.uniform(-0.99, 0.99, n).astype(np.float32); bl = api.rnd_scalar_vec(b"\x05" * 32, n)
dump = api.lib.rofl_timeline_dump; dump.argtypes = [ctypes.c_char_p]; dump.restype = None
for it in range(2):
    t0 = time.perf_counter(); rc, p, c = api.range_prove_shard(v, bl, m, 40, nch, 8, 16, 7, bytes([it + 1] * 32)); t1 = time.perf_counter()
    ok = api.range_verify_shard(p, c, m, 40, 8, bytes(32)); t2 = time.perf_counter()
    assert rc == 0 and ok == 1
    print("iter %d: prove %.1f ms verify %.1f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
    dump(b"/dev/null" if it == 0 else out.encode())
