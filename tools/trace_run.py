import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package(); api = pkg.context(0)
rng = np.random.default_rng(1); D = 62006
v = rng.uniform(-200, 200, D).astype(np.float32); bl = api.rnd_scalar_vec(b"\x01"*32, D)
for it in range(4):
    rc, p, c = api.range_prove(v, bl, 16, 64, 16, 7, bytes([it+1])*32)
    print("iter", it, rc, flush=True)
