import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package(); api = pkg.context(0); lib = api.lib
rng = np.random.default_rng(1); D = 62006
v = rng.uniform(-200, 200, D).astype(np.float32); bl = api.rnd_scalar_vec(b"\x01"*32, D)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
prof = len(sys.argv) > 2
for it in range(n):
    if prof: lib.rofl_prof_enable(1); lib.rofl_prof_reset()
    t0 = time.perf_counter()
    rc, p, c = api.range_prove(v, bl, 16, 64, 16, 7, bytes([it % 250 + 1])*32)
    w = (time.perf_counter() - t0) * 1e3
    extra = " rt_ms=%.1f (%d launches) frz_ms=%.1f tail_ms=%.1f" % (lib.rofl_prof_ms(4), lib.rofl_prof_launches(4), lib.rofl_prof_ms(6), lib.rofl_prof_ms(5)) if prof else ""
    print("iter", it, rc, "wall_ms=%.1f" % w + extra, flush=True)
