#!/usr/bin/env python3
"""Throughput of K concurrent callers of ONE context (each proving + verifying its own configs[1] update on its own lane of streams):
python tools/concurrent_clients.py [K ...]   (rofl_service proves several clients per process: bin/basic_client.rs:136-161)"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); api = pkg.context(0)
D = 62006; REPS = 6
rng = np.random.default_rng(1)
def work(k, out):
    v = rng.uniform(-255.9, 255.9, D).astype(np.float32); bl = api.rnd_scalar_vec(bytes([k + 1] * 32), D)
    for it in range(REPS):
        rc, p, c = api.range_prove(v, bl, 16, 64, 16, 7, bytes([it + 1] * 32)); assert rc == 0
        assert api.range_verify(p, c, 16, bytes([7] * 32)) == 1
    out[k] = True
work(0, {})                                   # warm-up: tables, block cache
for K in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4]:
    for g_ in ([3, 2, 1] if K > 1 else [3]):
        api.set_option("groups", g_)
        out = {}
        ths = [threading.Thread(target=work, args=(k, out)) for k in range(K)]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        dt = time.perf_counter() - t0
        print("callers %d groups/call %d: %.1f ms per (prove + verify) per caller, %.3f M elements/s in total" % (K, g_, dt / REPS * 1e3, K * REPS * D / dt / 1e6), flush=True)
