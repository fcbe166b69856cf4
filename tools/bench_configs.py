#!/usr/bin/env python3
"""Throughput of every BASELINE.json config on one GPU (informational; bench.py stays the contract line for configs[1]).
Writes profiles/r01_configs.json.  Times are wall-clock around the host-buffer C ABI calls (median of `reps` runs after one warm-up)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); api = pkg.context(0)


def med(f, reps=5):
    f(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


rng = np.random.default_rng(123); out = {}
# configs[0]: mnist 5k, 8-bit, P=64, fp16/7
D = 5000; v = rng.uniform(-0.99, 0.99, D).astype(np.float32); bl = api.rnd_scalar_vec(b"\x01" * 32, D); seed = b"\x02" * 32
rc, p, c = api.range_prove(v, bl, 8, 64, 16, 7, seed)
out["configs[0] mnist_dev_intrinsic_5k (8-bit, P=64)"] = dict(D=D, prove_ms=1e3 * med(lambda: api.range_prove(v, bl, 8, 64, 16, 7, seed)), verify_ms=1e3 * med(lambda: api.range_verify(p, c, 8, seed)))
# configs[1]: lenet5 62k, 16-bit
D = 62006; v = rng.uniform(-255.9, 255.9, D).astype(np.float32); bl = api.rnd_scalar_vec(b"\x03" * 32, D)
rc, p, c = api.range_prove(v, bl, 16, 64, 16, 7, seed)
out["configs[1] cifar_lenet5 (16-bit, P=64)"] = dict(D=D, prove_ms=1e3 * med(lambda: api.range_prove(v, bl, 16, 64, 16, 7, seed)), verify_ms=1e3 * med(lambda: api.range_verify(p, c, 16, seed)))
# configs[2]: resnet18 intrinsic 50k, L2: whole client message / whole server check (EncParamsL2Compressed)
D = 50000; v = (rng.integers(-24, 25, D) / 128).astype(np.float32); bl = api.rnd_scalar_vec(b"\x04" * 32, D)
rc, m = api.enc_l2_compressed_encrypt(v, bl, 8, 64, 32, 32, 7, seed)
out["configs[2] resnet18_intrinsic_50k (L2: 8-bit range + 50k square proofs + 32-bit sum proof + compressed rand proof)"] = dict(
    D=D, prove_ms=1e3 * med(lambda: api.enc_l2_compressed_encrypt(v, bl, 8, 64, 32, 32, 7, seed), 3), verify_ms=1e3 * med(lambda: api.enc_l2_compressed_verify(m, seed), 3))
# configs[3]: resnet18 full, one GPU's share of eight (8 chunks of 2^18 values, no generator tables at this size)
mlen = (1 << 24) // 64; n = 8 * mlen; v = rng.uniform(-0.99, 0.99, n).astype(np.float32); bl = api.rnd_scalar_vec(b"\x05" * 32, n)
rc, p, c = api.range_prove_shard(v, bl, mlen, 40, 8, 8, 16, 7, seed)
out["configs[3] resnet18 full, one rank of 8 (8 chunks x 2^18 values, 8-bit)"] = dict(
    D=n, prove_ms=1e3 * med(lambda: api.range_prove_shard(v, bl, mlen, 40, 8, 8, 16, 7, seed), 2), verify_ms=1e3 * med(lambda: api.range_verify_shard(p, c, mlen, 40, 8, seed), 2))
# configs[4]: server: verify one client message (above) + aggregate 48 clients + decrypt
D = 50000; pts = np.stack([api.commit((rng.integers(-24, 25, D) / 128).astype(np.float32), None, 32, 7) for _ in range(48)])
agg = api.aggregate(pts, 0)
out["configs[4] server: aggregate 48 x 50k + bsgs decrypt (table 2^16)"] = dict(D=D, clients=48, aggregate_ms=1e3 * med(lambda: api.aggregate(pts, 0), 3), dlog_ms=1e3 * med(lambda: api.dlog(agg, 1 << 16, 16, 32, 7), 3))
for k, r in out.items():
    if "prove_ms" in r:
        r["prove_elements_per_s"] = r["D"] / (r["prove_ms"] / 1e3); r["verify_elements_per_s"] = r["D"] / (r["verify_ms"] / 1e3)
    print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in r.items()}, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
