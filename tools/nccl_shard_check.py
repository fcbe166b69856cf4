#!/usr/bin/env python3
"""Run under torchrun on N GPUs: chunk-sharded create_rangeproof / verify_rangeproof of ONE update and the parameter-axis
aggregate + decrypt over NCCL (rofl-project-code_b200/sharding.py), compared with the single-GPU result on rank 0."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = g.load_package(); api = pkg.context(local); sh = pkg.sharding
dev = torch.device("cuda", local)
rng = np.random.default_rng(9)
D, rb, P, nb = 62006, 16, 64, 16
mx = ((1 << (rb - 1)) - 1) / 128.0
v = rng.uniform(-mx, mx, D).astype(np.float32); bl = api.rnd_scalar_vec(b"\x71" * 32, D); seed = b"\x72" * 32
rc, proofs, commits = sh.prove_range_sharded(api, v, bl, rb, P, nb, 7, seed, dist=dist, device=dev)
ok = sh.verify_range_sharded(api, proofs, commits, rb, seed, dist=dist, device=dev)
bad = proofs.copy(); bad[P - 1, 100] ^= 1
ok_bad = sh.verify_range_sharded(api, bad, commits, rb, seed, dist=dist, device=dev)
x = [(rng.integers(-24, 25, 5000) / 128).astype(np.float32) for _ in range(3)]
cs = np.stack([api.commit(r, None, 16, 7) for r in x])
rc_d, f = sh.decrypt_sharded(api, cs, 0, 1 << 16, 16, 16, 7, dist=dist, device=dev)
if rank == 0:
    rc1, p1, c1 = api.range_prove(v, bl, rb, P, nb, 7, seed)
    same = rc == rc1 == 0 and (p1 == proofs).all() and (c1 == commits).all()
    print(f"world={world} sharded==single-GPU bytes: {same}; verify={ok} tampered={ok_bad}; decrypt rc={rc_d} exact={(f == np.sum(np.stack(x), axis=0, dtype=np.float64).astype(np.float32)).all()}")
    assert same and ok == 1 and ok_bad == 0 and rc_d == 0
dist.barrier(); dist.destroy_process_group()
