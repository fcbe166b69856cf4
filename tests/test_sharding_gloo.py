"""N > 1 path on CPU: two processes, `gloo` backend, each rank driving the product's engine through the C ABI (the CUDA-on-CPU
emulation build tests/hostsim/libemul.so stands in for the GPU).  Chunk-sharded proving must give the single-process bytes,
sharded verification must AND the verdicts, and the parameter-axis aggregate + discrete log must match."""
import ctypes as C
import importlib.util
import os
import subprocess
import sys
import numpy as np
import pytest
from conftest import locked_make

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _load_ffi():
    spec = importlib.util.spec_from_file_location("rofl_ffi", os.path.join(ROOT, "rofl-project-code_b200", "_ffi.py"))
    ffi = importlib.util.module_from_spec(spec); spec.loader.exec_module(ffi)
    spec2 = importlib.util.spec_from_file_location("rofl_sharding", os.path.join(ROOT, "rofl-project-code_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec2); spec2.loader.exec_module(sh)
    return ffi, sh


def _inputs():
    sys.path.insert(0, ROOT)
    import oracle
    rng = np.random.default_rng(42)
    D, rb, P, nb = 11, 8, 4, 16                       # D' = 16, 4 chunks of 4 values; the last chunk is partly padding
    mn, mx = oracle.clip_bounds(rb, nb, 7)
    v = rng.uniform(mn, mx, D).astype(np.float32)
    bl = oracle.rnd_scalar_vec(b"\x51" * 32, D)
    return D, rb, P, nb, v, bl, bytes([5] * 32)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ffi, sh = _load_ffi()
        api = ffi.Api(C.CDLL(os.path.join(HERE, "hostsim", "libemul.so")))
        D, rb, P, nb, v, bl, seed = _inputs()
        rc, proofs, commits = sh.prove_range_sharded(api, v, bl, rb, P, nb, 7, seed, dist=dist)
        ok = sh.verify_range_sharded(api, proofs, commits, rb, seed, dist=dist)
        bad = proofs.copy(); bad[3, 70] ^= 1                      # corrupt the LAST chunk: only rank 1 sees it, every rank must learn the verdict
        ok_bad = sh.verify_range_sharded(api, bad, commits, rb, seed, dist=dist)
        x = np.array([[0.25, 1.25, -1.5, 2.0, -0.5], [-0.75, 1.25, -2.0, 1.0, 0.5], [0.5, 1.25, -3.0, -1.0, 0.25]], np.float32)
        cs = np.stack([api.commit(r, None, 16, 7) for r in x])
        rc_d, f = sh.decrypt_sharded(api, cs, 0, 1 << 10, 16, 16, 7, dist=dist)
        # a value outside the range in rank 1's slice only (element 9 lies in chunk 2): EVERY rank must report ValueOutOfRangeError (2), and a
        # negative error of one rank must win over the zeros of the others (MAX alone turned -98 / -1 into "ok")
        v2 = v.copy(); v2[9] = 5.0
        rc_oor = sh.prove_range_sharded(api, v2, bl, rb, P, nb, 7, seed, dist=dist)[0]
        v3 = v.copy(); v3[9] = np.nan
        rc_nan = sh.prove_range_sharded(api, v3, bl, rb, P, nb, 7, seed, dist=dist)[0]
        q.put((rank, rc, proofs.tobytes(), commits.tobytes(), ok, ok_bad, rc_d, f.tolist(), rc_oor, rc_nan))
        api.close()
    finally:
        dist.destroy_process_group()


def test_chunk_sharded_prove_verify_decrypt_two_ranks(oracle):
    import torch.multiprocessing as mp
    locked_make(os.path.join(HERE, "hostsim"), "libemul.so", env={"CXX": "g++"})
    D, rb, P, nb, v, bl, seed = _inputs()
    rc_o, p_o, c_o = oracle.range_prove(v, bl, rb, P, nb, 7, seed)
    assert rc_o == 0
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=600) for _ in range(2))
    for p in procs: p.join(timeout=60)
    for rank, rc, pb, cb, ok, ok_bad, rc_d, f, rc_oor, rc_nan in res:
        assert rc == 0 and rc_oor == 2 and rc_nan == -98
        assert pb == np.asarray(p_o).tobytes(), "sharded proofs differ from the single-process oracle bytes"
        assert cb == np.asarray(c_o).tobytes()
        assert ok == 1 and ok_bad == 0
        assert rc_d == 0 and f == [0.0, 3.75, -6.5, 2.0, 0.25]


def test_shard_layout_helpers():
    _, sh = _load_ffi()
    assert sh.chunk_layout(62006, 64) == (65536, 64, 1024)
    assert sh.chunk_layout(5, 64) == (8, 8, 1)
    cover = []
    for r in range(8):
        s = sh.shard_of(11689512, 64, r, 8)
        assert s["n_chunks"] == 8 and s["chunk_len"] == 2**18
        cover.append((s["elem_begin"], s["elem_end"]))
    assert cover[0][0] == 0 and cover[-1][1] == 11689512 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert sh.split_range(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
