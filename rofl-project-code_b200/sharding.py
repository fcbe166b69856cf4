"""Multi-GPU sharding of the prove / verify path: one process per GPU, `torch.distributed` for the plumbing.

The path shards without any data-path collective:
  * range proofs -- the P chunks of one update are independent proofs with independent transcripts
    (range_proof_vec/mod.rs:54-78,168-181): contiguous chunk ranges per rank, generator tables replicated;
  * per-element proofs (square proofs), aggregation and discrete log -- contiguous index ranges on the parameter axis;
  * server side -- (client, chunk) work items are independent.
Only proof bytes / commitments (prove), one verdict per rank (verify) or the decrypted f32 slice (decrypt) cross NVLink,
through all_gather / all_reduce(MIN) on the process group the caller passes in (NCCL on GPUs, gloo in the CPU tests).
There is no group-element reduction over NCCL anywhere: the layout avoids it.
"""
import numpy as np


def next_pow2(v):
    return 1 if v <= 1 else 1 << (int(v) - 1).bit_length()


def chunk_layout(D, n_partition):
    """(padded length D', number of chunks C, chunk length m) exactly as create_rangeproof lays an update out (range_proof_vec/mod.rs:46-55)."""
    Dp = next_pow2(D); C = min(Dp, n_partition)
    return Dp, C, Dp // C


def split_range(n, world):
    """Contiguous, balanced split of range(n) over `world` ranks: list of (begin, end)."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def shard_of(D, n_partition, rank, world):
    """Chunk range and element range of `rank`: dict(chunk_begin, n_chunks, chunk_len, elem_begin, elem_end) (elem range clipped to the real D)."""
    Dp, C, m = chunk_layout(D, n_partition)
    c0, c1 = split_range(C, world)[rank]
    return dict(chunk_begin=c0, n_chunks=c1 - c0, chunk_len=m, elem_begin=min(D, c0 * m), elem_end=min(D, c1 * m))


def _all_gather_bytes(dist, group, arr, device):
    """all_gather of variable-length uint8 arrays (lengths are exchanged first); returns the list of per-rank numpy arrays."""
    import torch
    world = dist.get_world_size(group)
    n = torch.tensor([arr.size], dtype=torch.int64, device=device)
    ns = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    mx = max(1, int(max(int(x.item()) for x in ns)))
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    if arr.size:
        buf[:arr.size] = torch.from_numpy(np.ascontiguousarray(arr.reshape(-1))).to(device)
    outs = [torch.zeros(mx, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return [o[:int(k.item())].cpu().numpy() for o, k in zip(outs, ns)]


_PINNED = {}


def _pinned(name, nbytes):
    """A cached pinned host buffer (torch) of at least nbytes: host<->device copies of the gathered commitments run at full PCIe / NVLink-C2C speed."""
    import torch
    t = _PINNED.get(name)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8)
        try:
            t = t.pin_memory()
        except Exception:  # noqa: BLE001  (no CUDA: the gloo tests)
            pass
        _PINNED[name] = t
    return t[:nbytes]


def prove_range_sharded(api, values, blind, range_bits, n_partition, n_bits, frac, seed, dist=None, group=None, device="cpu"):
    """create_rangeproof of ONE update over all ranks of `group`.  Every rank holds the full update (or at least its slice) and
    returns (rc, proofs[C, plen], commits[D, 32]) identical on all ranks and byte-identical to the single-GPU call.
    The gather moves every rank's chunk range as ONE equal-sized block (its padding rows are zero = the identity encoding the reference pads with,
    range_proof_vec/mod.rs:163-165), so the result lands in its final layout: pinned buffer -> device -> all_gather_into_tensor -> pinned buffer."""
    values = np.ascontiguousarray(values, np.float32); D = values.size
    blind = np.ascontiguousarray(blind, np.uint8).reshape(D, 32)
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist is not None else (0, 1)
    sh = shard_of(D, n_partition, rank, world)
    e0, e1 = sh["elem_begin"], sh["elem_end"]
    Dp, C, m = chunk_layout(D, n_partition)
    equal = world > 1 and C % world == 0              # every rank owns the same number of chunks: one fixed-size block per rank
    err, out_c = None, None
    if equal:
        rows = sh["n_chunks"] * m
        blk = _pinned("commit_block", rows * 32); blk.zero_()
        out_c = blk.numpy().reshape(rows, 32)[:e1 - e0]
    try:
        if sh["n_chunks"]:
            rc, proofs, commits = api.range_prove_shard(values[e0:e1], blind[e0:e1], sh["chunk_len"], sh["chunk_begin"], sh["n_chunks"], range_bits, n_bits, frac, seed, out_commits=out_c)
        else:
            rc, proofs, commits = 0, np.zeros((0, 0), np.uint8), np.zeros((0, 32), np.uint8)
    except Exception as ex:  # noqa: BLE001  (a CUDA error on one rank must not leave the others hanging in the collective below)
        err, rc, proofs, commits = ex, -100, np.zeros((0, 0), np.uint8), np.zeros((0, 32), np.uint8)
    if world == 1:
        if err is not None:
            raise err
        return rc, proofs, commits
    import torch
    # error codes are negative, the reference's domain errors positive, 0 = ok: the job fails if ANY rank failed (MIN finds errors, MAX domain errors)
    lo = torch.tensor([rc], dtype=torch.int32, device=device); hi = lo.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group); dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if err is not None:
        raise err
    if int(lo.item()) < 0 or int(hi.item()) != 0:
        return (int(lo.item()) if int(lo.item()) < 0 else int(hi.item())), None, None
    plen = api.range_proof_len(range_bits * sh["chunk_len"])
    parts_p = _all_gather_bytes(dist, group, proofs, device)
    all_p = np.concatenate([p.reshape(-1, plen) for p in parts_p if p.size], axis=0)
    if equal:
        src = blk.to(device, non_blocking=True)
        out = torch.empty(Dp * 32, dtype=torch.uint8, device=device)
        if hasattr(dist, "all_gather_into_tensor") and str(device) != "cpu":
            dist.all_gather_into_tensor(out, src, group=group)
        else:
            dist.all_gather(list(out.chunk(world)), src, group=group)
        host = _pinned("commit_all", Dp * 32); host.copy_(out)
        return 0, all_p, host.numpy().reshape(Dp, 32)[:D]
    parts_c = _all_gather_bytes(dist, group, commits, device)
    all_c = np.concatenate([c.reshape(-1, 32) for c in parts_c if c.size], axis=0)
    return 0, all_p, all_c


def verify_range_sharded(api, proofs, commits, range_bits, seed, dist=None, group=None, device="cpu"):
    """verify_rangeproof of ONE update over all ranks: each rank checks its chunk range, the verdicts are AND-ed with all_reduce(MIN).
    Returns 1 / 0 like the single-GPU call, or the (most negative) error code of any rank."""
    proofs = np.ascontiguousarray(proofs, np.uint8); proofs = proofs.reshape(proofs.shape[0], -1)
    commits = np.ascontiguousarray(commits, np.uint8).reshape(-1, 32)
    D, C = commits.shape[0], proofs.shape[0]
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist is not None else (0, 1)
    if world == 1:
        return api.range_verify(proofs, commits, range_bits, seed)
    m = next_pow2(D) // C                                        # range_proof_vec/mod.rs:168
    c0, c1 = split_range(C, world)[rank]
    e0, e1 = min(D, c0 * m), min(D, c1 * m)
    rc = api.range_verify_shard(proofs[c0:c1], commits[e0:e1], m, c0, range_bits, seed) if c1 > c0 and m > 0 else (1 if m > 0 else -3)
    import torch
    t = torch.tensor([rc], dtype=torch.int32, device=device); dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())


def decrypt_sharded(api, client_points, init_unity, table_size, bsgs_bits, n_bits, frac, dist=None, group=None, device="cpu"):
    """Homomorphic aggregate + discrete log sharded on the PARAMETER axis (rank r owns a contiguous index range for all clients),
    then all_gather of the decrypted f32 slices.  client_points: [clients, D, 32].  Returns (rc, f32[D])."""
    pts = np.ascontiguousarray(client_points, np.uint8); n_clients, D = pts.shape[0], pts.shape[1]
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist is not None else (0, 1)
    b, e = split_range(D, world)[rank]
    rc, f = 0, np.zeros(0, np.float32)
    if e > b:
        agg = api.aggregate(np.ascontiguousarray(pts[:, b:e]), init_unity)
        rc, _, f = api.dlog(agg, table_size, bsgs_bits, n_bits, frac)
    if world == 1:
        return rc, f
    import torch
    t = torch.tensor([rc], dtype=torch.int32, device=device); dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    parts = _all_gather_bytes(dist, group, np.ascontiguousarray(f, np.float32).view(np.uint8), device)
    return int(t.item()), np.concatenate([p.view(np.float32) for p in parts if p.size])
