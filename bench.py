#!/usr/bin/env python3
"""bench.py -- range-proved commitment elements/s (prove & verify) on N B200s (BASELINE.json metric).

A step = one pass of the hot path over one synthetic client update of BASELINE.json configs[1] (cifar_lenet5,
D = 62 006 parameters, L-inf 16-bit range proofs, n_partition 64, fp16/frac7): commit + create_rangeproof, then
verify_rangeproof.  elements/s = D / (t_prove + t_verify).
  value : inputs already resident in HBM (rofl_range_prove_dev / rofl_range_verify_dev), CUDA-event timed
  e2e   : the reference-facing host-buffer calls (rofl_range_prove / rofl_range_verify) from pinned host memory,
          host<->device copies inside the timed region
N > 1 (torchrun): one process per GPU, each rank proves and verifies its own client's update (sharding by client, no
data-path collective; the proof bytes are all-gathered to every rank over NCCL outside the kernels) -> weak scaling.
--impl reference : the CPU restatement of the reference (oracle/, OpenMP over chunks, all host cores) on a bounded
sample of the same workload; the real Rust reference cannot be built here (no cargo), see DESIGN.md.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(workload="cifar_lenet5 update: commit + L-inf range proofs, prove and verify", D=62006, range_bits=16, n_partition=64, n_bits=16, frac=7)
METRIC = "range-proved commitment elements/s (prove & verify)"
IMAD_PER_FIELD_MUL = 72                        # 8 x 8 limb schoolbook (64) + the 2^256 = 38 fold (8): the algorithmic cost of one field multiplication


def synth(D, rng_bits, n_bits, frac, seed):
    rng = np.random.default_rng(seed)
    mx = ((1 << (rng_bits - 1)) - 1) / float(1 << frac)
    v = rng.uniform(-mx, mx, D).astype(np.float32)
    return np.clip(v, -mx, mx)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.lines, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.p.stdout], daemon=True); self.t.start()
            t0 = time.time()                     # nvidia-smi's start-up (NVML attach to every GPU of the box) stalls CUDA calls of running
            while not self.lines and time.time() - t0 < 15 and self.p.poll() is None:      # processes: let it finish before any timing
                time.sleep(0.05)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        print("clock samples (sm MHz):", sm, file=sys.stderr)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(steps, warmup, sample_chunks=None):
    """CPU restatement of the reference on a bounded sample: `cores` chunks of the real chunk size (1024 values, 16 bit)."""
    import oracle
    oracle.set_num_threads(os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses all host threads it can
    cores = oracle.num_threads()
    w = WORKLOAD
    Dp = 1 << (w["D"] - 1).bit_length()
    m = Dp // w["n_partition"]
    chunks = sample_chunks or max(1, min(cores, w["n_partition"]))
    Ds = m * chunks
    v = synth(Ds, w["range_bits"], w["n_bits"], w["frac"], 1)
    bl = oracle.rnd_scalar_vec(b"\x01" * 32, Ds)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        rc, p, c = oracle.range_prove(v, bl, w["range_bits"], chunks, w["n_bits"], w["frac"], bytes([it % 256] * 32))
        t1 = time.perf_counter()
        ok = oracle.range_verify(p, c, w["range_bits"], bytes(32))
        t2 = time.perf_counter()
        assert rc == 0 and ok == 1
        if it >= warmup:
            times.append((t1 - t0, t2 - t1))
    tp = float(np.mean([a for a, _ in times])); tv = float(np.mean([b for _, b in times]))
    return dict(value=Ds / (tp + tv), prove_eps=Ds / tp, verify_eps=Ds / tv, cores=cores, sample=f"{chunks} chunks x {m} values ({Ds} elements) of the same workload, OpenMP over chunks", ms=(tp + tv) * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOAD
    # stdout carries exactly ONE JSON line: everything any library prints to fd 1 meanwhile (NCCL's version banner at NCCL_DEBUG=VERSION
    # ignores NCCL_DEBUG_FILE) is sent to stderr, and the line is written to the saved descriptor at the end
    sys.stdout.flush(); json_fd = os.dup(1); os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(max(1, args.steps), min(args.warmup, 1))
        line = dict(metric=METRIC, value=r["value"], unit="elements/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=r["ms"], higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="u32 limbs (GF(2^255-19), mod l) on CPU u64", data="synthetic", impl="reference", config=w,
                    cpu_baseline=dict(value=r["value"], unit="elements/s", cores=r["cores"], kind="port", sample=r["sample"], prove_eps=r["prove_eps"], verify_eps=r["verify_eps"]),
                    e2e=dict(value=r["value"], unit="elements/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        emit(line); return

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep NCCL's own version / debug lines off stdout: stdout carries the one JSON line
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api = pkg.context(local)                       # raises without the CUDA library / device: no CPU fallback
    lib = api.lib
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 16)
    host_threads = int(os.environ.get("BENCH_HOST_THREADS", "0")) or max(2, min(32, ncpu // max(1, world)))
    lib.rofl_set_host_threads(api.h, host_threads)  # the ranks of one box share its cores: split them for the transcript threads
    if rank == 0:
        print("host cpus %d, ranks %d -> %d transcript threads per rank" % (ncpu, world, host_threads), file=sys.stderr)
    stream = torch.cuda.ExternalStream(lib.rofl_ctx_stream(api.h), device=torch.device("cuda", local))
    D, rb, P, nb, fr = w["D"], w["range_bits"], w["n_partition"], w["n_bits"], w["frac"]
    v_h = torch.from_numpy(synth(D, rb, nb, fr, 1000 + rank)).pin_memory()
    bl_h = torch.from_numpy(api.rnd_scalar_vec(bytes([rank + 1] * 32), D)).pin_memory()
    v_d, bl_d = v_h.cuda(), bl_h.cuda()
    commits_d = torch.empty((D, 32), dtype=torch.uint8, device="cuda")
    commits_h = np.zeros((D, 32), np.uint8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")           # > 126 MB L2
    n_proofs, plen = api.range_proof_shape(D, rb, P)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(it):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        flush.fill_(it & 0xff); torch.cuda.synchronize()
        e0.record(stream)
        rc, proofs = api.range_prove_dev(v_d.data_ptr(), bl_d.data_ptr(), D, rb, P, nb, fr, bytes([it % 251 + 1] * 32), commits_d.data_ptr())
        e1.record(stream)
        ok = api.range_verify_dev(proofs, commits_d.data_ptr(), D, rb, bytes(32))
        e2.record(stream); e2.synchronize()
        assert rc == 0 and ok == 1
        return e0.elapsed_time(e1), e1.elapsed_time(e2), proofs

    def step_e2e(it):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        flush.fill_(it & 0xff); torch.cuda.synchronize()
        e0.record(stream)
        rc, proofs, commits = api.range_prove(v_h.numpy(), bl_h.numpy(), rb, P, nb, fr, bytes([it % 251 + 1] * 32))
        e1.record(stream)
        ok = api.range_verify(proofs, commits, rb, bytes(32))
        e2.record(stream); e2.synchronize()
        assert rc == 0 and ok == 1
        return e0.elapsed_time(e1), e1.elapsed_time(e2)

    sampler = ClockSampler(local)
    if not os.environ.get("BENCH_NO_CLOCKS"):          # (diagnostic switch: how much does the sampling itself perturb the steps?)
        sampler.start()      # started BEFORE the warm-up: nvidia-smi's own start-up (NVML init) stalls the driver for a moment
    for it in range(args.warmup):
        step_resident(it); step_e2e(it)
    sampler.lines.clear()                               # keep only the samples taken during the timed regions
    imad_peak = lib.rofl_probe_imad_wide(api.h) if rank == 0 else 0.0       # roofline denominator, measured on this GPU before the timed region
    print("== timed region starts", file=sys.stderr, flush=True)
    # ---- timed: resident (no per-kernel events inside the timed region)
    lib.rofl_prof_enable(0); lib.rofl_prof_reset()
    barrier()
    t_p = t_v = 0.0
    per_step = []
    for it in range(args.steps):
        if os.environ.get("ROFL_ALLOC_TRACE") or os.environ.get("ROFL_JITTER"):
            print("== resident step", it, "t=%.0f ms" % (time.monotonic() * 1e3), file=sys.stderr, flush=True)
        a, b, proofs = step_resident(100 + it); t_p += a; t_v += b; per_step.append((round(a, 2), round(b, 2)))
    barrier()
    if rank == 0:
        print("resident per-step (prove_ms, verify_ms):", per_step, file=sys.stderr)
    launches = lib.rofl_prof_launches(-1)
    # ---- timed: end to end through the host-buffer API
    barrier()
    e_p = e_v = 0.0
    per_step = []
    for it in range(args.steps):
        a, b = step_e2e(200 + it); e_p += a; e_v += b; per_step.append((round(a, 2), round(b, 2)))
    barrier()
    if rank == 0:
        print("e2e per-step (prove_ms, verify_ms):", per_step, file=sys.stderr)
    clocks = sampler.stop()
    print("== profiling pass", file=sys.stderr, flush=True)
    # ---- per-kernel breakdown: the same K resident steps again with CUDA events around every launch of the main kernel families
    # (separate pass: creating / recording the events costs host time that must not leak into `value`)
    api.set_option("groups", 1)             # one chunk group: every kernel is timed ALONE on the GPU (several groups overlap kernels of different rounds)
    step_resident(299); torch.cuda.synchronize()      # untimed: the one-group scratch sizes are new to the block cache
    lib.rofl_prof_enable(1); lib.rofl_prof_reset()
    prof_step_ms = 0.0
    prof_steps = []
    for it in range(args.steps):
        k0 = sum(lib.rofl_prof_ms(i) for i in range(7))
        a, b, _ = step_resident(300 + it); prof_step_ms += a + b
        torch.cuda.synchronize()
        prof_steps.append((round(a + b, 2), round(sum(lib.rofl_prof_ms(i) for i in range(7)) - k0, 2)))
    prof_step_ms /= args.steps
    if rank == 0:      # slow steps with an unchanged kernel sum = the GPU was waiting, not computing more slowly
        print("profiling pass per-step (step_ms, event-timed kernel ms of the instrumented families):", prof_steps, file=sys.stderr)
    torch.cuda.synchronize()
    api.set_option("groups", int(os.environ.get("BENCH_GROUPS", "3")))
    prof = dict(fold_ms=lib.rofl_prof_ms(0), msm_ms=lib.rofl_prof_ms(1), commit_ms=lib.rofl_prof_ms(2), rt_ms=lib.rofl_prof_ms(4), tail_ms=lib.rofl_prof_ms(5),
                rt_launches=lib.rofl_prof_launches(4), rt_madds=lib.rofl_prof_work(4))
    lib.rofl_prof_enable(0)
    if world > 1:      # proof pieces are the only thing that crosses NVLink: gather them (outside the timed kernels)
        pt = torch.from_numpy(proofs).cuda(); out = [torch.empty_like(pt) for _ in range(world)]; dist.all_gather(out, pt)
        tt = torch.tensor([t_p, t_v, e_p, e_v], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX); t_p, t_v, e_p, e_v = tt.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    K = args.steps
    ms_step = (t_p + t_v) / K
    value = world * D / (ms_step / 1e3)
    e2e_val = world * D / ((e_p + e_v) / K / 1e3)
    # ---- roofline of the dominant kernel: k_rt_msm (direct table MSM: S, the unfolded IPP rounds, the verifier's generator part).
    # The path is bound by the IMAD.WIDE.U32 pipe (32 lanes/clk/SM), not by HBM or the tensor cores.
    #   achieved = algorithmic multiply-adds / measured kernel time;  algorithmic = mixed additions x 7 field multiplications x 72
    #   IMAD.WIDE.U32 (8x8 schoolbook + 8 for the 2^256 = 38 fold);  peak = IMAD.WIDE issue rate measured now by rofl_probe_imad_wide
    madds = prof["rt_madds"]; rt_ms = prof["rt_ms"]
    achieved = madds * 7 * IMAD_PER_FIELD_MUL / (rt_ms / 1e3) / 1e12 if rt_ms > 0 else None
    peak = imad_peak / 1e12 if imad_peak else None
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_rt_msm_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    launches_rt = max(1, prof["rt_launches"])
    roofline = dict(bound="int-imad-wide (integer multiply pipe; not hbm / tensor: see DESIGN.md section 5)", kernel="k_rt_msm", achieved=achieved, peak=peak, unit="T IMAD.WIDE.U32/s",
                    frac=(achieved / peak) if (achieved and peak) else None, traffic=traffic,
                    algorithmic=dict(mixed_additions_per_launch=madds / launches_rt, field_muls_per_addition=7, imad_wide_per_field_mul=IMAD_PER_FIELD_MUL, launches_per_step=launches_rt / K),
                    kernel_ms_per_launch=rt_ms / launches_rt, kernel_ms_per_step=rt_ms / K, kernel_share_of_step=rt_ms / K / prof_step_ms,
                    share_note="share of the one-chunk-group step of the profiling pass (%.1f ms): the timed steps overlap three chunk groups" % prof_step_ms,
                    peak_source="rofl_probe_imad_wide: best of two mad.wide.u32 patterns (rotating multiplicands; carry-chained as in the field multiply), all SMs, best of 5, run before the timed region",
                    hbm=dict(achieved_gbs=madds * 96 / (rt_ms / 1e3) / 1e9 if rt_ms > 0 else None, peak_gbs=hbm_peak, note="table gathers: 96 B per mixed addition; MEASURED_PEAKS.json copy bandwidth" if peaks else "fallback 6650 GB/s"),
                    other_kernels_ms_per_step=dict(bucket_msm=prof["msm_ms"] / K, generator_fold=prof["fold_ms"] / K, ipp_tail=prof["tail_ms"] / K, commit=prof["commit_ms"] / K))
    line = dict(metric=METRIC, value=value, unit="elements/s", n_gpus=world, steps=K, warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="u32 (8 saturated 32-bit limbs of GF(2^255-19), IMAD.WIDE.U32 with 64-bit accumulate; scalars mod l)", data="synthetic",
                config=dict(w, l2_flush="256 MiB write between timed iterations", parallelism=f"client x{world}", generators="warm (cached tables)"),
                prove_eps=world * D / (t_p / K / 1e3), verify_eps=world * D / (t_v / K / 1e3), prove_ms=t_p / K, verify_ms=t_v / K,
                e2e=dict(value=e2e_val, unit="elements/s", h2d_bytes_per_step=int(v_h.numel() * 4 + bl_h.numel() + D * 32), d2h_bytes_per_step=int(D * 32 + n_proofs * plen),
                         prove_ms=e_p / K, verify_ms=e_v / K),
                gpu_launches=int(launches), clocks=clocks, roofline=roofline)
    if not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(1, 0)
            line["cpu_baseline"] = dict(value=r["value"], unit="elements/s", cores=r["cores"], kind="port", sample=r["sample"], prove_eps=r["prove_eps"], verify_eps=r["verify_eps"])
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = dict(value=None, unit="elements/s", cores=0, kind="port", sample=f"failed: {ex}")
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
