#!/usr/bin/env python
"""bench.py -- Hamming kNN-2 matching throughput on B200 (BASELINE.json metric: G Hamming cmp/s).

    python bench.py --gpus 1 --steps K --warmup W                      # our arm (default workload c5)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W      # sharded train set, one rank per GPU
    python bench.py --impl reference --steps K --warmup W              # the reference's CPU matcher, same metric

A "step" is one pass of the hot path over one batch of synthetic descriptors:
  c5 (default)  loop-closure query: 2000 queries vs a 10M-descriptor keyframe DB, kNN-2 + ratio 0.7;
                the DB is sharded in contiguous row blocks over the N ranks (strong scaling), per-rank
                top-2 keys are all-gathered with NCCL and merged.
  c4            BoW word assignment: 1M descriptors vs a 64k-word binary vocabulary, vocab sharded.
  c3 / c2 / c1  single-GPU shapes (keyframe batch 2016 x (2000x2000); 2000x20000 + cross-check; 1000x1000).
`value` times the device-resident path with CUDA events; `e2e` times the public host-buffer call
(pinned host arrays in, host results out, copies inside the timed region).  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "slam-1_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "G Hamming cmp/s (kNN-2 ratio match)"
UNIT = "Gcmp/s"
L2_BYTES = 126 * 1024 * 1024

WORKLOADS = {
    "c5": dict(name="c5 loop-closure query: 2000 x 10M-descriptor keyframe DB, kNN-2 + ratio 0.7, DB row-sharded",
               nq=2000, nt=10_000_000, ratio=(7, 10), cross=False, sharded=True),
    "c4": dict(name="c4 BoW word assignment: 1M descriptors x 64k-word binary vocabulary, vocab row-sharded",
               nq=1_000_000, nt=65536, ratio=None, cross=False, sharded=True),
    "c3": dict(name="c3 local-mapping batch: 64 keyframes x 2000 descriptors, 2016 unordered pairs, ratio 0.7",
               nq=2000, nt=2000, frames=64, ratio=(7, 10), cross=False, sharded=False),
    "c2": dict(name="c2 tracking vs local map: 2000 x 20000, kNN-2 + cross-check", nq=2000, nt=20000,
               ratio=None, cross=True, sharded=False),
    "c1": dict(name="c1 frame-to-frame: 1000 x 1000, kNN-2 + ratio 0.75", nq=1000, nt=1000, ratio=(3, 4),
               cross=False, sharded=False),
    # degenerate small-query streaming shapes for north_star's HBM clause
    "h1": dict(name="h1 streaming: 1 query x 10M descriptors (HBM-bound)", nq=1, nt=10_000_000, ratio=(7, 10),
               cross=False, sharded=False),
    "h4": dict(name="h4 streaming: 4 queries x 10M descriptors (HBM/POPC crossover)", nq=4, nt=10_000_000,
               ratio=(7, 10), cross=False, sharded=False),
}
BLOCK = 1_000_000   # synthetic train rows are generated in seeded 1M-row blocks


# ------------------------------------------------------------------------------------------------
# synthetic data: uniform train rows, half of the queries planted
# ------------------------------------------------------------------------------------------------
def train_rows(cfg_id: int, first: int, last: int) -> np.ndarray:
    out = np.empty((last - first, 32), dtype=np.uint8)
    b0, b1 = first // BLOCK, (last - 1) // BLOCK if last > first else first // BLOCK
    for b in range(b0, b1 + 1):
        lo, hi = max(first, b * BLOCK), min(last, (b + 1) * BLOCK)
        if hi <= lo:
            continue
        rng = np.random.default_rng(1000 * cfg_id + b)
        blk = rng.integers(0, 256, size=(BLOCK, 32), dtype=np.uint8)
        out[lo - first:hi - first] = blk[lo - b * BLOCK:hi - b * BLOCK]
    return out


def queries(cfg_id: int, nq: int, nt: int) -> np.ndarray:
    """Half of the queries are a train row from the first block with ~6 % of the bits flipped."""
    rng = np.random.default_rng(1000 * cfg_id + 999)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    head = train_rows(cfg_id, 0, min(nt, BLOCK))
    n_pl = min(nq // 2, head.shape[0])
    rows = rng.choice(nq, size=n_pl, replace=False)
    src = rng.choice(head.shape[0], size=n_pl, replace=False)
    noise = np.packbits(rng.random((n_pl, 256)) < 0.06, axis=1, bitorder="little")
    q[rows] = head[src] ^ noise
    return q


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def sample_now(self):
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit and name != "gpu_idle":
                    self.reasons.add(name)
        except Exception:
            pass

    def reset(self):
        self.samples, self.reasons = [], set()

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self.sample_now()
        self._stop.set()
        if self._thread:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's matcher family on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def cpu_matcher():
    """(callable(q, t) -> None, kind, description, cores).  cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) is the
    exhaustive member of the OpenCV matcher family the reference calls (tracking.py:17-22); without cv2 the
    C oracle port (OpenMP) is timed instead."""
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)

        def run(q, t):
            for s in range(0, t.shape[0], 262143):     # OpenCV's per-image row limit (matchers.cpp:860)
                bf.knnMatch(q, t[s:s + 262143], k=2)
        return run, "reference", f"cv2 {cv2.__version__} BFMatcher(NORM_HAMMING).knnMatch(k=2)", cv2.getNumThreads()
    except Exception:
        from oracle import oracle as orc

        def run(q, t):
            orc.c_knn2(q, t)
        return run, "port", "oracle/hamming_knn2.c (OpenMP)", orc.c_num_threads()


def cpu_sample(w, cfg_id, target_s: float):
    """Pick a bounded sample of the workload that takes ~target_s on this host."""
    run, kind, desc, cores = cpu_matcher()
    nq_s = min(w["nq"], 2000)
    q = queries(cfg_id, nq_s, w["nt"]) if w["nq"] <= 2000 else queries(cfg_id, 2000, w["nt"])
    nt_probe = min(w["nt"], 131072)
    t = train_rows(cfg_id, 0, nt_probe)
    run(q[:256], t[:4096])   # warm the thread pool
    t0 = time.perf_counter()
    run(q, t)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = nq_s * nt_probe / dt
    nt_s = int(min(w["nt"], max(nt_probe, rate * target_s / nq_s)))
    t = train_rows(cfg_id, 0, nt_s)
    return run, kind, desc, cores, q, t


# ------------------------------------------------------------------------------------------------
def reference_arm(args, w, cfg_id):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, kind, desc, cores, q, t = cpu_sample(w, cfg_id, target_s=3.0)
    for _ in range(args.warmup):
        run(q, t)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(q, t)
    dt = time.perf_counter() - t0
    cmp_per_step = q.shape[0] * t.shape[0]
    val = cmp_per_step * args.steps / dt / 1e9
    sample = f"{q.shape[0]} queries x first {t.shape[0]} train rows of the workload per step"
    # the same matcher on ONE host thread, on a slice sized for about two seconds
    single = None
    try:
        import cv2
        n_thr = cv2.getNumThreads()
        cv2.setNumThreads(1)
        try:
            rows_1 = int(max(1, min(t.shape[0], val * 1e9 / max(cores, 1) * 2.0 / q.shape[0])))
            t1 = time.perf_counter()
            run(q, t[:rows_1])
            d1 = max(time.perf_counter() - t1, 1e-6)
            single = {"value": q.shape[0] * rows_1 / d1 / 1e9, "unit": UNIT, "cores": 1,
                      "sample": f"{q.shape[0]} queries x first {rows_1} train rows, one pass"}
        finally:
            cv2.setNumThreads(n_thr)
    except Exception:
        pass
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if w["sharded"] else "replicas", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": {"workload": w["name"], "sample": sample, "matcher": desc},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "single_thread": single},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def ours(args, w, cfg_id):
    import torch
    import torch.distributed as dist
    import slammatch
    from slammatch import _lib
    from slammatch.sharded import ShardedMatcher, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout, and honours NCCL_DEBUG_FILE only above the VERSION level: raise the
        # level to WARN and send the log to stderr so that the JSON line is the only thing on stdout
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    sharded = w["sharded"] and world > 1
    if world > 1 and not w["sharded"]:
        # path does not shard: independent replicas
        pass

    ctx = _lib.context(local)
    ctx.set_variant(args.variant)
    nq, nt = w["nq"], w["nt"]
    num, den = w["ratio"] if w["ratio"] else (0, 1)

    # ---- inputs (host, pinned) and device-resident copies ----------------------------------------
    first, last = shard_bounds(nt, world)[rank] if sharded else (0, nt)
    if args.workload == "c3":
        frames = w["frames"]
        base = queries(cfg_id, nq, nq)
        rng = np.random.default_rng(3000)
        desc_h = np.stack([base ^ np.packbits(rng.random((nq, 256)) < 0.02 * (1 + f % 5), axis=1, bitorder="little")
                           for f in range(frames)])
        pairs = np.array([(i, j) for i in range(frames) for j in range(i + 1, frames)], dtype=np.int32)
        desc_pin = torch.from_numpy(desc_h).pin_memory()
        desc_d = desc_pin.to(dev)
        P = pairs.shape[0]
        idx_d = torch.empty((P, nq, 2), dtype=torch.int32, device=dev)
        dist_d = torch.empty((P, nq, 2), dtype=torch.int32, device=dev)
        acc_d = torch.empty((P, nq), dtype=torch.uint8, device=dev)
        cmp_per_step = float(P) * nq * nq
        in_bytes = desc_h.nbytes
    else:
        q_h = queries(cfg_id, nq, nt)
        t_h = train_rows(cfg_id, first, last)
        q_pin = torch.from_numpy(q_h).pin_memory()
        t_pin = torch.from_numpy(t_h).pin_memory()
        q_d = q_pin.to(dev)
        t_d = t_pin.to(dev)
        cmp_per_step = float(nq) * nt            # whole job, all ranks together
        in_bytes = q_h.nbytes + t_h.nbytes
        sm = ShardedMatcher(t_d, first, ratio=w["ratio"], variant=args.variant, exchange=args.exchange)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def step_device():
        if args.workload == "c3":
            _lib.check(ctx.lib.slm_knn2_batched(ctx.handle, desc_d.data_ptr(), frames, nq, pairs.ctypes.data, P, num, den,
                                                idx_d.data_ptr(), dist_d.data_ptr(), acc_d.data_ptr(), stream))
            return idx_d, dist_d, acc_d
        if sharded:
            return sm.knn2(q_d)
        return slammatch.knn2(q_d, t_d, ratio=w["ratio"], cross_check=w["cross"])

    # L2 hygiene: inputs larger than L2 stream from HBM every step; smaller workloads get an L2 flush
    flush = None
    if in_bytes < 2 * L2_BYTES:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # The clock sampler starts BEFORE the warm-up: its first NVML queries take milliseconds and contend with
    # kernel submission for the driver lock, which would otherwise land inside the timed region.
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()

    # ---- timed region: device-resident inputs, CUDA events on the launching stream -----------------
    ctx.profile(True)
    ctx.profile_read()
    launches0 = ctx.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.reset()
    # one more untimed step queued directly in front of the start event: the barrier above idles the GPU for
    # a moment and the first kernel after an idle gap runs at a lower clock.  The timed region is still
    # exactly K steps between two events on the launching stream, with a barrier + synchronize on both sides.
    step_device()
    launches0 = ctx.launch_count()
    wall0 = time.perf_counter()
    if flush is None:
        evs[0][0].record()
        for _ in range(args.steps):
            out = step_device()
        evs[0][1].record()
        sampler.sample_now()          # the GPU is still working through the queued steps here
        barrier()
        dev_ms = evs[0][0].elapsed_time(evs[0][1])
    else:
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            evs[k][0].record()
            out = step_device()
            evs[k][1].record()
        barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    kern_ms, kern_n = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launch_count() - launches0
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms.item())
    jobs = world if (world > 1 and not w["sharded"]) else 1     # replicas: N independent copies of the job
    value = cmp_per_step * jobs * args.steps / (dev_ms * 1e-3) / 1e9
    matched = int(out[2].sum().item())

    # ---- e2e: the public host-buffer call, H2D and D2H inside the timed region ---------------------
    def step_e2e():
        if args.workload == "c3":
            desc_d.copy_(desc_pin, non_blocking=True)
            step_device()
            return idx_d.cpu(), dist_d.cpu(), acc_d.cpu()
        if sharded:
            q_d.copy_(q_pin, non_blocking=True)
            t_d.copy_(t_pin, non_blocking=True)
            i, d, a = sm.knn2(q_d)
            return i.cpu(), d.cpu(), a.cpu()
        return slammatch.knn2(q_pin.numpy(), t_pin.numpy(), ratio=w["ratio"], cross_check=w["cross"], device=local)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_val = cmp_per_step * jobs * e2e_steps / float(t_e.item()) / 1e9
    out_bytes = (P * nq if args.workload == "c3" else nq) * 17

    # The object-level drop-in call the unmodified reference makes (tracking.py:22): matcher.knnMatch(des1, des2, k=2)
    # returning tuples of cv2.DMatch -- same copies as e2e plus the construction of 2 * nq result objects.
    matcher_line = None
    if world == 1 and args.workload in ("c1", "c2"):
        m = slammatch.Matcher(crossCheck=False, device=local)
        qn, tn = q_pin.numpy(), t_pin.numpy()
        for _ in range(2):
            m.knnMatch(qn, tn, k=2)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            rows = m.knnMatch(qn, tn, k=2)
        dt = (time.perf_counter() - t0) / e2e_steps
        matcher_line = {"value": cmp_per_step / dt / 1e9, "unit": UNIT, "us_per_call": dt * 1e6, "rows": len(rows),
                        "api": "slammatch.Matcher().knnMatch(des1, des2, k=2) -> tuple[nq] of tuple[2] of cv2.DMatch"}

    # Same call with the train set held in a persistent device-resident collection (OpenCV's matcher.add([...]) +
    # knnMatch(q, k) form; slammatch.KeyframeDB): the DB is uploaded once, outside the timed region, and every step
    # moves only the query descriptors in and the results out.  Reported NEXT TO e2e, never instead of it.
    resident = None
    if world == 1 and args.workload in ("c5", "c4"):
        db = slammatch.KeyframeDB(device=local, capacity=nt)
        db.add(t_d)
        qn = q_pin.numpy()
        for _ in range(2):
            db.query(qn, ratio=w["ratio"])
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            db.query(qn, ratio=w["ratio"])
        torch.cuda.synchronize(dev)
        resident = {"value": cmp_per_step * e2e_steps / (time.perf_counter() - t0) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": int(q_h.nbytes), "d2h_bytes_per_step": int(out_bytes),
                    "api": "slammatch.KeyframeDB.add(train) once, then .query(host queries) per step"}
        del db

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        variant = ctx.last_variant()
        kernel_name = ctx.last_kernel()
        per_rank_cmp = cmp_per_step / (world if sharded else 1)
        kern_avg_ms = kern_ms / max(kern_n, 1)
        kernels_per_step = max(kern_n // (args.steps + 1), 1)   # the profile also holds the ramp step
        cmp_per_launch = per_rank_cmp / kernels_per_step
        if variant == "tensor":
            # one comparison = a 256-term dot product of +-1 fp8 values = 512 flop on the tcgen05 pipe
            ach = cmp_per_launch * 512 / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
            bf16 = peaks.get("bf16_tflops", 1590.0)
            probe = {}
            try:
                probe = json.load(open(os.path.join(ROOT, "profiles", "peaks_probe.json")))
            except Exception:
                pass
            # The kernel issues tcgen05.mma kind::f8f6f4.  MEASURED_PEAKS.json only holds a cuBLAS bf16 figure,
            # so the denominator is the tensor pipe's own fp8 ceiling: this repo's back-to-back tcgen05 probe
            # measured 8191 MAC/clk/SM (profiles/r1_tc_probe_v2.txt) = the architectural 8192; times 2 flop,
            # the SM count and the MAX SM clock (an upper bound: the kernel cannot clock higher), or 2 x the
            # measured bf16 figure if that is larger.
            macs = probe.get("tcgen05_f8f6f4_mac_per_clk_per_sm", 8192)
            clk_mhz = clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak = max(2.0 * bf16, macs * 2.0 * sms * clk_mhz * 1e6 / 1e12)
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": (ach / peak) if ach else None, "traffic": None,
                    "kernel": kernel_name or "knn2_tc2_kernel", "kernel_ms": kern_avg_ms,
                    "peak_note": ("max(2 x measured cuBLAS bf16 burst %.1f TF/s [%s], tcgen05 kind::f8f6f4 ceiling = %d MAC/clk/SM "
                                  "(own probe) x 2 x %d SMs x %.0f MHz max SM clock)"
                                  % (bf16, "MEASURED_PEAKS.json" if peaks else "fallback", macs, sms, clk_mhz)),
                    "frac_vs_2x_bf16_measured": (ach / (2.0 * bf16)) if ach else None,
                    "algorithmic_unit": "1 cmp = 512 fp8 flop (256-bit +-1 dot product)"}
        elif kernel_name == "knn2_stream_kernel":
            # nq <= 8 runs knn2_stream_kernel: the train set streams through the SMs once
            alg_bytes = 32.0 * (nq + nt / (world if sharded else 1)) + 16.0 * nq
            ach = alg_bytes / (kern_avg_ms * 1e-3) / 1e9 if kern_avg_ms > 0 else None
            peak = peaks.get("hbm_gbs", 6650.0)
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                    "traffic": None, "kernel": "knn2_stream_kernel", "kernel_ms": kern_avg_ms,
                    "peak_note": ("measured copy bandwidth (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"),
                    "algorithmic_bytes": alg_bytes}
        elif variant == "bmma":
            # b1 mma.sync is emulated by ptxas on sm_100a (8 IMMA + ~100 logic/move instructions per MMA): the
            # kernel is issue-slot-bound (ncu: issue active 70 %, legacy tensor pipe 30 %), no single pipe peak applies
            ach = cmp_per_launch / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
            roof = {"bound": "issue_slots(b1 mma.sync emulation)", "achieved": ach, "peak": None, "unit": "Tcmp/s",
                    "frac": None, "traffic": None, "kernel": "knn2_bmma_kernel", "kernel_ms": kern_avg_ms,
                    "peak_note": "no native b1 tensor instruction on sm_100a; see profiles/r1_ncu_bmma_c5.txt"}
        else:
            # integer pipe: 8 POPC32 per comparison on the XU pipe; measured 15.8 lanes/clk/SM (profiles/r1_pipe_rates.txt)
            probe = {}
            try:
                probe = json.load(open(os.path.join(ROOT, "profiles", "peaks_probe.json")))
            except Exception:
                pass
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            popc_rate = probe.get("popc32_lanes_per_clk_per_sm", 16.0) * sms * (clocks.get("sm_max_mhz") or 1965.0) * 1e6
            ach = cmp_per_launch / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
            peak = popc_rate / 8 / 1e12
            roof = {"bound": "int_popc(xu pipe)", "achieved": ach, "peak": peak, "unit": "Tcmp/s",
                    "frac": (ach / peak) if ach else None, "traffic": None, "kernel": kernel_name or "knn2_popc_kernel",
                    "kernel_ms": kern_avg_ms,
                    "peak_note": "measured POPC32 lanes/clk/SM (own probe) x SMs x max SM clock / 8 POPC per cmp"}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                roof["traffic"] = json.load(open(tr)).get(roof["kernel"], {}).get(args.workload)
            except Exception:
                pass

        cpu = None
        if world == 1 and not args.no_cpu:
            run, kind, desc, cores, qs, ts = cpu_sample(w, cfg_id, target_s=12.0)
            t0 = time.perf_counter()
            run(qs, ts)
            dt = time.perf_counter() - t0
            cpu = {"value": qs.shape[0] * ts.shape[0] / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{qs.shape[0]} queries x first {ts.shape[0]} train rows, one pass ({dt:.1f} s)",
                   "matcher": desc}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if w["sharded"] else "replicas", "vs_baseline": None,
            "dtype": "e4m3(+-1 bits), f32 accumulate (exact)" if variant == "tensor" else "u32",
            "data": "synthetic",
            "config": {"workload": w["name"], "nq": nq, "nt": nt, "variant": variant, "variant_requested": args.variant,
                       "l2": "inputs larger than L2 (streamed from HBM every step)" if flush is None
                             else "256 MiB L2 flush between timed steps",
                       "parallelism": (f"train rows sharded over {world} ranks, exchange of packed top-2 keys ({sm.last_exchange}) + merge"
                                       if sharded else ("single GPU" if world == 1 else f"{world} replicas"))},
            "matched_queries_per_s": matched * jobs / (dev_ms / args.steps * 1e-3),
            "matched_per_step": matched,
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(out_bytes),
                    "steps": e2e_steps, "api": "slammatch.knn2(host arrays) -> slm_knn2_host" if not (sharded or args.workload == "c3")
                    else "pinned host -> device copy + device entry points + result read-back",
                    "resident_db": resident, "matcher_knnMatch": matcher_line},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="auto", choices=["auto", "popc", "tensor", "bmma"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "nvlink", "a2a"],
                    help="sharded path: how per-rank keys are exchanged (auto = NVLink peer stores when available)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    cfg_id = int(args.workload[1]) + (5 if args.workload[0] == "h" else 0)
    if args.impl == "reference":
        reference_arm(args, w, cfg_id)
    else:
        ours(args, w, cfg_id)


if __name__ == "__main__":
    main()
