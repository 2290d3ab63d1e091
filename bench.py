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
    # the same job cut the other way (note): vocabulary replicated, query slices, no candidate exchange
    "c4q": dict(name="c4 BoW word assignment, QUERY-sharded alternative: 1M descriptors sliced over the ranks, 64k-word vocabulary replicated",
                nq=1_000_000, nt=65536, ratio=None, cross=False, sharded=True, shard_by="queries"),
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


def cpu_model() -> str:
    """Host CPU model and logical core count (stated next to every CPU figure)."""
    name = "unknown"
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.lower().startswith("model name"):
                    name = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return f"{name}; os.cpu_count() = {os.cpu_count()}"


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
                         "single_thread": single, "cpu": cpu_model()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# parity check outside the timed region: a slice of every rank's output against the C oracle on the FULL train set
# ------------------------------------------------------------------------------------------------
def _oracle():
    from oracle import oracle as orc          # test infrastructure: used here only as the checker
    orc.build_c()
    return orc


def parity_expected(wl_id, w, cfg_id, q_h, t_full, extra):
    """rank 0: (query rows checked, expected idx, dist, accept) from oracle/hamming_knn2.c."""
    orc = _oracle()
    nq = w["nq"]
    if wl_id == "c3":
        desc_h, pairs = extra
        sel_pairs = sorted({0, len(pairs) // 2, len(pairs) - 1})
        idx, dist, acc = [], [], []
        for p in sel_pairs:
            i, d = orc.c_knn2(desc_h[pairs[p][0]], desc_h[pairs[p][1]])
            idx.append(i); dist.append(d); acc.append(orc.c_ratio(d, *w["ratio"]))
        return np.asarray(sel_pairs), np.stack(idx), np.stack(dist), np.stack(acc)
    if w["cross"] or nq * w["nt"] <= 2_000_000_000 and nq <= 4096:
        sel = np.arange(nq)                                      # small problems: every query
    else:
        n = 64 if nq <= 4096 else 1024
        sel = np.unique(np.linspace(0, nq - 1, n).astype(np.int64))
    i, d = orc.c_knn2(q_h[sel], t_full)
    acc = orc.c_ratio(d, *w["ratio"]) if w["ratio"] else (i[:, 0] >= 0).astype(np.uint8)
    if w["cross"]:
        acc = acc & orc.c_cross_check(q_h, t_full, i)
    return sel, i, d, acc


# ------------------------------------------------------------------------------------------------
def measure(args, wl_id, env, headline):
    """One workload on the current process group: device-resident value, e2e, roofline, parity check."""
    import torch
    import torch.distributed as dist
    import slammatch
    from slammatch import _lib
    from slammatch.sharded import QueryShardedMatcher, ShardedMatcher, shard_bounds

    rank, world, local, dev, ctx, peaks = env["rank"], env["world"], env["local"], env["dev"], env["ctx"], env["peaks"]
    w = WORKLOADS[wl_id]
    cfg_id = int(wl_id[1]) + (5 if wl_id[0] == "h" else 0)
    steps = args.steps if headline else max(3, min(args.steps, args.config_steps))
    e2e_steps = max(1, min(steps, args.e2e_steps if headline else 2))
    sharded = w["sharded"] and world > 1
    by_q = sharded and w.get("shard_by") == "queries"
    nq, nt = w["nq"], w["nt"]
    num, den = w["ratio"] if w["ratio"] else (0, 1)
    ctx.set_variant(args.variant)

    # ---- inputs (host, pinned) and device-resident copies ----------------------------------------
    first, last = shard_bounds(nt, world)[rank] if (sharded and not by_q) else (0, nt)
    extra = None
    if wl_id == "c3":
        frames = w["frames"]
        base = queries(cfg_id, nq, nq)
        rng = np.random.default_rng(3000)
        desc_h = np.stack([base ^ np.packbits(rng.random((nq, 256)) < 0.02 * (1 + f % 5), axis=1, bitorder="little")
                           for f in range(frames)])
        pairs = np.array([(i, j) for i in range(frames) for j in range(i + 1, frames)], dtype=np.int32)
        extra = (desc_h, pairs)
        desc_pin = torch.from_numpy(desc_h).pin_memory()
        desc_d = desc_pin.to(dev)
        P = pairs.shape[0]
        idx_d = torch.empty((P, nq, 2), dtype=torch.int32, device=dev)
        dist_d = torch.empty((P, nq, 2), dtype=torch.int32, device=dev)
        acc_d = torch.empty((P, nq), dtype=torch.uint8, device=dev)
        # e2e: the caller's result buffers, pinned and reused from step to step
        out_pin = (torch.empty((P, nq, 2), dtype=torch.int32).pin_memory(), torch.empty((P, nq, 2), dtype=torch.int32).pin_memory(),
                   torch.empty((P, nq), dtype=torch.uint8).pin_memory())
        cmp_per_step = float(P) * nq * nq
        in_bytes = desc_h.nbytes
        q_h = t_h = None
    else:
        q_h = queries(cfg_id, nq, nt)
        t_h = train_rows(cfg_id, first, last)
        q_pin = torch.from_numpy(q_h).pin_memory()
        t_pin = torch.from_numpy(t_h).pin_memory()
        q_d = q_pin.to(dev)
        t_d = t_pin.to(dev)
        cmp_per_step = float(nq) * nt            # whole job, all ranks together
        in_bytes = (q_h.nbytes // world if by_q else q_h.nbytes) + t_h.nbytes      # what THIS rank uploads per e2e step
        sm = (QueryShardedMatcher(t_d, ratio=w["ratio"], variant=args.variant) if by_q else
              ShardedMatcher(t_d, first, ratio=w["ratio"], variant=args.variant, exchange=args.exchange, total_rows=nt))
    stream = torch.cuda.current_stream(dev).cuda_stream

    def step_device():
        if wl_id == "c3":
            _lib.check(ctx.lib.slm_knn2_batched(ctx.handle, desc_d.data_ptr(), frames, nq, pairs.ctypes.data, P, num, den,
                                                idx_d.data_ptr(), dist_d.data_ptr(), acc_d.data_ptr(), stream))
            return idx_d, dist_d, acc_d
        if sharded:
            return sm.knn2(q_d)
        return slammatch.knn2(q_d, t_d, ratio=w["ratio"], cross_check=w["cross"])

    # L2 hygiene: inputs larger than L2 stream from HBM every step; smaller workloads get an L2 flush
    flush = None
    if in_bytes < 2 * L2_BYTES:
        flush = env.get("flush")
        if flush is None:
            flush = env["flush"] = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

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
    # (the library's own kernel timing is OFF here: its event records sit between the search kernel and the
    # programmatically launched refine / merge kernels and would serialise them; the dominant kernel is timed in a
    # separate pass below)
    ctx.profile(False)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 2)]
    barrier()
    sampler.reset()
    # one more untimed step queued directly in front of the start event: the barrier above idles the GPU for
    # a moment and the first kernel after an idle gap runs at a lower clock.  The timed region is still
    # exactly K steps between two events on the launching stream, with a barrier + synchronize on both sides.
    step_device()
    launches0 = ctx.launch_count()
    wall0 = time.perf_counter()
    if flush is None:
        evs[0].record()
        for k in range(steps):
            out = step_device()
            evs[k + 1].record()            # also the start of step k + 1: per-step times for min / median
        sampler.sample_now()              # the GPU is still working through the queued steps here
        barrier()
        dev_ms = evs[0].elapsed_time(evs[steps])
        per_step = [evs[k].elapsed_time(evs[k + 1]) for k in range(steps)]
    else:
        for k in range(steps):
            flush.fill_(k & 0xFF)
            evs[2 * k].record()
            out = step_device()
            evs[2 * k + 1].record()
        barrier()
        per_step = [evs[2 * k].elapsed_time(evs[2 * k + 1]) for k in range(steps)]
        dev_ms = sum(per_step)
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    variant = ctx.last_variant()
    kernel_name = ctx.last_kernel()
    # ---- the dominant kernel, timed by events recorded inside the library around its launches (separate pass) ----
    ctx.profile(True)
    try:
        ctx.profile_read()
    except Exception:
        pass
    prof_steps = max(3, min(steps, 10))
    step_device()                          # ramp step, as above
    for k in range(prof_steps):
        if flush is not None:
            flush.fill_(k & 0xFF)
        step_device()
    barrier()
    kern_ms, kern_n = ctx.profile_read()
    ctx.profile(False)
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms.item())
    jobs = world if (world > 1 and not w["sharded"]) else 1     # replicas: N independent copies of the job
    value = cmp_per_step * jobs * steps / (dev_ms * 1e-3) / 1e9
    matched = int(out[2].sum().item())

    # ---- parity: a slice of THIS rank's output of the timed path against the oracle on the full train set ----
    parity = None
    if not args.no_parity:
        n_sel = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            t_full = t_h if (wl_id == "c3" or not sharded or by_q) else train_rows(cfg_id, 0, nt)
            sel, e_i, e_d, e_a = parity_expected(wl_id, w, cfg_id, q_h, t_full, extra)
            del t_full
            n_sel[0] = sel.shape[0]
        if world > 1:
            dist.broadcast(n_sel, 0)
        n = int(n_sel.item())
        shape_i = (n, nq, 2) if wl_id == "c3" else (n, 2)
        shape_a = (n, nq) if wl_id == "c3" else (n,)
        if rank == 0:
            sel_t = torch.from_numpy(np.ascontiguousarray(sel, dtype=np.int64)).to(dev)
            e_i_t, e_d_t = torch.from_numpy(e_i).to(dev), torch.from_numpy(e_d).to(dev)
            e_a_t = torch.from_numpy(np.ascontiguousarray(e_a, dtype=np.uint8)).to(dev)
        else:
            sel_t = torch.empty((n,), dtype=torch.int64, device=dev)
            e_i_t = torch.empty(shape_i, dtype=torch.int32, device=dev)
            e_d_t = torch.empty(shape_i, dtype=torch.int32, device=dev)
            e_a_t = torch.empty(shape_a, dtype=torch.uint8, device=dev)
        if world > 1:
            for x in (sel_t, e_i_t, e_d_t, e_a_t):
                dist.broadcast(x, 0)
        ok = bool(torch.equal(out[0][sel_t], e_i_t) and torch.equal(out[1][sel_t], e_d_t) and torch.equal(out[2][sel_t], e_a_t))
        ok_t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        parity = {"queries": (n * nq if wl_id == "c3" else n), "ok": bool(ok_t.item()), "ranks_checked": world,
                  "against": "oracle/hamming_knn2.c on the full train set (idx, dist, accept of the timed path's output)"}

    # ---- e2e: the public host-buffer call, H2D and D2H inside the timed region ---------------------
    def step_e2e():
        if wl_id == "c3":
            return slammatch.knn2_batched(desc_pin, pairs, ratio=w["ratio"], out=out_pin, device=local)
        if sharded:
            return sm.knn2_host(q_pin.numpy(), train_host=t_pin)
        return slammatch.knn2(q_pin.numpy(), t_pin.numpy(), ratio=w["ratio"], cross_check=w["cross"], device=local)

    def time_e2e(fn, n):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize(dev)
        t_e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        return cmp_per_step * jobs * n / float(t_e.item()) / 1e9

    e2e_val = time_e2e(step_e2e, e2e_steps)
    out_bytes = (P * nq if wl_id == "c3" else nq) * 17
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(out_bytes),
           "steps": e2e_steps,
           "api": ("slammatch.knn2(pinned host arrays) -> slm_knn2_host" if not (sharded or wl_id == "c3") else
                   "QueryShardedMatcher.knn2_host(host query slice, host train set)" if by_q else
                   "ShardedMatcher.knn2_host(host queries, host shard)" if sharded else
                   "slammatch.knn2_batched(pinned host descriptors, pairs, out=pinned host buffers) -> slm_knn2_batched")}
    if world == 1 and wl_id != "c3":
        # the drop-in caller hands PAGEABLE numpy arrays (orb.py:23-24): same call, inputs not pinned
        q_pg, t_pg = q_h.copy(), t_h.copy()
        e2e["pageable"] = {"value": time_e2e(lambda: slammatch.knn2(q_pg, t_pg, ratio=w["ratio"], cross_check=w["cross"],
                                                                     device=local), e2e_steps), "unit": UNIT}
        del q_pg, t_pg

    # The object-level drop-in call the unmodified reference makes (tracking.py:22): matcher.knnMatch(des1, des2, k=2)
    # returning tuples of cv2.DMatch -- same copies as e2e plus the construction of 2 * nq result objects.
    if world == 1 and wl_id in ("c1", "c2"):
        m = slammatch.Matcher(crossCheck=False, device=local)
        qn, tn = q_pin.numpy(), t_pin.numpy()
        for _ in range(2):
            m.knnMatch(qn, tn, k=2)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            rows = m.knnMatch(qn, tn, k=2)
        dt = (time.perf_counter() - t0) / e2e_steps
        e2e["matcher_knnMatch"] = {"value": cmp_per_step / dt / 1e9, "unit": UNIT, "us_per_call": dt * 1e6, "rows": len(rows),
                                   "api": "slammatch.Matcher().knnMatch(des1, des2, k=2) -> tuple[nq] of tuple[2] of cv2.DMatch"}

    # Same call with the train set held in a persistent device-resident collection (OpenCV's matcher.add([...]) +
    # knnMatch(q, k) form; slammatch.KeyframeDB): the DB is uploaded once, outside the timed region, and every step
    # moves only the query descriptors in and the results out.  Reported NEXT TO e2e, never instead of it.
    if world == 1 and wl_id in ("c5", "c4"):
        db = slammatch.KeyframeDB(device=local, capacity=nt)
        db.add(t_d)
        qn = q_pin.numpy()
        e2e["resident_db"] = {"value": time_e2e(lambda: db.query(qn, ratio=w["ratio"]), e2e_steps), "unit": UNIT,
                              "h2d_bytes_per_step": int(q_h.nbytes), "d2h_bytes_per_step": int(out_bytes),
                              "api": "slammatch.KeyframeDB.add(train) once, then .query(host queries) per step"}
        del db

    # ---- roofline of the dominant kernel, against ceilings measured now, in this process ----------------
    per_rank_cmp = cmp_per_step / (world if sharded else 1)
    kern_avg_ms = kern_ms / max(kern_n, 1)
    kernels_per_step = max(kern_n // (prof_steps + 1), 1)   # the profile also holds the ramp step
    cmp_per_launch = per_rank_cmp / kernels_per_step
    popc_tcmp, popc_lanes = ctx.probe_popc_peak(3)
    roof = {"kernel": kernel_name, "kernel_ms": kern_avg_ms, "traffic": None}
    if variant in ("tensor", "tensor4"):
        kind = "mxf4" if variant == "tensor4" else "f8f6f4"
        tf_probe, mac_probe = ctx.probe_tensor_peak(kind, 4096, 5)
        ach = cmp_per_launch * 512 / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
        bf16 = peaks.get("bf16_tflops")
        roof.update({"bound": "tensor", "achieved": ach, "peak": tf_probe, "unit": "TFLOP/s",
                     "frac": (ach / tf_probe) if ach else None,
                     "peak_note": ("tcgen05.mma kind::%s ceiling measured in this process right after the timed region "
                                   "(slm_probe_tensor_peak: back-to-back MMAs from shared memory on every SM, best of 5 launches; "
                                   "%.0f MAC/clk/SM by clock64); MEASURED_PEAKS.json has no fp8/fp4 entry (cuBLAS bf16 burst %s TF/s)"
                                   % (kind, mac_probe, bf16)),
                     "mac_per_clk_per_sm_probe": mac_probe,
                     "frac_vs_2x_bf16_measured": (ach / (2.0 * bf16)) if (ach and bf16) else None,
                     "algorithmic_unit": "1 cmp = 512 flop on the tensor pipe (256-term +-1 dot product) = 8 POPC32 on the integer pipe"})
    elif kernel_name == "knn2_stream_kernel":
        alg_bytes = 32.0 * (nq + nt / (world if sharded else 1)) + 16.0 * nq
        ach = alg_bytes / (kern_avg_ms * 1e-3) / 1e9 if kern_avg_ms > 0 else None
        peak = peaks.get("hbm_gbs", 6650.0)
        roof.update({"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                     "peak_note": ("measured copy bandwidth (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"),
                     "algorithmic_bytes": alg_bytes})
    elif variant == "bmma":
        ach = cmp_per_launch / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
        roof.update({"bound": "issue_slots(b1 mma.sync emulation)", "achieved": ach, "peak": None, "unit": "Tcmp/s", "frac": None,
                     "peak_note": "no native b1 tensor instruction on sm_100a; see profiles/r1_ncu_bmma_c5.txt"})
    else:
        ach = cmp_per_launch / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
        roof.update({"bound": "int_popc(xu pipe)", "achieved": ach, "peak": popc_tcmp, "unit": "Tcmp/s",
                     "frac": (ach / popc_tcmp) if ach else None,
                     "peak_note": "variant P's comparison loop on every SM, measured in this process (slm_probe_popc_peak)"})
    # the ALGORITHMIC integer roofline SURVEY.md section 8(d) names (8 POPC32 per comparison), whatever pipe the kernel used
    roof["popc_algorithmic_peak_tcmp_s"] = popc_tcmp
    roof["popc_lanes_per_clk_per_sm_probe"] = popc_lanes
    kernel_tcmp = cmp_per_launch / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
    roof["frac_vs_popc_algorithmic"] = (kernel_tcmp / popc_tcmp) if (kernel_tcmp and popc_tcmp) else None
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            ent = json.load(open(tr)).get(roof["kernel"], {}).get(wl_id)
            if isinstance(ent, dict):
                roof["traffic"], roof["traffic_source"] = ent.get("bytes"), ent.get("source")
        except Exception:
            pass

    res = {
        "workload": w["name"], "value": value, "unit": UNIT, "steps": steps, "ms_per_step": dev_ms / steps,
        "ms_min": float(np.min(per_step)), "ms_median": float(np.median(per_step)),
        "variant": variant, "nq": nq, "nt": nt,
        "l2": "inputs larger than L2 (streamed from HBM every step)" if flush is None else "256 MiB L2 flush between timed steps",
        "parallelism": (f"queries sliced over {world} ranks, train set replicated, {sm.last_exchange}" if by_q else
                        f"train rows sharded over {world} ranks, exchange of packed top-2 keys ({sm.last_exchange}) + merge"
                        if sharded else ("single GPU" if world == 1 else f"{world} replicas")),
        "scaling": "strong" if w["sharded"] else "replicas",
        "matched_per_step": matched, "matched_queries_per_s": matched * jobs / (dev_ms / steps * 1e-3),
        "wall_ms_per_step": wall_ms / steps, "gpu_launches": int(launches), "clocks": clocks,
        "e2e": e2e, "roofline": roof, "parity_check": parity,
    }
    # release this workload's device memory before the next one
    del out
    torch.cuda.empty_cache()
    return res


def ours(args):
    import torch
    import torch.distributed as dist
    from slammatch import _lib

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
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    env = dict(rank=rank, world=world, local=local, dev=dev, ctx=_lib.context(local), peaks=peaks)

    head = measure(args, args.workload, env, headline=True)
    # every other BASELINE config rides in the same line: all of them on one GPU, the sharded ones (c4) under torchrun
    if args.configs == "auto":
        others = [c for c in ("c4", "c3", "c2", "c1", "h1") if c != args.workload] if world == 1 else \
                 [c for c in ("c4", "c4q", "c5") if c != args.workload]
    elif args.configs == "none":
        others = []
    else:
        others = [c for c in args.configs.split(",") if c and c != args.workload]
    configs = {}
    failed = []
    for c in others:
        # a failure in one of the brief extra workloads must not cost the headline line: an extra that RAISED is recorded under
        # its name ("error") and listed in "configs_failed", and the run still ends with exit code 0 -- the headline was measured
        # and checked; an extra (or the headline) whose RESULT differs from the oracle ends the run with a non-zero exit code.
        # (Under torchrun the remaining extras are skipped: the ranks may no longer agree on the sequence of collectives.)
        try:
            r = measure(args, c, env, headline=False)
        except (Exception, SystemExit) as e:   # noqa: BLE001
            import traceback
            traceback.print_exc(file=sys.stderr)
            configs[c] = {"error": f"{type(e).__name__}: {e}"[:400], "parity_check": None}
            failed.append(c)
            if world > 1:
                break
            continue
        configs[c] = {k: r[k] for k in ("workload", "value", "unit", "steps", "ms_per_step", "ms_min", "ms_median", "variant",
                                        "parallelism", "scaling", "gpu_launches", "clocks", "parity_check")}
        configs[c]["kernel"] = r["roofline"]["kernel"]
        configs[c]["kernel_ms"] = r["roofline"]["kernel_ms"]
        configs[c]["roofline_frac"] = r["roofline"].get("frac")
        configs[c]["roofline_bound"] = r["roofline"].get("bound")
        configs[c]["frac_vs_popc_algorithmic"] = r["roofline"].get("frac_vs_popc_algorithmic")
        configs[c]["e2e"] = {k: v for k, v in r["e2e"].items() if k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step",
                                                                          "pageable", "matcher_knnMatch")}

    bad = [c for c, r in [(args.workload, head)] + list(configs.items()) if r["parity_check"] and not r["parity_check"]["ok"]]
    if rank == 0:
        w = WORKLOADS[args.workload]
        cpu = None
        if world == 1 and not args.no_cpu:
            cfg_id = int(args.workload[1]) + (5 if args.workload[0] == "h" else 0)
            run, kind, desc, cores, qs, ts = cpu_sample(w, cfg_id, target_s=12.0)
            t0 = time.perf_counter()
            run(qs, ts)
            dt = time.perf_counter() - t0
            cpu = {"value": qs.shape[0] * ts.shape[0] / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{qs.shape[0]} queries x first {ts.shape[0]} train rows, one pass ({dt:.1f} s)",
                   "matcher": desc, "cpu": cpu_model()}
        variant = head["variant"]
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": head["steps"],
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "ms_min": head["ms_min"],
            "ms_median": head["ms_median"], "higher_is_better": True,
            "scaling": head["scaling"], "vs_baseline": None,
            "dtype": {"tensor4": "e2m1(+-1 bits, UE8M0 scales 1.0), f32 accumulate (exact)",
                      "tensor": "e4m3(+-1 bits), f32 accumulate (exact)"}.get(variant, "u32"),
            "data": "synthetic",
            "config": {"workload": head["workload"], "nq": head["nq"], "nt": head["nt"], "variant": variant,
                       "variant_requested": args.variant, "l2": head["l2"], "parallelism": head["parallelism"]},
            "matched_queries_per_s": head["matched_queries_per_s"], "matched_per_step": head["matched_per_step"],
            "wall_ms_per_step": head["wall_ms_per_step"],
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
            "roofline": head["roofline"], "cpu_baseline": cpu, "parity_check": head["parity_check"],
            "configs": configs, "configs_failed": failed,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        raise SystemExit(f"FAILED for {bad}: the GPU result differs from the oracle (see parity_check of that entry)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="auto", choices=["auto", "popc", "tensor", "tensor4", "bmma"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--configs", default="auto",
                    help="other BASELINE configs measured briefly after the headline workload and reported under 'configs': "
                         "auto (1 GPU: c4,c3,c2,c1,h1; torchrun: the sharded c4), none, or a comma list")
    ap.add_argument("--config-steps", type=int, default=10, help="timed steps of each entry under 'configs'")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "nvlink", "a2a"],
                    help="sharded path: how per-rank keys are exchanged (auto = NVLink peer stores when available)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check of the timed path's output")
    args = ap.parse_args()
    if args.impl == "reference":
        w = WORKLOADS[args.workload]
        cfg_id = int(args.workload[1]) + (5 if args.workload[0] == "h" else 0)
        reference_arm(args, w, cfg_id)
    else:
        ours(args)


if __name__ == "__main__":
    main()
