"""Calibration of SLM_VARIANT_AUTO's frame-kernel threshold (slm_ctx::frame_max_clk): device time per call of the
single-launch frame kernel against the general path (tensor / popc + refine/merge + finalize) over frame-sized
shapes.  Run on the B200 box:  python scripts/calib_frame.py > gpurun_out/calib_frame.txt"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "slam-1_b200")]

import torch  # noqa: E402
from slammatch import _lib, synth  # noqa: E402


def make_ctx(max_clk, warps=8):
    os.environ["SLM_FRAME_MAX_CLK"] = str(max_clk)
    os.environ["SLM_FRAME_WARPS"] = str(warps)
    ctx = _lib.Context(0)
    del os.environ["SLM_FRAME_MAX_CLK"], os.environ["SLM_FRAME_WARPS"]
    return ctx


SPLIT_CHOICES = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18, 20, 24, 28, 32, 40, 48, 56, 64)


def model_clk(nq, nt, cross, sms=148):
    """frame_plan() of csrc/knn2_frame.cu, restated: (estimated clocks, KQ, slices) of the chosen plan."""
    best = None
    for kq in (4, 2, 1):
        qc = 32 * kq
        for s0 in SPLIT_CHOICES:
            if s0 > 1 and -(-nt // s0) < 16:
                break
            rps0 = -(-nt // s0)
            ctas, worst, s1 = -(-nq // qc) * s0, rps0, 1
            if cross:
                s1 = min(64, max(1, -(-nq // rps0)))
                ctas += -(-nt // qc) * s1
                worst = max(worst, -(-nq // s1))
            tiles = -(-worst // 512)
            est = (qc * worst // 2 + 1200 + 700 * (tiles - 1) + 20 * max(s0, s1)) * -(-ctas // sms)
            if best is None or est < best[0]:
                best = (est, kq, s0)
    return best


def time_call(ctx, qd, td, cross, iters=40):
    nq, nt = qd.shape[0], td.shape[0]
    idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    dist = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    acc = torch.empty((nq,), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def call():
        _lib.check(ctx.lib.slm_knn2_filter(ctx.handle, qd.data_ptr(), nq, td.data_ptr(), nt, 0, 7, 10, int(cross),
                                           idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), stream))
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        call()
    b.record()
    torch.cuda.synchronize()
    ctx.profile(True)
    ctx.profile_read()
    for _ in range(10):
        call()
    torch.cuda.synchronize()
    kms, kn = ctx.profile_read()
    ctx.profile(False)
    kern_us = kms / max(kn, 1) * 1e3
    t0 = time.perf_counter()
    for _ in range(20):          # latency of one synchronous call (what a host caller waits for)
        call()
        torch.cuda.synchronize()
    lat = (time.perf_counter() - t0) / 20 * 1e6
    return a.elapsed_time(b) / iters * 1e3, ctx.last_kernel(), (idx.clone(), dist.clone(), acc.clone()), lat, kern_us


def main():
    frames = {w: make_ctx(1 << 60, w) for w in (4, 8, 16)}
    general = make_ctx(0)
    shapes = [(100, 100), (500, 500), (1000, 1000), (1500, 1500), (2000, 2000), (1000, 4000), (3000, 3000),
              (4000, 4000), (2000, 8000), (500, 20000), (2000, 20000), (6000, 6000), (64, 100000)]
    print(f"{'nq':>6} {'nt':>7} {'cross':>5} {'Mcmp':>8} {'f4 us':>7} {'f8 us':>7} {'f16 us':>7} {'model clk (kq,S0)':>18} {'general us':>10}  general kernel   same")
    for nq, nt in shapes:
        q, t = synth.planted(nq, nt, nq + nt)
        qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
        for cross in (False, True):
            g_us, gk, go, g_lat, _ = time_call(general, qd, td, cross)
            f_us, same = {}, True
            for w, fctx in frames.items():
                f_us[w], _, fo, f_lat, f_k = time_call(fctx, qd, td, cross)
                if w == 8:
                    f8_lat, f8_k = f_lat, f_k
                same = same and all(torch.equal(x, y) for x, y in zip(fo, go))
            est, kq, S = model_clk(nq, nt, cross)
            print(f"{nq:6d} {nt:7d} {int(cross):5d} {nq * nt * (2 if cross else 1) / 1e6:8.1f} {f_us[4]:7.1f} {f_us[8]:7.1f} {f_us[16]:7.1f} "
                  f"{est:10d} ({kq},{S:2d}) {g_us:10.1f}  {gk:16s} {same}  sync latency f8 {f8_lat:6.1f} general {g_lat:6.1f}  f8 kernel {f8_k:6.1f} us")


if __name__ == "__main__":
    main()
