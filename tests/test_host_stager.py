"""The host-side staging ring of slm_knn2_host (slam-1_b200/csrc/host_stager.h) against a simulated consumer, on the CPU:
worker threads may only overwrite a ring slot after the consumer has released it, every chunk arrives complete, ragged
sizes and early aborts do not hang.  Compiled with g++ (ThreadSanitizer when the toolchain has it)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

HARNESS = r"""
#include "host_stager.h"
#include <cstdio>
#include <cstdlib>
#include <deque>
int run(size_t total, size_t chunk, int threads, bool abort_early)
{
    std::vector<uint8_t> src(total), dst(total, 0), ring((size_t)kStageSlots * chunk);
    for (size_t i = 0; i < total; ++i) src[i] = (uint8_t)((i * 2654435761u) >> 13);
    HostStager st;
    if (st.start(src.data(), total, chunk, ring.data(), threads) != 0) return 2;
    const int64_t n = (int64_t)((total + chunk - 1) / chunk);
    // the "H2D copy" of a chunk is executed as LATE as the protocol allows: only when its slot is about to be released
    // (cudaEventSynchronize in api.cu) or at the end -- a worker that overwrites a slot too early corrupts dst
    std::deque<int64_t> pending;
    auto do_copy = [&](int64_t c) {
        const size_t c0 = (size_t)c * chunk, len = std::min(chunk, total - c0);
        memcpy(dst.data() + c0, ring.data() + (size_t)(c % kStageSlots) * chunk, len);
    };
    for (int64_t c = 0; c < n; ++c) {
        if (abort_early && c == n / 2) return 0;              // destructor must stop and join the workers
        if (c >= kStageSlots) {
            do_copy(pending.front());
            pending.pop_front();
            st.release(c - kStageSlots + 1);
        }
        st.wait(c);
        pending.push_back(c);
    }
    while (!pending.empty()) { do_copy(pending.front()); pending.pop_front(); }
    return memcmp(src.data(), dst.data(), total) == 0 ? 0 : 1;
}
int main()
{
    const size_t sizes[][2] = {{1, 64}, {63, 64}, {64, 64}, {1000003, 4096}, {(size_t)8 << 20, (size_t)1 << 20},
                               {((size_t)9 << 20) + 17, (size_t)1 << 20}, {5000, 100}};
    for (auto &s : sizes)
        for (int t : {1, 2, 3, 8})
            for (int ab = 0; ab < 2; ++ab) {
                int r = run(s[0], s[1], t, ab != 0);
                if (r) { printf("FAIL total=%zu chunk=%zu threads=%d abort=%d -> %d\n", s[0], s[1], t, ab, r); return 1; }
            }
    printf("ok\n");
    return 0;
}
"""


@pytest.mark.parametrize("tsan", [False, True])
def test_host_stager_ring_protocol(tmp_path, tsan):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not found")
    src = tmp_path / "harness.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / ("harness_tsan" if tsan else "harness")
    cmd = [cxx, "-O1" if tsan else "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "slam-1_b200", "csrc"), str(src),
           "-o", str(exe)]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if tsan:
            pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
        raise AssertionError(r.stderr)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    if tsan and r.returncode != 0 and "FATAL: ThreadSanitizer" in r.stderr and "WARNING: ThreadSanitizer" not in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container: " + r.stderr[-200:])
    assert r.returncode == 0 and "ok" in r.stdout and "WARNING: ThreadSanitizer" not in r.stderr, r.stdout + r.stderr[-3000:]
