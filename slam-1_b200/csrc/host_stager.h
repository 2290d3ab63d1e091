// host_stager.h -- a few host threads copy a long PAGEABLE array, chunk by chunk, into a ring of pinned chunks; the
// consumer (slm_knn2_host, api.cu) sends every finished chunk over PCIe and frees the ring slot when that copy is done.
// Plain C++ (no CUDA): tests/test_host_stager.py compiles this header with g++ and runs it against a simulated consumer.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <vector>

static constexpr int kStageSlots = 4;

struct HostStager {
    const uint8_t *src = nullptr;
    uint8_t *ring = nullptr;
    size_t total = 0, chunk = 0;
    int64_t n_chunks = 0;
    int n_threads = 0;
    std::atomic<int64_t> released{0};                 // chunks [0, released + kStageSlots) may be written into the ring
    std::atomic<int> *arrived = nullptr;              // per chunk: threads that have copied their piece
    std::atomic<bool> abort{false};
    std::vector<std::thread> threads;

    int start(const uint8_t *src_, size_t total_, size_t chunk_, uint8_t *ring_, int n_threads_)
    {
        src = src_, total = total_, chunk = chunk_, ring = ring_, n_threads = n_threads_;
        n_chunks = (int64_t)((total + chunk - 1) / chunk);
        arrived = new (std::nothrow) std::atomic<int>[n_chunks];
        if (!arrived) return -1;
        for (int64_t c = 0; c < n_chunks; ++c) arrived[c].store(0, std::memory_order_relaxed);
        try {
            for (int t = 0; t < n_threads; ++t) threads.emplace_back([this, t] { work(t); });
        } catch (...) {
            stop();
            return -1;
        }
        return 0;
    }
    void work(int t)
    {
        for (int64_t c = 0; c < n_chunks; ++c) {
            while (c >= released.load(std::memory_order_acquire) + kStageSlots) {
                if (abort.load(std::memory_order_relaxed)) return;
                std::this_thread::yield();
            }
            const size_t c0 = (size_t)c * chunk, len = std::min(chunk, total - c0);
            // 64-byte aligned pieces, one per thread
            const size_t piece = ((len + n_threads - 1) / n_threads + 63) & ~(size_t)63;
            const size_t a = std::min(len, piece * t), b = std::min(len, piece * (t + 1));
            if (b > a) memcpy(ring + (size_t)(c % kStageSlots) * chunk + a, src + c0 + a, b - a);
            arrived[c].fetch_add(1, std::memory_order_release);
        }
    }
    // chunk c is complete in its ring slot (returns its address)
    const uint8_t *wait(int64_t c)
    {
        while (arrived[c].load(std::memory_order_acquire) < (int)threads.size()) std::this_thread::yield();
        return ring + (size_t)(c % kStageSlots) * chunk;
    }
    void release(int64_t n) { released.store(n, std::memory_order_release); }
    void stop()
    {
        abort.store(true, std::memory_order_relaxed);
        for (std::thread &th : threads)
            if (th.joinable()) th.join();
        threads.clear();
        delete[] arrived;
        arrived = nullptr;
    }
    ~HostStager() { stop(); }
};
