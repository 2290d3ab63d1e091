// api.cu -- the extern "C" surface declared in include/slammatch.h.
//
// Host-side orchestration only: argument checks, workspace, variant dispatch, stream ordering.
// There is no CPU implementation behind any entry point: without a CUDA device they fail.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <new>
#include <numeric>
#include <thread>
#include <vector>

#include "slm_internal.cuh"

// ---- error plumbing ------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int slm_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int slm_buf_reserve(slm_ctx *ctx, slm_buf *buf, size_t bytes)
{
    if (bytes <= buf->bytes) return SLM_OK;
    const size_t want = bytes + bytes / 4 + 4096;  // slack so that slowly growing inputs do not realloc each call
    // Stream-ordered growth: the old block is released after everything already queued on the call's stream (slm_enter
    // has ordered that stream behind the previous call's), the new one is usable by everything queued after this point.
    // No device-wide synchronisation, no host stall in the middle of a stream of calls.
    cudaStream_t s = ctx->cur_stream;
    if (buf->p) {
        SLM_CUDA(cudaFreeAsync(buf->p, s));
        buf->p = nullptr;
        buf->bytes = 0;
    }
    cudaError_t e = cudaMallocAsync(&buf->p, want, s);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        buf->p = nullptr;
        return slm_fail(SLM_ERR_NOMEM, "cudaMallocAsync(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    buf->bytes = want;
    return SLM_OK;
}

int slm_enter(slm_ctx *ctx, cudaStream_t stream)
{
    // The workspace of a ctx is shared by all its calls.  A call on another stream than the previous one first waits,
    // on the device, for everything queued on that stream so far (which includes the previous call's kernels).
    if (ctx->have_last && ctx->last_stream != stream) {
        cudaError_t e = cudaEventRecord(ctx->last_ev, ctx->last_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, ctx->last_ev, 0);
        if (e != cudaSuccess) {                     // e.g. the caller destroyed the previous stream
            (void)cudaGetLastError();
            SLM_CUDA(cudaDeviceSynchronize());
        }
    }
    ctx->cur_stream = stream;
    ctx->last_stream = stream;
    ctx->have_last = true;
    return SLM_OK;
}

int slm_leave(slm_ctx *ctx, cudaStream_t stream)
{
    (void)stream;
    (void)ctx;
    return SLM_OK;
}

static int pin_reserve(slm_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->pin_bytes) return SLM_OK;
    if (ctx->pin) {
        SLM_CUDA(cudaDeviceSynchronize());
        SLM_CUDA(cudaFreeHost(ctx->pin));
        ctx->pin = nullptr;
        ctx->pin_bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&ctx->pin, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        ctx->pin = nullptr;
        return slm_fail(SLM_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    ctx->pin_bytes = want;
    return SLM_OK;
}

// ---- pageable host inputs --------------------------------------------------------------------------------------------
// cudaMemcpyAsync from PAGEABLE memory is staged by the driver on the calling thread at ~11 GB/s (measured: config 5's
// 320 MB train set, 29 ms per call against 6 ms from pinned memory) -- and the drop-in caller hands exactly such arrays
// (numpy from ORB, orb.py:23-24).  So a long pageable train set is copied by a few host threads into a ring of pinned
// chunks, each chunk going out over PCIe (and being searched) while the next ones are filled.
#include "host_stager.h"

static bool host_is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

static int stage_reserve(slm_ctx *ctx, size_t bytes)
{
    if (bytes > ctx->stage_bytes) {
        if (ctx->stage_pin) {
            SLM_CUDA(cudaDeviceSynchronize());
            SLM_CUDA(cudaFreeHost(ctx->stage_pin));
            ctx->stage_pin = nullptr;
            ctx->stage_bytes = 0;
        }
        cudaError_t e = cudaMallocHost(&ctx->stage_pin, bytes);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            ctx->stage_pin = nullptr;
            return slm_fail(SLM_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        }
        ctx->stage_bytes = bytes;
    }
    for (cudaEvent_t &ev : ctx->stage_ev) {
        if (!ev) SLM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        else SLM_CUDA(cudaEventSynchronize(ev));      // an earlier upload may still be reading the ring
    }
    return SLM_OK;
}

// One host array -> device on `stream`.  Pinned (or small) sources are handed to the driver as they are; a long pageable one
// goes through the pinned ring in 8 MB pieces filled by the host threads.
static int upload_host(slm_ctx *ctx, void *dst, const uint8_t *src, size_t bytes, cudaStream_t stream)
{
    const size_t kPiece = (size_t)8 << 20;
    if (bytes == 0) return SLM_OK;
    if (ctx->host_threads <= 0 || bytes < 2 * kPiece || !host_is_pageable(src)) {
        SLM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return SLM_OK;
    }
    SLM_TRY(stage_reserve(ctx, (size_t)kStageSlots * kPiece));
    HostStager stager;
    if (stager.start(src, bytes, kPiece, reinterpret_cast<uint8_t *>(ctx->stage_pin), ctx->host_threads) != 0)
        return slm_fail(SLM_ERR_NOMEM, "could not start the host staging threads");
    for (int64_t c = 0; c < stager.n_chunks; ++c) {
        if (c >= kStageSlots) {
            SLM_CUDA(cudaEventSynchronize(ctx->stage_ev[c % kStageSlots]));
            stager.release(c - kStageSlots + 1);
        }
        const uint8_t *piece = stager.wait(c);
        const size_t c0 = (size_t)c * kPiece, len = std::min(kPiece, bytes - c0);
        SLM_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(dst) + c0, piece, len, cudaMemcpyHostToDevice, stream));
        SLM_CUDA(cudaEventRecord(ctx->stage_ev[c % kStageSlots], stream));
    }
    return SLM_OK;
}

int slm_prof_mark(slm_ctx *ctx, cudaStream_t stream, int tag)
{
    if (!ctx->profile) return SLM_OK;
    if (ctx->prof_n >= slm_ctx::kMaxProf) {       // full: count what is lost, slm_profile_read reports it
        ctx->prof_dropped += 1;
        return SLM_OK;
    }
    if (!ctx->prof_ev) {
        ctx->prof_ev = new (std::nothrow) cudaEvent_t[slm_ctx::kMaxProf];
        ctx->prof_tag = new (std::nothrow) unsigned char[slm_ctx::kMaxProf];
        if (!ctx->prof_ev || !ctx->prof_tag) return slm_fail(SLM_ERR_NOMEM, "out of host memory");
        for (int i = 0; i < slm_ctx::kMaxProf; ++i) SLM_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
    }
    SLM_CUDA(cudaEventRecord(ctx->prof_ev[ctx->prof_n], stream));
    ctx->prof_tag[ctx->prof_n] = (unsigned char)tag;
    ctx->prof_n += 1;
    return SLM_OK;
}

static int check_ctx(slm_ctx *ctx)
{
    if (!ctx) return slm_fail(SLM_ERR_INVALID, "ctx is NULL");
    SLM_CUDA(cudaSetDevice(ctx->device));
    return SLM_OK;
}

static int check_sizes(int64_t nq, int64_t nt, int64_t base)
{
    if (nq < 0 || nt < 0 || base < 0) return slm_fail(SLM_ERR_INVALID, "negative size (nq=%lld nt=%lld base=%lld)",
                                                       (long long)nq, (long long)nt, (long long)base);
    if (nq > 0x7FFFFFFFll || base + nt > 0x7FFFFFFFll)
        return slm_fail(SLM_ERR_UNSUPPORTED, "indices must fit int32 (nq=%lld base+nt=%lld)", (long long)nq,
                        (long long)(base + nt));
    return SLM_OK;
}

__global__ void fill_keys_none_kernel(unsigned long long *keys, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = kKeyNone;
}

// Variant dispatch: produce packed keys uint64[nq][2] for one problem.
static int knn2_keys_dispatch(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt,
                              int64_t base, uint64_t *keys_out, cudaStream_t stream)
{
    if (nq == 0) return SLM_OK;
    if (nt == 0) {
        fill_keys_none_kernel<<<(unsigned)((2 * nq + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<unsigned long long *>(keys_out), 2 * nq);
        SLM_CUDA(cudaGetLastError());
        ctx->launches += 1;
        return SLM_OK;
    }
    if (((uintptr_t)q & 15) || ((uintptr_t)t & 15))
        return slm_fail(SLM_ERR_INVALID, "descriptor pointers must be 16-byte aligned");
    switch (ctx->variant) {
    case SLM_VARIANT_TENSOR:
        return slm_tc_knn2_keys(ctx, q, nq, t, nt, base, keys_out, stream, false);
    case SLM_VARIANT_TENSOR4:
        return slm_tc_knn2_keys(ctx, q, nq, t, nt, base, keys_out, stream, true);
    case SLM_VARIANT_BMMA:
        return slm_bmma_knn2_keys(ctx, q, nq, t, nt, base, keys_out, stream);
    case SLM_VARIANT_POPC:
        return slm_popc_knn2_keys(ctx, q, nq, t, nt, base, keys_out, stream);
    default:
        return slm_auto_knn2_keys(ctx, q, nq, t, nt, base, keys_out, stream);
    }
}

// ---- public ABI ----------------------------------------------------------------------------------
extern "C" {

const char *slm_last_error(void) { return g_err; }
int slm_version(void) { return SLM_VERSION; }

int slm_create(int device, slm_ctx **ctx_out)
{
    if (!ctx_out) return slm_fail(SLM_ERR_INVALID, "ctx_out is NULL");
    *ctx_out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return slm_fail(SLM_ERR_CUDA, "no CUDA device available (%s); libslammatch has no CPU path",
                        e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return slm_fail(SLM_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    SLM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SLM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return slm_fail(SLM_ERR_CUDA, "device %d is sm_%d%d; libslammatch is built for sm_100a (B200) only", device,
                        prop.major, prop.minor);
    slm_ctx *ctx = new (std::nothrow) slm_ctx();
    if (!ctx) return slm_fail(SLM_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->force_1cta = getenv("SLM_TC_1CTA") != nullptr;
    ctx->trace = getenv("SLM_TRACE") != nullptr;
    if (const char *e = getenv("SLM_TC_MAX_CPG")) {
        int v = atoi(e);
        if (v >= 1) ctx->max_cpg = v;
    }
    ctx->no_frame_refine = getenv("SLM_TC_NO_FRAME_REFINE") != nullptr;
    ctx->tc_plan_mt = getenv("SLM_TC_PLAN_MT") != nullptr;
    if (const char *e = getenv("SLM_TC_FP4")) ctx->tc_fp4 = atoi(e) != 0;
    if (const char *e = getenv("SLM_TC4_CHUNK")) {
        int v = atoi(e);
        if (v == 120 || v == 40) ctx->tc4_chunk = v;
    }
    ctx->no_pdl = getenv("SLM_NO_PDL") != nullptr;
    {
        const unsigned hw = std::thread::hardware_concurrency();
        ctx->host_threads = hw >= 16 ? 8 : hw >= 8 ? 4 : hw >= 4 ? 2 : 0;
        if (const char *e = getenv("SLM_HOST_STAGE_THREADS")) {
            int v = atoi(e);
            if (v >= 0 && v <= 32) ctx->host_threads = v;
        }
    }
    ctx->tc4_timing = getenv("SLM_TC4_TIMING") != nullptr;
    ctx->exchange_max_blocks = 2ll * prop.multiProcessorCount;
    if (const char *e = getenv("SLM_EXCHANGE_MAX_BLOCKS")) {
        long long v = atoll(e);
        if (v >= 1) ctx->exchange_max_blocks = v;
    }
    if (const char *e = getenv("SLM_EXCHANGE_TWO_PHASE_MIN")) ctx->exchange_two_phase_min = atoll(e);
    if (const char *e = getenv("SLM_EXCHANGE_TWO_PHASE_WORLD")) ctx->exchange_two_phase_world = atoi(e);
    if (const char *e = getenv("SLM_EXCHANGE_MAX_POLLS")) {
        long long v = atoll(e);
        if (v >= 1 && v <= 0xFFFFFFFFll) ctx->exchange_max_polls = (unsigned)v;
    }
    ctx->exchange_wide_keys = getenv("SLM_EXCHANGE_WIDE_KEYS") != nullptr;
    if (const char *e = getenv("SLM_TC_CHAIN")) ctx->tc_chain_max = atoi(e);
    if (const char *e = getenv("SLM_TC_CHAIN_MIN")) ctx->tc_chain_min = atoi(e);
    if (const char *e = getenv("SLM_FRAME_MAX_CLK")) ctx->frame_max_clk = atoll(e);
    if (const char *e = getenv("SLM_FRAME_WARPS")) {
        int v = atoi(e);
        if (v == 4 || v == 8 || v == 16) ctx->frame_warps = v;
    }
    if (const char *e = getenv("SLM_TC_EPOCH_TILES")) {
        int v = atoi(e);
        if (v >= 8 && v <= 4096) ctx->epoch_tiles = v;
    }   // A/B switch: single-CTA tcgen05 kernel only
    SLM_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    SLM_CUDA(cudaEventCreateWithFlags(&ctx->ev[0], cudaEventDisableTiming));
    SLM_CUDA(cudaEventCreateWithFlags(&ctx->ev[1], cudaEventDisableTiming));
    SLM_CUDA(cudaEventCreateWithFlags(&ctx->last_ev, cudaEventDisableTiming));
    // last-block-done counter of the frame kernel and the exchange producers: zeroed HERE, synchronously -- a lazy
    // cudaMemset on the legacy stream is not ordered with kernels on the caller's non-blocking streams
    SLM_CUDA(cudaMalloc(&ctx->done_counter, sizeof(unsigned)));
    SLM_CUDA(cudaMemset(ctx->done_counter, 0, sizeof(unsigned)));
    SLM_CUDA(cudaDeviceSynchronize());
    *ctx_out = ctx;
    return SLM_OK;
}

int slm_destroy(slm_ctx *ctx)
{
    if (!ctx) return SLM_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    slm_buf *bufs[] = {&ctx->scratch, &ctx->keys, &ctx->rev, &ctx->misc, &ctx->io, &ctx->tickets, &ctx->chi2_leaves};
    for (slm_buf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->stage_pin) cudaFreeHost(ctx->stage_pin);
    for (cudaEvent_t ev : ctx->stage_ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->done_counter) cudaFree(ctx->done_counter);
    if (ctx->exchange_status) cudaFreeHost(ctx->exchange_status);
    if (ctx->last_ev) cudaEventDestroy(ctx->last_ev);
    if (ctx->prof_ev) {
        for (int i = 0; i < slm_ctx::kMaxProf; ++i) cudaEventDestroy(ctx->prof_ev[i]);
        delete[] ctx->prof_ev;
        delete[] ctx->prof_tag;
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (cudaEvent_t ev : ctx->chunk_ev)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    delete ctx;
    (void)cudaGetLastError();
    return SLM_OK;
}

int slm_set_variant(slm_ctx *ctx, int variant)
{
    if (!ctx) return slm_fail(SLM_ERR_INVALID, "ctx is NULL");
    if (variant < SLM_VARIANT_AUTO || variant > SLM_VARIANT_TENSOR4)
        return slm_fail(SLM_ERR_INVALID, "unknown variant %d", variant);
    ctx->variant = variant;
    return SLM_OK;
}

int64_t slm_launch_count(const slm_ctx *ctx) { return ctx ? ctx->launches : 0; }
int slm_last_variant(const slm_ctx *ctx) { return ctx ? ctx->last_variant : 0; }
const char *slm_last_kernel(const slm_ctx *ctx) { return ctx ? ctx->last_kernel : ""; }

int slm_profile_enable(slm_ctx *ctx, int enable)
{
    if (!ctx) return slm_fail(SLM_ERR_INVALID, "ctx is NULL");
    ctx->profile = enable ? 1 : 0;
    return SLM_OK;
}

int slm_profile_read(slm_ctx *ctx, double *kernel_ms_out, int64_t *launches_out)
{
    SLM_TRY(check_ctx(ctx));
    double total = 0.0;
    int64_t launches = 0;
    static const char *names[] = {"call_begin", "main_begin", "main_end", "call_end"};
    for (int i = 0; i + 1 < ctx->prof_n || i < ctx->prof_n; ++i) {
        SLM_CUDA(cudaEventSynchronize(ctx->prof_ev[i]));
        if (i + 1 >= ctx->prof_n) break;
        float ms = 0.f;
        SLM_CUDA(cudaEventSynchronize(ctx->prof_ev[i + 1]));
        SLM_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[i], ctx->prof_ev[i + 1]));
        if (ctx->prof_tag[i] == SLM_TAG_MAIN_BEGIN && ctx->prof_tag[i + 1] == SLM_TAG_MAIN_END) {
            total += ms;
            launches += 1;
        }
        if (ctx->trace)
            fprintf(stderr, "slm trace %4d  %-10s -> %-10s %9.3f us\n", i, names[ctx->prof_tag[i] & 3],
                    names[ctx->prof_tag[i + 1] & 3], ms * 1e3);
    }
    if (kernel_ms_out) *kernel_ms_out = total;
    if (launches_out) *launches_out = launches;
    ctx->prof_n = 0;
    if (ctx->prof_dropped > 0) {
        const long long d = ctx->prof_dropped;
        ctx->prof_dropped = 0;
        return slm_fail(SLM_ERR_UNSUPPORTED, "profile buffer overflowed: %lld event marks were dropped (read at least every %d marks)",
                        d, slm_ctx::kMaxProf);
    }
    return SLM_OK;
}

int slm_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                  uint64_t *keys_out, void *stream)
{
    SLM_TRY(check_ctx(ctx));
    SLM_TRY(check_sizes(nq, nt, base));
    if (nq == 0) return SLM_OK;
    if (!q || !keys_out || (nt > 0 && !t)) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    SLM_TRY(slm_enter(ctx, (cudaStream_t)stream));
    SLM_TRY(slm_prof_mark(ctx, (cudaStream_t)stream, SLM_TAG_CALL_BEGIN));
    SLM_TRY(knn2_keys_dispatch(ctx, q, nq, t, nt, base, keys_out, (cudaStream_t)stream));
    return slm_prof_mark(ctx, (cudaStream_t)stream, SLM_TAG_CALL_END);
}

int slm_knn2_filter(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                    int32_t ratio_num, int32_t ratio_den, int32_t cross_check, int32_t *idx_out,
                    int32_t *dist_out, uint8_t *accept_out, void *stream_)
{
    SLM_TRY(check_ctx(ctx));
    SLM_TRY(check_sizes(nq, nt, base));
    if (ratio_num > 0 && ratio_den <= 0) return slm_fail(SLM_ERR_INVALID, "ratio_den must be > 0");
    if (nq == 0) return SLM_OK;
    if (!q || (nt > 0 && !t)) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    SLM_TRY(slm_enter(ctx, stream));
    const bool cross = cross_check && accept_out && nt > 0;
    if (ctx->variant == SLM_VARIANT_AUTO && nt > 0 && slm_frame_eligible(ctx, nq, nt, cross)) {
        // frame-to-frame shapes: search, merge, ratio and cross-check in one launch
        if (((uintptr_t)q & 15) || ((uintptr_t)t & 15))
            return slm_fail(SLM_ERR_INVALID, "descriptor pointers must be 16-byte aligned");
        SLM_TRY(slm_prof_mark(ctx, stream, SLM_TAG_CALL_BEGIN));
        SLM_TRY(slm_frame_knn2(ctx, q, nq, t, nt, base, ratio_num, ratio_den, cross ? 1 : 0, nullptr, idx_out, dist_out,
                               accept_out, stream));
        return slm_prof_mark(ctx, stream, SLM_TAG_CALL_END);
    }
    SLM_TRY(slm_buf_reserve(ctx, &ctx->keys, (size_t)nq * 16));
    uint64_t *keys = reinterpret_cast<uint64_t *>(ctx->keys.p);
    SLM_TRY(slm_prof_mark(ctx, stream, SLM_TAG_CALL_BEGIN));
    SLM_TRY(knn2_keys_dispatch(ctx, q, nq, t, nt, base, keys, stream));
    const uint64_t *rev = nullptr;
    int rev_by_query = 0;
    if (cross && nq < nt) {
        // Reverse search, reduced: query i can only be mutual with its own best match, so only the <= nq train rows
        // best(i) are searched against the queries (nq x nq comparisons instead of nt x nq; lowest query index wins
        // ties either way).  Frame-sized nq: one launch that also applies the ratio test and writes the outputs.
        const int fwd_variant = ctx->last_variant;
        const char *fwd_kernel = ctx->last_kernel;
        if (ctx->variant == SLM_VARIANT_AUTO && slm_frame_eligible(ctx, nq, nq, false)) {
            SLM_TRY(slm_frame_revcheck(ctx, q, nq, t, base, ratio_num, ratio_den, keys, idx_out, dist_out, accept_out, stream));
            ctx->last_variant = fwd_variant;
            ctx->last_kernel = fwd_kernel;
            return slm_prof_mark(ctx, stream, SLM_TAG_CALL_END);
        }
        SLM_TRY(slm_buf_reserve(ctx, &ctx->misc, (size_t)nq * 32));
        SLM_TRY(slm_buf_reserve(ctx, &ctx->rev, (size_t)nq * 16));
        uint32_t *best_rows = reinterpret_cast<uint32_t *>(ctx->misc.p);
        SLM_TRY(slm_gather_best_rows(ctx, keys, nq, base, t, best_rows, stream));
        SLM_TRY(knn2_keys_dispatch(ctx, best_rows, nq, q, nq, 0, reinterpret_cast<uint64_t *>(ctx->rev.p), stream));
        rev = reinterpret_cast<const uint64_t *>(ctx->rev.p);
        rev_by_query = 1;
    } else if (cross) {
        // reverse search: every train row against all queries (lowest query index wins ties)
        SLM_TRY(slm_buf_reserve(ctx, &ctx->rev, (size_t)nt * 16));
        SLM_TRY(knn2_keys_dispatch(ctx, t, nt, q, nq, 0, reinterpret_cast<uint64_t *>(ctx->rev.p), stream));
        rev = reinterpret_cast<const uint64_t *>(ctx->rev.p);
    }
    SLM_TRY(slm_finalize(ctx, keys, nq, ratio_num, ratio_den, rev, nt, base, idx_out, dist_out, accept_out, stream,
                         rev_by_query));
    return slm_prof_mark(ctx, stream, SLM_TAG_CALL_END);
}

int slm_knn2_masked(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                    const uint8_t *mask, int64_t mask_row_stride, int32_t ratio_num, int32_t ratio_den, int32_t *idx_out,
                    int32_t *dist_out, uint8_t *accept_out, void *stream_)
{
    SLM_TRY(check_ctx(ctx));
    SLM_TRY(check_sizes(nq, nt, base));
    if (ratio_num > 0 && ratio_den <= 0) return slm_fail(SLM_ERR_INVALID, "ratio_den must be > 0");
    if (nq == 0) return SLM_OK;
    if (!q || (nt > 0 && (!t || !mask))) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    if (mask_row_stride < nt) return slm_fail(SLM_ERR_INVALID, "mask_row_stride (%lld) < nt (%lld)", (long long)mask_row_stride,
                                              (long long)nt);
    if (((uintptr_t)q & 15) || ((uintptr_t)t & 15))
        return slm_fail(SLM_ERR_INVALID, "descriptor pointers must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    SLM_TRY(slm_enter(ctx, stream));
    SLM_TRY(slm_prof_mark(ctx, stream, SLM_TAG_CALL_BEGIN));
    if (nt == 0) {
        // empty train set: every row is empty, exactly like the unmasked call
        SLM_TRY(slm_buf_reserve(ctx, &ctx->keys, (size_t)nq * 16));
        uint64_t *keys = reinterpret_cast<uint64_t *>(ctx->keys.p);
        SLM_TRY(knn2_keys_dispatch(ctx, q, nq, t, nt, base, keys, stream));
        SLM_TRY(slm_finalize(ctx, keys, nq, ratio_num, ratio_den, nullptr, nt, base, idx_out, dist_out, accept_out, stream));
    } else {
        SLM_TRY(slm_masked_knn2(ctx, q, nq, t, nt, base, mask, mask_row_stride, ratio_num, ratio_den, idx_out, dist_out,
                                accept_out, stream));
    }
    return slm_prof_mark(ctx, stream, SLM_TAG_CALL_END);
}

int slm_knn2(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
             int32_t *idx_out, int32_t *dist_out, void *stream)
{
    return slm_knn2_filter(ctx, q, nq, t, nt, base, 0, 1, 0, idx_out, dist_out, nullptr, stream);
}

int slm_knn2_batched(slm_ctx *ctx, const uint32_t *desc, int64_t n_frames, int64_t n_per_frame,
                     const int32_t *pairs_host, int64_t n_pairs, int32_t ratio_num, int32_t ratio_den,
                     int32_t *idx_out, int32_t *dist_out, uint8_t *accept_out, void *stream_)
{
    SLM_TRY(check_ctx(ctx));
    if (n_frames < 0 || n_per_frame < 0 || n_pairs < 0) return slm_fail(SLM_ERR_INVALID, "negative size");
    if (ratio_num > 0 && ratio_den <= 0) return slm_fail(SLM_ERR_INVALID, "ratio_den must be > 0");
    if (n_pairs == 0 || n_per_frame == 0) return SLM_OK;
    if (!desc || !pairs_host) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    if (n_per_frame > 0x7FFFFFFFll || n_pairs * n_per_frame > 0x7FFFFFFFll)
        return slm_fail(SLM_ERR_UNSUPPORTED, "batch too large");
    for (int64_t i = 0; i < 2 * n_pairs; ++i)
        if (pairs_host[i] < 0 || pairs_host[i] >= n_frames)
            return slm_fail(SLM_ERR_INVALID, "pair %lld references frame %d outside [0,%lld)", (long long)(i / 2),
                            pairs_host[i], (long long)n_frames);
    if ((uintptr_t)desc & 15) return slm_fail(SLM_ERR_INVALID, "descriptor pointers must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    SLM_TRY(slm_enter(ctx, stream));
    const int64_t rows = n_pairs * n_per_frame;
    SLM_TRY(slm_buf_reserve(ctx, &ctx->keys, (size_t)rows * 16));
    // Chain plan (tensor variant): pairs sorted by query frame, cut into units of <= L pairs that share it, longest
    // units first.  L keeps ~8 waves of clusters for balance while the cluster start-up is paid once per unit.
    std::vector<int32_t> sorted, prob, units;
    if (ctx->tc_chain_max > 1 && n_pairs >= 2 && n_pairs <= 32768) {
        // query tiles per cluster: 2 x 8 on the fp4 kernel, 2 x 4 on the fp8 kernel
        const bool fp4 = ctx->variant == SLM_VARIANT_TENSOR4 || (ctx->variant == SLM_VARIANT_AUTO && ctx->tc_fp4);
        const int64_t m_tiles = (n_per_frame + 127) / 128, n_gpairs = fp4 ? (m_tiles + 15) / 16 : (m_tiles + 7) / 8;
        int64_t L = n_pairs * n_gpairs / ((int64_t)(ctx->sm_count / 2) * 8);
        L = std::max<int64_t>(std::max(1, ctx->tc_chain_min), std::min<int64_t>(L, ctx->tc_chain_max));
        if (L > 1) {
            std::vector<int32_t> order((size_t)n_pairs);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(),
                             [&](int32_t a, int32_t b) { return pairs_host[2 * a] < pairs_host[2 * b]; });
            sorted.resize((size_t)n_pairs * 2);
            prob.resize((size_t)n_pairs);
            for (int64_t k = 0; k < n_pairs; ++k) {
                sorted[2 * k] = pairs_host[2 * order[k]];
                sorted[2 * k + 1] = pairs_host[2 * order[k] + 1];
                prob[k] = order[k];
            }
            std::vector<std::pair<int32_t, int32_t>> u;
            for (int64_t k = 0; k < n_pairs;) {
                int64_t e = k;
                while (e < n_pairs && sorted[2 * e] == sorted[2 * k] && e - k < L) ++e;
                u.emplace_back((int32_t)k, (int32_t)(e - k));
                k = e;
            }
            std::stable_sort(u.begin(), u.end(), [](const auto &a, const auto &b) { return a.second > b.second; });
            for (const auto &x : u) { units.push_back(x.first); units.push_back(x.second); }
        }
    }
    // one upload: [pairs | sorted pairs | units | prob]  (the int2-read arrays first: 8-byte aligned for any n_pairs)
    const size_t n_ints = (size_t)n_pairs * 2 + sorted.size() + prob.size() + units.size();
    SLM_TRY(slm_buf_reserve(ctx, &ctx->misc, n_ints * 4));
    SLM_TRY(pin_reserve(ctx, n_ints * 4));
    // the pinned copy must not be overwritten while a previous call's H2D is in flight
    SLM_CUDA(cudaEventSynchronize(ctx->ev[0]));
    int32_t *pin = reinterpret_cast<int32_t *>(ctx->pin);
    memcpy(pin, pairs_host, (size_t)n_pairs * 8);
    if (!units.empty()) {
        memcpy(pin + 2 * n_pairs, sorted.data(), sorted.size() * 4);
        memcpy(pin + 4 * n_pairs, units.data(), units.size() * 4);
        memcpy(pin + 4 * n_pairs + units.size(), prob.data(), prob.size() * 4);
    }
    SLM_CUDA(cudaMemcpyAsync(ctx->misc.p, ctx->pin, n_ints * 4, cudaMemcpyHostToDevice, stream));
    SLM_CUDA(cudaEventRecord(ctx->ev[0], stream));
    const int32_t *dev = reinterpret_cast<const int32_t *>(ctx->misc.p);
    slm_chain chain{dev + 2 * n_pairs, dev + 4 * n_pairs + units.size(), dev + 4 * n_pairs, (int)(units.size() / 2)};
    uint64_t *keys = reinterpret_cast<uint64_t *>(ctx->keys.p);
    SLM_TRY(slm_batched_knn2_keys(ctx, desc, n_per_frame, dev, n_pairs, keys, stream, units.empty() ? nullptr : &chain));
    return slm_finalize(ctx, keys, rows, ratio_num, ratio_den, nullptr, 0, 0, idx_out, dist_out, accept_out, stream);
}

int slm_merge_top2(slm_ctx *ctx, const uint64_t *gathered, int32_t n_shards, int64_t nq, int32_t ratio_num,
                   int32_t ratio_den, int32_t *idx_out, int32_t *dist_out, uint8_t *accept_out, void *stream_)
{
    SLM_TRY(check_ctx(ctx));
    if (n_shards <= 0 || nq < 0) return slm_fail(SLM_ERR_INVALID, "n_shards must be > 0 and nq >= 0");
    if (ratio_num > 0 && ratio_den <= 0) return slm_fail(SLM_ERR_INVALID, "ratio_den must be > 0");
    if (nq == 0) return SLM_OK;
    if (!gathered) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    SLM_TRY(slm_enter(ctx, (cudaStream_t)stream_));
    SLM_TRY(slm_prof_mark(ctx, (cudaStream_t)stream_, SLM_TAG_CALL_BEGIN));
    SLM_TRY(slm_merge_finalize(ctx, gathered, n_shards, nq, ratio_num, ratio_den, idx_out, dist_out, accept_out,
                               (cudaStream_t)stream_));
    return slm_prof_mark(ctx, (cudaStream_t)stream_, SLM_TAG_CALL_END);
}

static int check_exchange_args(int64_t nq, int64_t nq_capacity, int64_t nt_global, const uint64_t *peer_keys_host,
                               const uint64_t *peer_flags_host, int32_t rank, int32_t world, uint32_t step, int32_t ratio_num,
                               int32_t ratio_den)
{
    if (world < 1 || world > kSlmMaxWorld || rank < 0 || rank >= world)
        return slm_fail(SLM_ERR_INVALID, "bad rank/world (%d/%d; at most %d ranks)", rank, world, kSlmMaxWorld);
    if (nq < 1 || nq > nq_capacity) return slm_fail(SLM_ERR_INVALID, "nq must be in 1..nq_capacity (nq=%lld capacity=%lld)",
                                                    (long long)nq, (long long)nq_capacity);
    if (nt_global < 0) return slm_fail(SLM_ERR_INVALID, "nt_global must be >= 0");
    if (step == 0) return slm_fail(SLM_ERR_INVALID, "step starts at 1");
    if (ratio_num > 0 && ratio_den <= 0) return slm_fail(SLM_ERR_INVALID, "ratio_den must be > 0");
    if (!peer_keys_host || !peer_flags_host) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    return SLM_OK;
}

int slm_exchange_merge(slm_ctx *ctx, const uint64_t *local_keys, int64_t nq, int64_t nq_capacity, int64_t nt_global,
                       const uint64_t *peer_keys_host, const uint64_t *peer_flags_host, int32_t rank, int32_t world,
                       uint32_t step, int32_t ratio_num, int32_t ratio_den, int32_t *idx_out, int32_t *dist_out,
                       uint8_t *accept_out, void *stream_)
{
    SLM_TRY(check_ctx(ctx));
    SLM_TRY(slm_exchange_check(ctx));
    SLM_TRY(check_exchange_args(nq, nq_capacity, nt_global, peer_keys_host, peer_flags_host, rank, world, step, ratio_num, ratio_den));
    if (!local_keys) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    SLM_TRY(slm_enter(ctx, stream));
    SLM_TRY(slm_prof_mark(ctx, stream, SLM_TAG_CALL_BEGIN));
    slm_exchange ex;
    SLM_TRY(slm_exchange_setup(ctx, &ex, peer_keys_host, peer_flags_host, rank, world, step, nq_capacity, nt_global));
    SLM_TRY(slm_exchange_store(ctx, ex, local_keys, nq, stream));
    SLM_TRY(slm_exchange_wait_merge(ctx, ex, 0, nq, ratio_num, ratio_den, idx_out, dist_out, accept_out, stream));
    return slm_prof_mark(ctx, stream, SLM_TAG_CALL_END);
}

int slm_knn2_exchange(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                      int64_t nq_capacity, int64_t nt_global, const uint64_t *peer_keys_host, const uint64_t *peer_flags_host,
                      int32_t rank, int32_t world, uint32_t step, int32_t ratio_num, int32_t ratio_den, int32_t *idx_out,
                      int32_t *dist_out, uint8_t *accept_out, void *stream_)
{
    SLM_TRY(check_ctx(ctx));
    SLM_TRY(slm_exchange_check(ctx));
    SLM_TRY(check_sizes(nq, nt, base));
    SLM_TRY(check_exchange_args(nq, nq_capacity, nt_global, peer_keys_host, peer_flags_host, rank, world, step, ratio_num, ratio_den));
    if (nt_global > 0 && base + nt > nt_global) return slm_fail(SLM_ERR_INVALID, "train_index_base + nt exceeds nt_global");
    if (!q || (nt > 0 && !t)) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    SLM_TRY(slm_enter(ctx, stream));
    const bool tensor = nt > 0 && (ctx->variant == SLM_VARIANT_TENSOR || ctx->variant == SLM_VARIANT_TENSOR4 ||
                                   (ctx->variant == SLM_VARIANT_AUTO && nq > 8 && nq * nt >= (1ll << 18)));
    SLM_TRY(slm_prof_mark(ctx, stream, SLM_TAG_CALL_BEGIN));
    slm_exchange ex;
    SLM_TRY(slm_exchange_setup(ctx, &ex, peer_keys_host, peer_flags_host, rank, world, step, nq_capacity, nt_global));
    int phase = 0;
    if (tensor) {
        // the refine kernel is the producer: every query's exact keys go straight into the peers' buffers (phase 0), or --
        // many queries -- the ranks first exchange candidate-chunk keys and only the owners refine (exact keys in phase 1)
        const bool fp4 = ctx->variant == SLM_VARIANT_TENSOR4 || (ctx->variant == SLM_VARIANT_AUTO && ctx->tc_fp4);
        SLM_TRY(slm_tc_knn2_exchange(ctx, q, nq, t, nt, base, ex, stream, fp4, &phase));
    } else {
        SLM_TRY(slm_buf_reserve(ctx, &ctx->keys, (size_t)nq * 16));
        uint64_t *keys = reinterpret_cast<uint64_t *>(ctx->keys.p);
        SLM_TRY(knn2_keys_dispatch(ctx, q, nq, t, nt, base, keys, stream));
        SLM_TRY(slm_exchange_store(ctx, ex, keys, nq, stream));
    }
    SLM_TRY(slm_exchange_wait_merge(ctx, ex, phase, nq, ratio_num, ratio_den, idx_out, dist_out, accept_out, stream));
    return slm_prof_mark(ctx, stream, SLM_TAG_CALL_END);
}

int slm_exchange_status(slm_ctx *ctx)
{
    if (!ctx) return slm_fail(SLM_ERR_INVALID, "ctx is NULL");
    return slm_exchange_check(ctx);
}

int slm_filter_points3d(slm_ctx *ctx, const double *pts3d, int64_t nq, double max_distance, uint8_t *accept, void *stream)
{
    SLM_TRY(check_ctx(ctx));
    if (nq < 0) return slm_fail(SLM_ERR_INVALID, "negative size");
    if (nq == 0) return SLM_OK;
    if (!pts3d || !accept) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    SLM_TRY(slm_enter(ctx, (cudaStream_t)stream));
    return slm_filter_points3d_impl(ctx, pts3d, nq, max_distance, accept, (cudaStream_t)stream);
}

int slm_compact_matches(slm_ctx *ctx, const int32_t *idx, const int32_t *dist, const uint8_t *accept, int64_t nq,
                        int32_t stop_at_short_row, int32_t *matches_out, int32_t *count_out, void *stream)
{
    SLM_TRY(check_ctx(ctx));
    if (nq < 0) return slm_fail(SLM_ERR_INVALID, "negative size");
    if (!count_out || (nq > 0 && (!idx || !dist || !accept || !matches_out)))
        return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    SLM_TRY(slm_enter(ctx, (cudaStream_t)stream));
    return slm_compact(ctx, idx, dist, accept, nq, stop_at_short_row, matches_out, count_out, (cudaStream_t)stream);
}

int slm_gather_rows(slm_ctx *ctx, const void *src, int32_t row_bytes, const int32_t *matches, const int32_t *count,
                    int64_t capacity, int32_t column, void *out, void *stream)
{
    SLM_TRY(check_ctx(ctx));
    if (row_bytes <= 0 || (row_bytes & 3)) return slm_fail(SLM_ERR_INVALID, "row_bytes must be a positive multiple of 4");
    if (column != 0 && column != 1) return slm_fail(SLM_ERR_INVALID, "column must be 0 (queryIdx) or 1 (trainIdx)");
    if (capacity < 0) return slm_fail(SLM_ERR_INVALID, "negative capacity");
    if (capacity == 0) return SLM_OK;
    if (!src || !matches || !count || !out) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    SLM_TRY(slm_enter(ctx, (cudaStream_t)stream));
    return slm_gather(ctx, src, row_bytes, matches, count, capacity, column, out, (cudaStream_t)stream);
}

int slm_bow_hist(slm_ctx *ctx, const int32_t *words, int64_t n, int32_t stride, int32_t n_words, int32_t *hist_out,
                 void *stream)
{
    SLM_TRY(check_ctx(ctx));
    if (n < 0 || stride < 1 || n_words < 1) return slm_fail(SLM_ERR_INVALID, "bad size (n=%lld stride=%d n_words=%d)",
                                                              (long long)n, stride, n_words);
    if (!hist_out || (n > 0 && !words)) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    SLM_TRY(slm_enter(ctx, (cudaStream_t)stream));
    return slm_bow_hist_impl(ctx, words, n, stride, n_words, hist_out, (cudaStream_t)stream);
}

int slm_chi2_scan(slm_ctx *ctx, const int32_t *hist, const int32_t *db, int64_t n_db, int32_t n_words, double *dist_out,
                  int32_t *best_idx, double *best_val, void *stream)
{
    SLM_TRY(check_ctx(ctx));
    if (n_db < 0 || n_words < 1 || n_words > kChi2MaxWords) return slm_fail(SLM_ERR_INVALID, "bad size (n_db=%lld n_words=%d)",
                                                                      (long long)n_db, n_words);
    if (n_db == 0) return SLM_OK;
    if (!hist || !db || !dist_out || !best_idx || !best_val) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    SLM_TRY(slm_enter(ctx, (cudaStream_t)stream));
    return slm_chi2_scan_impl(ctx, hist, db, n_db, n_words, dist_out, best_idx, best_val, (cudaStream_t)stream);
}

int slm_vocab_update(slm_ctx *ctx, const uint32_t *desc, int64_t n, const int32_t *words, int32_t stride, uint32_t *vocab,
                     int32_t n_words, int32_t *counts_out, int32_t *changed_out, void *stream)
{
    SLM_TRY(check_ctx(ctx));
    if (n < 0 || n > 0x7FFFFFFFll || stride < 1 || n_words < 1)
        return slm_fail(SLM_ERR_INVALID, "bad size (n=%lld stride=%d n_words=%d)", (long long)n, stride, n_words);
    if (!vocab || (n > 0 && (!desc || !words))) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    SLM_TRY(slm_enter(ctx, (cudaStream_t)stream));
    return slm_vocab_update_impl(ctx, desc, n, words, stride, vocab, n_words, counts_out, changed_out, (cudaStream_t)stream);
}

int slm_knn2_host(slm_ctx *ctx, const uint8_t *q_host, int64_t nq, const uint8_t *t_host, int64_t nt,
                  int32_t ratio_num, int32_t ratio_den, int32_t cross_check, int32_t *idx_out,
                  int32_t *dist_out, uint8_t *accept_out)
{
    SLM_TRY(check_ctx(ctx));
    SLM_TRY(check_sizes(nq, nt, 0));
    if (ratio_num > 0 && ratio_den <= 0) return slm_fail(SLM_ERR_INVALID, "ratio_den must be > 0");
    if (nq == 0) return SLM_OK;
    if (!q_host || (nt > 0 && !t_host)) return slm_fail(SLM_ERR_INVALID, "NULL pointer argument");
    cudaStream_t s = ctx->own_stream;
    SLM_TRY(slm_enter(ctx, s));
    // device layout: [q | t | idx | dist | accept], every block 256-byte aligned
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t q_b = up((size_t)nq * 32), t_b = up((size_t)nt * 32), i_b = up((size_t)nq * 8), a_b = up((size_t)nq);
    SLM_TRY(slm_buf_reserve(ctx, &ctx->io, q_b + t_b + 2 * i_b + a_b));
    uint8_t *base = reinterpret_cast<uint8_t *>(ctx->io.p);
    uint32_t *q_dev = reinterpret_cast<uint32_t *>(base);
    uint32_t *t_dev = reinterpret_cast<uint32_t *>(base + q_b);
    int32_t *idx_dev = reinterpret_cast<int32_t *>(base + q_b + t_b);
    int32_t *dist_dev = reinterpret_cast<int32_t *>(base + q_b + t_b + i_b);
    uint8_t *acc_dev = base + q_b + t_b + 2 * i_b;
    // Frame-to-frame shapes (the reference's own call: ~1000 x 1000) are all latency.  Inputs are packed into
    // one pinned block and travel in a single H2D copy; the one-launch frame kernel writes idx / dist / accept
    // straight into pinned host memory (mapped into the device address space under UVA), so there is no D2H
    // copy at all: memcpy in, 1 copy, 1 kernel, 1 synchronize, memcpy out.
    const bool cross = cross_check != 0 && accept_out != nullptr && nt > 0;
    if (ctx->variant == SLM_VARIANT_AUTO && nt > 0 && (size_t)(nq + nt) * 32 <= (512u << 10) &&
        slm_frame_eligible(ctx, nq, nt, cross)) {
        SLM_TRY(pin_reserve(ctx, q_b + t_b + 2 * i_b + a_b));
        uint8_t *pin = reinterpret_cast<uint8_t *>(ctx->pin);
        SLM_CUDA(cudaEventSynchronize(ctx->ev[0]));
        memcpy(pin, q_host, (size_t)nq * 32);
        memcpy(pin + q_b, t_host, (size_t)nt * 32);
        SLM_CUDA(cudaMemcpyAsync(q_dev, pin, q_b + (size_t)nt * 32, cudaMemcpyHostToDevice, s));
        int32_t *idx_pin = reinterpret_cast<int32_t *>(pin + q_b + t_b);
        int32_t *dist_pin = reinterpret_cast<int32_t *>(pin + q_b + t_b + i_b);
        uint8_t *acc_pin = pin + q_b + t_b + 2 * i_b;
        SLM_TRY(slm_knn2_filter(ctx, q_dev, nq, t_dev, nt, 0, ratio_num, ratio_den, cross_check, idx_pin, dist_pin,
                                accept_out ? acc_pin : nullptr, s));
        SLM_CUDA(cudaStreamSynchronize(s));
        if (idx_out) memcpy(idx_out, idx_pin, (size_t)nq * 8);
        if (dist_out) memcpy(dist_out, dist_pin, (size_t)nq * 8);
        if (accept_out) memcpy(accept_out, acc_pin, (size_t)nq);
        return SLM_OK;
    }
    // results come back through pinned staging so the D2H copies are truly asynchronous
    SLM_TRY(pin_reserve(ctx, 2 * i_b + a_b));
    uint8_t *pin = reinterpret_cast<uint8_t *>(ctx->pin);
    SLM_CUDA(cudaEventSynchronize(ctx->ev[0]));  // a batched call may still be reading the pinned block

    SLM_TRY(upload_host(ctx, q_dev, q_host, (size_t)nq * 32, s));
    // Large train sets: the H2D copy (PCIe) takes longer than the search, so it is cut into chunks on a second
    // stream and every chunk is searched as soon as it has landed -- chunk = shard: per-chunk packed top-2 keys
    // are merged by global index exactly like the cross-GPU path.  (Cross-check needs the whole train set
    // resident for the reverse search, so it takes the plain path.)
    const int64_t kChunkRows = 1 << 20;          // 32 MB
    const int64_t n_chunks = (nt + kChunkRows - 1) / kChunkRows;
    if (n_chunks >= 3 && n_chunks <= kMaxHostChunks && !cross_check) {
        SLM_TRY(slm_buf_reserve(ctx, &ctx->rev, (size_t)n_chunks * nq * 16));
        uint64_t *chunk_keys = reinterpret_cast<uint64_t *>(ctx->rev.p);
        if (!ctx->copy_stream) SLM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        // a pageable train set goes through the pinned ring, filled by host threads (see HostStager)
        HostStager stager;
        const bool staged = ctx->host_threads > 0 && host_is_pageable(t_host);
        if (staged) {
            SLM_TRY(stage_reserve(ctx, (size_t)kStageSlots * kChunkRows * 32));
            if (stager.start(t_host, (size_t)nt * 32, (size_t)kChunkRows * 32, reinterpret_cast<uint8_t *>(ctx->stage_pin),
                             ctx->host_threads) != 0)
                return slm_fail(SLM_ERR_NOMEM, "could not start the host staging threads");
        }
        for (int64_t c = 0; c < n_chunks; ++c) {
            const int64_t r0 = c * kChunkRows, rows = (nt - r0 < kChunkRows) ? nt - r0 : kChunkRows;
            if (!ctx->chunk_ev[c]) SLM_CUDA(cudaEventCreateWithFlags(&ctx->chunk_ev[c], cudaEventDisableTiming));
            const uint8_t *src = t_host + r0 * 32;
            if (staged) {
                if (c >= kStageSlots) {          // the slot chunk c is written into was last read by the copy of chunk c - slots
                    SLM_CUDA(cudaEventSynchronize(ctx->stage_ev[c % kStageSlots]));
                    stager.release(c - kStageSlots + 1);
                }
                src = stager.wait(c);
            }
            SLM_CUDA(cudaMemcpyAsync(t_dev + r0 * 8, src, (size_t)rows * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (staged) SLM_CUDA(cudaEventRecord(ctx->stage_ev[c % kStageSlots], ctx->copy_stream));
            SLM_CUDA(cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream));
            SLM_CUDA(cudaStreamWaitEvent(s, ctx->chunk_ev[c], 0));
            SLM_TRY(knn2_keys_dispatch(ctx, q_dev, nq, t_dev + r0 * 8, rows, r0, chunk_keys + c * nq * 2, s));
        }
        SLM_TRY(slm_merge_finalize(ctx, chunk_keys, (int32_t)n_chunks, nq, ratio_num, ratio_den, idx_dev, dist_dev,
                                   accept_out ? acc_dev : nullptr, s));
    } else {
        SLM_TRY(upload_host(ctx, t_dev, t_host, (size_t)nt * 32, s));
        SLM_TRY(slm_knn2_filter(ctx, q_dev, nq, t_dev, nt, 0, ratio_num, ratio_den, cross_check, idx_dev, dist_dev,
                                accept_out ? acc_dev : nullptr, s));
    }
    if (idx_out && dist_out && accept_out && 2 * i_b + a_b <= (1u << 20)) {
        // small results: [idx | dist | accept] are contiguous on the device, one copy instead of three
        SLM_CUDA(cudaMemcpyAsync(pin, idx_dev, 2 * i_b + (size_t)nq, cudaMemcpyDeviceToHost, s));
    } else {
        if (idx_out) SLM_CUDA(cudaMemcpyAsync(pin, idx_dev, (size_t)nq * 8, cudaMemcpyDeviceToHost, s));
        if (dist_out) SLM_CUDA(cudaMemcpyAsync(pin + i_b, dist_dev, (size_t)nq * 8, cudaMemcpyDeviceToHost, s));
        if (accept_out) SLM_CUDA(cudaMemcpyAsync(pin + 2 * i_b, acc_dev, (size_t)nq, cudaMemcpyDeviceToHost, s));
    }
    SLM_CUDA(cudaStreamSynchronize(s));
    if (idx_out) memcpy(idx_out, pin, (size_t)nq * 8);
    if (dist_out) memcpy(dist_out, pin + i_b, (size_t)nq * 8);
    if (accept_out) memcpy(accept_out, pin + 2 * i_b, (size_t)nq);
    return SLM_OK;
}

}  // extern "C"
