// probe.cu -- in-process ceilings of the pipes the distance kernels run on (slm_probe_* in slammatch.h).
//
// bench.py divides the achieved rate by a ceiling measured in the SAME process, on the same GPU, right after the timed
// region (same clocks, same power state) instead of a constant from an earlier session:
//   kind 0  tcgen05.mma kind::f8f6f4 (variant T):  one CTA per SM issues back-to-back 128 x 256 x 32 MMAs from shared memory
//   kind 1  tcgen05.mma kind::mxf4.block_scale (variant T4): one CTA pair per SM pair, 256 x 240 x 64 MMAs, scales = 1.0
//   POPC    the inner loop of variant P (8 XOR + 8 POPC + adds + packed top-2 update) on every SM
// The operands are whatever shared memory holds (the values do not change the issue rate); nothing is read back.
#include <algorithm>
#include <vector>

#include "slm_internal.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kProbeThreads = 128;

struct ProbeBar {
    uint64_t bar;
    uint32_t tmem_base;
};

__device__ __forceinline__ void probe_fill_smem(uint8_t *smem, int bytes, uint32_t word)
{
    for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4 *>(smem + i) = make_uint4(word, word, word, word);
}

__global__ void __launch_bounds__(kProbeThreads, 1) probe_f8_kernel(long long *cycles, int loops)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t kA = 128 * 256, kB = 256 * 256;
    ProbeBar *pb = reinterpret_cast<ProbeBar *>(smem + kA + kB);
    const int warp = threadIdx.x >> 5;
    probe_fill_smem(smem, kA + kB, tc::kFp8PlusOne);
    if (threadIdx.x == 0) {
        tc::mbar_init(&pb->bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(&pb->tmem_base, 512);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = pb->tmem_base;
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::idesc_e4m3_f32(128, 256);
        const uint32_t a_lo = tc::smem_desc_lo(tc::smem_u32(smem)), b_lo = tc::smem_desc_lo(tc::smem_u32(smem + kA));
        for (int l = 0; l < loops; ++l) tc::umma_job<1>(tmem + (l & 1) * 256, a_lo, b_lo, idesc);
        tc::umma_commit(&pb->bar);
    }
    tc::mbar_wait(&pb->bar, 0, 90);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kProbeThreads, 1) probe_mxf4_kernel(long long *cycles, int loops)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t kA = 128 * 128, kB = 120 * 128;
    ProbeBar *pb = reinterpret_cast<ProbeBar *>(smem + kA + kB);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = tc::cluster_ctarank();
    probe_fill_smem(smem, kA + kB, tc4::kE2m1PlusOne);
    if (threadIdx.x == 0) {
        tc::mbar_init(&pb->bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc_2cta(&pb->tmem_base, 512);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    tc::cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem = pb->tmem_base;
    tc4::tmem_fill32(tmem + ((uint32_t)(warp * 32) << 16) + tc4::kSfCol, tc4::kScaleOnes);   // 4 warps = 128 lanes
    tc::tc_fence_before();
    tc::cluster_sync();
    tc::tc_fence_after();
    const long long t0 = clock64();
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = tc4::idesc_mxf4(256, tc4::kTileN);
        const uint32_t a_lo = tc4::smem_desc_lo(tc::smem_u32(smem)), b_lo = tc4::smem_desc_lo(tc::smem_u32(smem + kA));
        for (int l = 0; l < loops; ++l) tc4::umma_job(tmem + (l & 1) * tc4::kTileN, a_lo, b_lo, idesc, tmem + tc4::kSfCol);
        tc::umma_commit_2cta(&pb->bar, 3);
    }
    tc::mbar_wait(&pb->bar, 0, 91);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    tc::tc_fence_before();
    tc::cluster_sync();
    if (warp == 0) tc::tmem_dealloc_2cta(tmem, 512);
}

constexpr int kPopcIters = 4096;
__global__ void __launch_bounds__(1024) probe_popc_kernel(unsigned *out, long long *cycles, unsigned seed)
{
    unsigned q[8], b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
#pragma unroll
    for (int w = 0; w < 8; ++w) q[w] = seed * (threadIdx.x + 1) + w * 0x9E3779B9u;
    unsigned tw = seed;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < kPopcIters; ++it) {
        unsigned d = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) d += __popc(q[w] ^ (tw + w * 0x85EBCA6Bu));
        const unsigned key = (d << 23) + it;
        const unsigned m = max(b1, key);
        b1 = min(b1, key);
        b2 = min(b2, m);
        tw = tw * 1664525u + 1013904223u;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = b1 ^ b2;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

}  // namespace

extern "C" {

int slm_probe_tensor_peak(slm_ctx *ctx, int32_t kind, int32_t loops, int32_t reps, double *tflops_out,
                          double *mac_per_clk_per_sm_out)
{
    if (!ctx) return slm_fail(SLM_ERR_INVALID, "ctx is NULL");
    if (kind != 0 && kind != 1) return slm_fail(SLM_ERR_INVALID, "kind must be 0 (f8f6f4) or 1 (mxf4)");
    if (loops < 1 || loops > (1 << 20) || reps < 1 || reps > 1000) return slm_fail(SLM_ERR_INVALID, "bad loops / reps");
    SLM_CUDA(cudaSetDevice(ctx->device));
    const int sms = kind == 0 ? ctx->sm_count : ctx->sm_count / 2 * 2;
    SLM_TRY(slm_buf_reserve(ctx, &ctx->misc, (size_t)sms * sizeof(long long)));
    long long *cyc = reinterpret_cast<long long *>(ctx->misc.p);
    const size_t smem = (kind == 0 ? 128 * 256 + 256 * 256 : 128 * 128 + 120 * 128) + 64;
    if (kind == 0) SLM_CUDA(cudaFuncSetAttribute(probe_f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else SLM_CUDA(cudaFuncSetAttribute(probe_mxf4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t s = ctx->own_stream;
    cudaEvent_t e0, e1;
    SLM_CUDA(cudaEventCreate(&e0));
    SLM_CUDA(cudaEventCreate(&e1));
    auto launch = [&]() {
        if (kind == 0) probe_f8_kernel<<<sms, kProbeThreads, smem, s>>>(cyc, loops);
        else probe_mxf4_kernel<<<sms, kProbeThreads, smem, s>>>(cyc, loops);
    };
    launch();                                   // warm-up (module load, clocks)
    double best_ms = 1e30;
    for (int r = 0; r < reps; ++r) {
        SLM_CUDA(cudaEventRecord(e0, s));
        launch();
        SLM_CUDA(cudaEventRecord(e1, s));
        SLM_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SLM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best_ms = std::min(best_ms, (double)ms);
    }
    SLM_CUDA(cudaGetLastError());
    std::vector<long long> h((size_t)sms);
    SLM_CUDA(cudaMemcpyAsync(h.data(), cyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, s));
    SLM_CUDA(cudaStreamSynchronize(s));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::sort(h.begin(), h.end());
    // MACs per SM per launch: f8: 128 x 256 x 256 per job on one SM; mxf4: 256 x 240 x 256 per job on a pair
    const double macs_per_sm = kind == 0 ? (double)loops * 128 * 256 * 256 : (double)loops * 128 * tc4::kTileN * 256;
    if (mac_per_clk_per_sm_out) *mac_per_clk_per_sm_out = macs_per_sm / (double)h[h.size() / 2];
    if (tflops_out) *tflops_out = 2.0 * macs_per_sm * sms / (best_ms * 1e-3) / 1e12;
    ctx->launches += reps + 1;
    return SLM_OK;
}

int slm_probe_popc_peak(slm_ctx *ctx, int32_t reps, double *tcmp_per_s_out, double *popc_lanes_per_clk_per_sm_out)
{
    if (!ctx) return slm_fail(SLM_ERR_INVALID, "ctx is NULL");
    if (reps < 1 || reps > 1000) return slm_fail(SLM_ERR_INVALID, "bad reps");
    SLM_CUDA(cudaSetDevice(ctx->device));
    const int sms = ctx->sm_count, threads = 1024;
    SLM_TRY(slm_buf_reserve(ctx, &ctx->misc, (size_t)sms * sizeof(long long) + (size_t)sms * threads * 4));
    long long *cyc = reinterpret_cast<long long *>(ctx->misc.p);
    unsigned *out = reinterpret_cast<unsigned *>(cyc + sms);
    cudaStream_t s = ctx->own_stream;
    cudaEvent_t e0, e1;
    SLM_CUDA(cudaEventCreate(&e0));
    SLM_CUDA(cudaEventCreate(&e1));
    probe_popc_kernel<<<sms, threads, 0, s>>>(out, cyc, 12345u);
    double best_ms = 1e30;
    for (int r = 0; r < reps; ++r) {
        SLM_CUDA(cudaEventRecord(e0, s));
        probe_popc_kernel<<<sms, threads, 0, s>>>(out, cyc, 54321u + r);
        SLM_CUDA(cudaEventRecord(e1, s));
        SLM_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SLM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best_ms = std::min(best_ms, (double)ms);
    }
    SLM_CUDA(cudaGetLastError());
    std::vector<long long> h((size_t)sms);
    SLM_CUDA(cudaMemcpyAsync(h.data(), cyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, s));
    SLM_CUDA(cudaStreamSynchronize(s));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::sort(h.begin(), h.end());
    const double cmps_per_sm = (double)kPopcIters * threads;
    if (popc_lanes_per_clk_per_sm_out) *popc_lanes_per_clk_per_sm_out = 8.0 * cmps_per_sm / (double)h[h.size() / 2];
    if (tcmp_per_s_out) *tcmp_per_s_out = cmps_per_sm * sms / (best_ms * 1e-3) / 1e12;
    ctx->launches += reps + 1;
    return SLM_OK;
}

}  // extern "C"
