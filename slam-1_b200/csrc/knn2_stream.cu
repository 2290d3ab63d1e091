// knn2_stream.cu -- integer-pipe kernel for a HANDFUL of queries against a long train set (nq <= 8).
//
// With so few queries the arithmetic intensity (nq / 32 cmp per train byte) is below
// the compute/HBM crossover (nq ~ 2-3), so the train set must simply stream through the SMs once at HBM
// speed: this is the shape BASELINE.json's north_star asks to be reported against the HBM roofline.
//   * queries live in registers (broadcast loads once);
//   * every thread walks train rows with a grid stride, 4 rows in flight (eight independent 16-byte loads,
//     consecutive lanes -> consecutive 32-byte rows, so a warp request covers whole 128-byte lines);
//   * per (row, query): 8 XOR + 8 POPC and ONE compare against the thread's current second-best distance;
//     rows arrive in increasing index order, so a strict '<' keeps the lowest index on ties and the update
//     branch is taken O(log rows) times per thread;
//   * per-thread results become packed (distance << 32 | global index) keys, reduced by warp shuffles and
//     shared memory to one partial per CTA, then by merge_keys_kernel across CTAs.
// Same exhaustive knnMatch(k=2) semantics as the other variants (tracking.py:22).
#include "slm_internal.cuh"

namespace {

constexpr int kStreamThreads = 256;
constexpr int kRowsInFlight = 4;

__device__ __forceinline__ void top2_min(unsigned long long &k1, unsigned long long &k2, unsigned long long key)
{
    unsigned long long m = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, m);
}

template <int NQ>
__global__ void __launch_bounds__(kStreamThreads) knn2_stream_kernel(const uint32_t *__restrict__ q, int nq,
                                                                     const uint4 *__restrict__ t, long long nt,
                                                                     long long base, unsigned long long *part)
{
    __shared__ unsigned long long red[kStreamThreads / 32][NQ][2];
    uint32_t qr[NQ][8];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
        const uint32_t *row = q + (long long)min(k, nq - 1) * 8;
#pragma unroll
        for (int w = 0; w < 8; ++w) qr[k][w] = __ldg(row + w);
    }
    int d1[NQ], d2[NQ];
    long long i1[NQ], i2[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) { d1[k] = 1 << 20; d2[k] = 1 << 20; i1[k] = -1; i2[k] = -1; }

    const long long stride = (long long)gridDim.x * kStreamThreads;
    const long long first = (long long)blockIdx.x * kStreamThreads + threadIdx.x;
    for (long long r0 = first; r0 < nt; r0 += stride * kRowsInFlight) {
        uint4 lo[kRowsInFlight], hi[kRowsInFlight];
#pragma unroll
        for (int u = 0; u < kRowsInFlight; ++u) {
            const long long r = r0 + u * stride;
            if (r < nt) {
                lo[u] = __ldcs(t + 2 * r);        // streaming loads: every train row is touched exactly once
                hi[u] = __ldcs(t + 2 * r + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < kRowsInFlight; ++u) {
            const long long r = r0 + u * stride;
            if (r < nt) {
#pragma unroll
                for (int k = 0; k < NQ; ++k) {
                    const int d = __popc(qr[k][0] ^ lo[u].x) + __popc(qr[k][1] ^ lo[u].y) + __popc(qr[k][2] ^ lo[u].z) +
                                  __popc(qr[k][3] ^ lo[u].w) + __popc(qr[k][4] ^ hi[u].x) + __popc(qr[k][5] ^ hi[u].y) +
                                  __popc(qr[k][6] ^ hi[u].z) + __popc(qr[k][7] ^ hi[u].w);
                    if (d < d2[k]) {
                        if (d < d1[k]) { d2[k] = d1[k]; i2[k] = i1[k]; d1[k] = d; i1[k] = r; }
                        else { d2[k] = d; i2[k] = r; }
                    }
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
        unsigned long long k1 = i1[k] < 0 ? kKeyNone : ((unsigned long long)d1[k] << 32) | (unsigned long long)(base + i1[k]);
        unsigned long long k2 = i2[k] < 0 ? kKeyNone : ((unsigned long long)d2[k] << 32) | (unsigned long long)(base + i2[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, k2, o);
            top2_min(k1, k2, o1);
            top2_min(k1, k2, o2);
        }
        if (lane == 0) { red[warp][k][0] = k1; red[warp][k][1] = k2; }
    }
    __syncthreads();
    if (threadIdx.x < NQ && (int)threadIdx.x < nq) {
        unsigned long long k1 = kKeyNone, k2 = kKeyNone;
        for (int w = 0; w < kStreamThreads / 32; ++w) {
            top2_min(k1, k2, red[w][threadIdx.x][0]);
            top2_min(k1, k2, red[w][threadIdx.x][1]);
        }
        // layout of the partials = gathered keys of merge_keys_kernel: [cta][nq][2]
        reinterpret_cast<ulonglong2 *>(part)[(long long)blockIdx.x * nq + threadIdx.x] = make_ulonglong2(k1, k2);
    }
}

}  // namespace

int slm_stream_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                         uint64_t *keys_out, cudaStream_t stream)
{
    ctx->last_variant = SLM_VARIANT_POPC;
    ctx->last_kernel = "knn2_stream_kernel";
    if (nq < 1 || nq > 8) return slm_fail(SLM_ERR_INVALID, "stream kernel handles 1..8 queries");
    long long ctas = (nt + (long long)kStreamThreads * kRowsInFlight - 1) / ((long long)kStreamThreads * kRowsInFlight);
    const long long max_ctas = (long long)ctx->sm_count * 8;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    SLM_TRY(slm_buf_reserve(ctx, &ctx->scratch, (size_t)ctas * nq * 16));
    unsigned long long *part = reinterpret_cast<unsigned long long *>(ctx->scratch.p);
    const uint4 *t4 = reinterpret_cast<const uint4 *>(t);
    SLM_TRY(slm_prof_begin(ctx, stream));
    if (nq == 1) knn2_stream_kernel<1><<<(unsigned)ctas, kStreamThreads, 0, stream>>>(q, (int)nq, t4, nt, base, part);
    else if (nq == 2) knn2_stream_kernel<2><<<(unsigned)ctas, kStreamThreads, 0, stream>>>(q, (int)nq, t4, nt, base, part);
    else if (nq <= 4) knn2_stream_kernel<4><<<(unsigned)ctas, kStreamThreads, 0, stream>>>(q, (int)nq, t4, nt, base, part);
    else knn2_stream_kernel<8><<<(unsigned)ctas, kStreamThreads, 0, stream>>>(q, (int)nq, t4, nt, base, part);
    SLM_CUDA(cudaGetLastError());
    SLM_TRY(slm_prof_end(ctx, stream));
    ctx->launches += 1;
    return slm_merge_keys(ctx, reinterpret_cast<const uint64_t *>(part), (int32_t)ctas, nq, keys_out, stream);
}
