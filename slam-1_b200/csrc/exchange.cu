// exchange.cu -- NVLink key exchange + cross-shard merge of the sharded path (BASELINE configs 4 and 5; protocol in
// exchange.cuh).  Two kernels:
//   exchange_store_kernel       producer for variants whose search leaves packed keys in local memory (the tensor
//                               path's refine kernel is its own producer, knn2_tc.cu)
//   exchange_wait_merge_kernel  acquires every rank's flag of this step (bounded poll: a lost peer is REPORTED through
//                               the ctx's status word, the kernel never traps or hangs), then every block merges its
//                               share of the queries -- top-2 of 2 x world keys in unsigned key order == (distance,
//                               global train index), OpenCV's collection order -- and applies the integer ratio test.
// The merge kernel is launched with programmatic dependent launch, so its blocks are resident and polling while the
// producer drains; it needs no griddepcontrol.wait because the flags carry the dependency (release / acquire, system scope).
#include <cstdio>
#include <cstdlib>

#include "exchange.cuh"

namespace {

__global__ void __launch_bounds__(256) exchange_store_kernel(slm_exchange ex, const unsigned long long *local_keys, long long nq)
{
    slm_pdl_launch_dependents();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) {
        const ulonglong2 k = reinterpret_cast<const ulonglong2 *>(local_keys)[i];
        for (int r = 0; r < ex.world; ++r) slm_exchange_store_to(ex, r, i, k.x, k.y);     // consecutive threads, consecutive queries
    }
    slm_exchange_publish(ex);
}

__device__ __forceinline__ void top2_insert(unsigned long long &k1, unsigned long long &k2, unsigned long long key)
{
    const unsigned long long m = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, m);
}
__device__ __forceinline__ void top2_insert32(unsigned &k1, unsigned &k2, unsigned key)
{
    const unsigned m = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, m);
}

__global__ void __launch_bounds__(256) exchange_wait_merge_kernel(slm_exchange ex, int phase, long long nq, int ratio_num,
                                                                  int ratio_den, int *idx_out, int *dist_out,
                                                                  unsigned char *accept_out)
{
    if (!slm_exchange_wait_flags(ex, phase)) return;      // outputs are left untouched; the host raises SLM_ERR_TIMEOUT
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
        unsigned long long k1 = kKeyNone, k2 = kKeyNone;
        if (ex.key_bytes == 4) {
            const uint2 *g = reinterpret_cast<const uint2 *>(ex.peer_keys[ex.rank]) + slm_exchange_slot(ex, phase, 0);
            unsigned c1 = 0xFFFFFFFFu, c2 = 0xFFFFFFFFu;
            for (int r = 0; r < ex.world; ++r) {
                const uint2 v = g[(long long)r * ex.cap + i];
                top2_insert32(c1, c2, v.x);
                top2_insert32(c1, c2, v.y);
            }
            k1 = slm_key_widen(c1);
            k2 = slm_key_widen(c2);
        } else {
            const ulonglong2 *g = reinterpret_cast<const ulonglong2 *>(ex.peer_keys[ex.rank]) + slm_exchange_slot(ex, phase, 0);
            for (int r = 0; r < ex.world; ++r) {
                const ulonglong2 v = g[(long long)r * ex.cap + i];
                top2_insert(k1, k2, v.x);
                top2_insert(k1, k2, v.y);
            }
        }
        const bool has1 = k1 != kKeyNone, has2 = k2 != kKeyNone;
        const int i1 = has1 ? (int)(k1 & 0xFFFFFFFFull) : -1, d1 = has1 ? (int)(k1 >> 32) : -1;
        const int i2 = has2 ? (int)(k2 & 0xFFFFFFFFull) : -1, d2 = has2 ? (int)(k2 >> 32) : -1;
        if (idx_out) reinterpret_cast<int2 *>(idx_out)[i] = make_int2(i1, i2);
        if (dist_out) reinterpret_cast<int2 *>(dist_out)[i] = make_int2(d1, d2);
        if (accept_out)
            accept_out[i] = (ratio_num > 0 ? (has1 && has2 && (long long)ratio_den * d1 < (long long)ratio_num * d2) : has1) ? 1 : 0;
    }
}

}  // namespace

int slm_exchange_setup(slm_ctx *ctx, slm_exchange *ex, const uint64_t *peer_keys_host, const uint64_t *peer_flags_host,
                       int32_t rank, int32_t world, uint32_t step, int64_t cap, int64_t nt_global)
{
    if (!ctx->exchange_status) {
        // mapped pinned host memory: the merge kernel reports a lost peer here, the next API call reads it
        SLM_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->exchange_status), 4 * sizeof(int), cudaHostAllocMapped));
        for (int i = 0; i < 4; ++i) ctx->exchange_status[i] = 0;
    }
    *ex = slm_exchange{};
    for (int r = 0; r < world; ++r) {
        ex->peer_keys[r] = reinterpret_cast<unsigned char *>(peer_keys_host[r]);
        ex->peer_flags[r] = reinterpret_cast<unsigned *>(peer_flags_host[r]);
    }
    ex->rank = rank;
    ex->world = world;
    ex->step = step;
    ex->cap = cap;
    // every global train index < 65 536 (config 4's vocabulary): 32-bit keys, half the NVLink bytes
    ex->key_bytes = (nt_global > 0 && nt_global <= 65536 && !ctx->exchange_wide_keys) ? 4 : 8;
    ex->done_counter = ctx->done_counter;
    ex->max_polls = ctx->exchange_max_polls;
    SLM_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&ex->status), ctx->exchange_status, 0));
    return SLM_OK;
}

int slm_exchange_store(slm_ctx *ctx, const slm_exchange &ex, const uint64_t *local_keys, int64_t nq, cudaStream_t stream)
{
    SLM_CUDA(slm_launch(exchange_store_kernel, dim3((unsigned)((nq + 255) / 256)), dim3(256), 0, stream, false, ex,
                        reinterpret_cast<const unsigned long long *>(local_keys), (long long)nq));
    ctx->launches += 1;
    return SLM_OK;
}

int slm_exchange_wait_merge(slm_ctx *ctx, const slm_exchange &ex, int phase, int64_t nq, int32_t ratio_num, int32_t ratio_den,
                            int32_t *idx_out, int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream)
{
    long long blocks = (nq + 255) / 256;
    if (blocks > ctx->exchange_max_blocks) blocks = ctx->exchange_max_blocks;
    if (blocks < 1) blocks = 1;
    SLM_CUDA(slm_launch(exchange_wait_merge_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, /*pdl=*/!ctx->no_pdl, ex, phase,
                        (long long)nq, (int)ratio_num, (int)ratio_den, idx_out, dist_out, accept_out));
    ctx->launches += 1;
    return SLM_OK;
}

int slm_exchange_check(slm_ctx *ctx)
{
    if (ctx->exchange_status && ctx->exchange_status[0] != 0) {
        const int r = ctx->exchange_status[1], step = ctx->exchange_status[2], seen = ctx->exchange_status[3];
        ctx->exchange_status[0] = 0;
        return slm_fail(SLM_ERR_TIMEOUT, "sharded exchange: rank %d never delivered its keys of step %d (its flag shows %d); "
                        "the results of that step were not written", r, step, seen);
    }
    return SLM_OK;
}
