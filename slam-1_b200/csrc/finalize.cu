// finalize.cu -- everything after the distance kernels, shared by all variants:
//   * packed keys -> (idx, dist) + the reference's Lowe ratio test in integer form
//     (`if m.distance < 0.7 * n.distance`, tracking.py:27 / keypoint.py:48 / Point3D.py:44;
//     den*d1 < num*d2 is exact for every d in 0..256) + cross-check;
//   * cross-shard merge of all-gathered keys (order = (distance, global train index));
//   * ordered compaction of accepted rows (the `good` list of tracking.py:24-30).
#include <cstdio>
#include "slm_internal.cuh"

namespace {

__global__ void finalize_kernel(const unsigned long long *keys, long long n, int ratio_num, int ratio_den,
                                const unsigned long long *rev_keys, long long nt, long long base,
                                int *idx_out, int *dist_out, unsigned char *accept_out, int rev_by_query)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ulonglong2 k = reinterpret_cast<const ulonglong2 *>(keys)[i];
    const bool has1 = k.x != kKeyNone, has2 = k.y != kKeyNone;
    int i1 = has1 ? (int)(k.x & 0xFFFFFFFFull) : -1, d1 = has1 ? (int)(k.x >> 32) : -1;
    int i2 = has2 ? (int)(k.y & 0xFFFFFFFFull) : -1, d2 = has2 ? (int)(k.y >> 32) : -1;
    if (idx_out) reinterpret_cast<int2 *>(idx_out)[i] = make_int2(i1, i2);
    if (dist_out) reinterpret_cast<int2 *>(dist_out)[i] = make_int2(d1, d2);
    if (accept_out) {
        bool ok = ratio_num > 0 ? (has1 && has2 && (long long)ratio_den * d1 < (long long)ratio_num * d2) : has1;
        if (ok && rev_keys != nullptr) {
            if (rev_by_query) {
                ok = (long long)(rev_keys[2 * i] & 0xFFFFFFFFull) == i;
            } else {
                long long j = (long long)i1 - base;
                ok = j >= 0 && j < nt && (long long)(rev_keys[2 * j] & 0xFFFFFFFFull) == i;
            }
        }
        accept_out[i] = ok ? 1 : 0;
    }
}

// Rows of the train set that are some query's best match, in query order (input of the reduced reverse search).
__global__ void gather_best_rows_kernel(const unsigned long long *keys, long long n, long long base, const uint4 *t,
                                        uint4 *out)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 16-byte half row
    if (g >= 2 * n) return;
    const unsigned long long k = keys[2 * (g >> 1)];
    out[g] = k == kKeyNone ? make_uint4(0, 0, 0, 0) : __ldg(t + 2 * ((long long)(k & 0xFFFFFFFFull) - base) + (g & 1));
}

__global__ void merge_keys_kernel(const unsigned long long *gathered, int n_shards, long long nq,
                                  unsigned long long *keys_out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    unsigned long long k1 = kKeyNone, k2 = kKeyNone;
    for (int s = 0; s < n_shards; ++s) {
        ulonglong2 v = reinterpret_cast<const ulonglong2 *>(gathered)[(long long)s * nq + i];
        unsigned long long m = max(k1, v.x);
        k1 = min(k1, v.x);
        k2 = min(k2, m);
        m = max(k1, v.y);
        k1 = min(k1, v.y);
        k2 = min(k2, m);
    }
    reinterpret_cast<ulonglong2 *>(keys_out)[i] = make_ulonglong2(k1, k2);
}

// Cross-shard merge fused with the finalize step (one launch on the latency-critical sharded path).
__global__ void merge_finalize_kernel(const unsigned long long *gathered, int n_shards, long long nq, int ratio_num,
                                      int ratio_den, int *idx_out, int *dist_out, unsigned char *accept_out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    unsigned long long k1 = kKeyNone, k2 = kKeyNone;
    for (int s = 0; s < n_shards; ++s) {
        ulonglong2 v = reinterpret_cast<const ulonglong2 *>(gathered)[(long long)s * nq + i];
        unsigned long long m = max(k1, v.x);
        k1 = min(k1, v.x);
        k2 = min(k2, m);
        m = max(k1, v.y);
        k1 = min(k1, v.y);
        k2 = min(k2, m);
    }
    const bool has1 = k1 != kKeyNone, has2 = k2 != kKeyNone;
    const int i1 = has1 ? (int)(k1 & 0xFFFFFFFFull) : -1, d1 = has1 ? (int)(k1 >> 32) : -1;
    const int i2 = has2 ? (int)(k2 & 0xFFFFFFFFull) : -1, d2 = has2 ? (int)(k2 >> 32) : -1;
    if (idx_out) reinterpret_cast<int2 *>(idx_out)[i] = make_int2(i1, i2);
    if (dist_out) reinterpret_cast<int2 *>(dist_out)[i] = make_int2(d1, d2);
    if (accept_out)
        accept_out[i] = (ratio_num > 0 ? (has1 && has2 && (long long)ratio_den * d1 < (long long)ratio_num * d2) : has1) ? 1 : 0;
}

// Point3D.find_2D_and_3D_correspondenses (Point3D.py:45-46): a match is kept only if the query's triangulated point has
// |X|, |Y|, |Z| < max_Distance (strict, float64 like the reference's numpy array).
__global__ void filter_points3d_kernel(const double *pts, long long n, double max_distance, unsigned char *accept)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !accept[i]) return;
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (!(fabs(x) < max_distance && fabs(y) < max_distance && fabs(z) < max_distance)) accept[i] = 0;
}

// Ordered compaction by one CTA: block-wide exclusive scan over 1024-row chunks.
constexpr int kCompactThreads = 1024;
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const int *idx, const int *dist, const unsigned char *accept, long long nq, int stop_at_short_row,
               int *matches_out, int *count_out)
{
    __shared__ int warp_sums[32];
    __shared__ int chunk_base;
    __shared__ long long first_short;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { chunk_base = 0; first_short = nq; }
    __syncthreads();
    if (stop_at_short_row) {
        long long mine = nq;
        for (long long i = tid; i < nq; i += kCompactThreads)
            if (idx[2 * i] < 0 || idx[2 * i + 1] < 0) { mine = i; break; }
        if (mine < nq) atomicMin(&first_short, mine);
        __syncthreads();
    }
    const long long limit = first_short;
    for (long long c0 = 0; c0 < limit; c0 += kCompactThreads) {
        long long i = c0 + tid;
        int flag = (i < limit && accept[i]) ? 1 : 0;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (lane >= o) wi += v;
            }
            warp_sums[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        int pos = chunk_base + warp_sums[warp] + incl - flag;
        if (flag) {
            matches_out[3 * pos + 0] = (int)i;
            matches_out[3 * pos + 1] = idx[2 * i];
            matches_out[3 * pos + 2] = dist[2 * i];
        }
        __syncthreads();
        if (tid == kCompactThreads - 1) chunk_base = pos + flag;
        __syncthreads();
    }
    if (tid == 0) *count_out = chunk_base;
}

// out[m] = src[matches[m][column]] in 4-byte words; one thread per output word.
__global__ void gather_rows_kernel(const unsigned *src, int row_words, const int *matches, const int *count,
                                   long long capacity, int column, unsigned *out)
{
    const long long n = min((long long)*count, capacity) * row_words;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / row_words;
        const int w = (int)(i % row_words);
        out[i] = src[(long long)matches[3 * m + column] * row_words + w];
    }
}

}  // namespace

int slm_gather(slm_ctx *ctx, const void *src, int32_t row_bytes, const int32_t *matches, const int32_t *count,
               int64_t capacity, int32_t column, void *out, cudaStream_t stream)
{
    if (capacity <= 0) return SLM_OK;
    const int row_words = row_bytes / 4;
    long long blocks = (capacity * row_words + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const unsigned *>(src), row_words, matches,
                                                             count, capacity, column, reinterpret_cast<unsigned *>(out));
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SLM_OK;
}

int slm_finalize(slm_ctx *ctx, const uint64_t *keys, int64_t n, int32_t ratio_num, int32_t ratio_den,
                 const uint64_t *rev_keys, int64_t nt, int64_t train_index_base, int32_t *idx_out,
                 int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream, int rev_by_query)
{
    if (n <= 0) return SLM_OK;
    finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const unsigned long long *>(keys), n, ratio_num, ratio_den,
        reinterpret_cast<const unsigned long long *>(rev_keys), nt, train_index_base, idx_out, dist_out, accept_out,
        rev_by_query);
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SLM_OK;
}

int slm_gather_best_rows(slm_ctx *ctx, const uint64_t *keys, int64_t n, int64_t train_index_base, const uint32_t *t,
                         uint32_t *out, cudaStream_t stream)
{
    if (n <= 0) return SLM_OK;
    gather_best_rows_kernel<<<(unsigned)((2 * n + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const unsigned long long *>(keys), n, train_index_base, reinterpret_cast<const uint4 *>(t),
        reinterpret_cast<uint4 *>(out));
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SLM_OK;
}

int slm_merge_keys(slm_ctx *ctx, const uint64_t *gathered, int32_t n_shards, int64_t nq, uint64_t *keys_out,
                   cudaStream_t stream)
{
    if (nq <= 0) return SLM_OK;
    merge_keys_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const unsigned long long *>(gathered), n_shards, nq,
        reinterpret_cast<unsigned long long *>(keys_out));
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SLM_OK;
}

int slm_merge_finalize(slm_ctx *ctx, const uint64_t *gathered, int32_t n_shards, int64_t nq, int32_t ratio_num,
                       int32_t ratio_den, int32_t *idx_out, int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream)
{
    if (nq <= 0) return SLM_OK;
    merge_finalize_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const unsigned long long *>(gathered), n_shards, nq, ratio_num, ratio_den, idx_out, dist_out,
        accept_out);
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SLM_OK;
}

int slm_filter_points3d_impl(slm_ctx *ctx, const double *pts3d, int64_t n, double max_distance, uint8_t *accept,
                             cudaStream_t stream)
{
    filter_points3d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pts3d, n, max_distance, accept);
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SLM_OK;
}

int slm_compact(slm_ctx *ctx, const int32_t *idx, const int32_t *dist, const uint8_t *accept, int64_t nq,
                int32_t stop_at_short_row, int32_t *matches_out, int32_t *count_out, cudaStream_t stream)
{
    compact_kernel<<<1, kCompactThreads, 0, stream>>>(idx, dist, accept, nq, stop_at_short_row, matches_out, count_out);
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SLM_OK;
}
