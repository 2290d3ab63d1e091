// knn2_masked.cu -- knnMatch(q, t, k=2, mask=M): the search restricted to the pairs a caller's mask allows.
//
// OpenCV's DescriptorMatcher honours an optional mask uint8[nq][nt]: pair (i, j) takes
// part iff M[i][j] != 0; a query with fewer than two allowed train rows gets a short (or empty) result row.  The
// reference never passes one (tracking.py:22, keypoint.py:44, Point3D.py:40), so this is the boundary's optional
// argument, not the hot path: one integer-pipe kernel, no tensor variant (a masked pair cannot be dropped from an
// MMA tile's running maximum without reading the mask anyway, and the mask is one byte per pair -- 8 x the train
// bytes of a 2000-query call -- so the kernel is bound by the XOR + POPC work and the mask stream, not by the tile shape).
//   * one warp owns kMaskQ = 4 queries (32 registers) and walks a slice of the train rows: lane l takes rows l, l + 32, ...
//     -- consecutive lanes read consecutive 32-byte rows (whole 128-byte lines per request) and consecutive mask bytes
//     (one 32-byte sector per query and step);
//   * the eight warps of a block walk the SAME slice for different queries, so the train rows are fetched from L2
//     once per block and served from L1 to the other warps;
//   * rows arrive in increasing index order per lane: a strict '<' keeps the lowest index on ties; lanes, then train
//     slices, are merged through packed (distance << 32 | global index) keys, whose unsigned order is OpenCV's
//     (distance, trainIdx) order (merge_finalize_kernel, the cross-shard merge of finalize.cu).
#include "slm_internal.cuh"

namespace {

constexpr int kMaskThreads = 256;
constexpr int kMaskQ = 4;                                    // queries per warp
constexpr int kMaskQPB = (kMaskThreads / 32) * kMaskQ;       // queries per block
constexpr int kMaskRows = 2;                                 // train rows per lane in flight

__device__ __forceinline__ void top2_min(unsigned long long &k1, unsigned long long &k2, unsigned long long key)
{
    const unsigned long long m = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, m);
}

__global__ void __launch_bounds__(kMaskThreads) knn2_masked_kernel(const uint32_t *__restrict__ q, int nq,
                                                                   const uint4 *__restrict__ t, int nt, long long base,
                                                                   const unsigned char *__restrict__ mask, long long mask_stride,
                                                                   int rows_per_slice, unsigned long long *part)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = (blockIdx.x * (kMaskThreads / 32) + warp) * kMaskQ;
    if (q0 >= nq) return;                                    // whole warps leave; nothing below synchronises the block
    const int r_begin = blockIdx.y * rows_per_slice;
    const int r_end = min(nt, r_begin + rows_per_slice);
    uint32_t qr[kMaskQ][8];
    const unsigned char *mrow[kMaskQ];
#pragma unroll
    for (int k = 0; k < kMaskQ; ++k) {
        const int qi = min(q0 + k, nq - 1);
        const uint4 *row = reinterpret_cast<const uint4 *>(q + (long long)qi * 8);
        const uint4 a = __ldg(row), b = __ldg(row + 1);
        qr[k][0] = a.x; qr[k][1] = a.y; qr[k][2] = a.z; qr[k][3] = a.w;
        qr[k][4] = b.x; qr[k][5] = b.y; qr[k][6] = b.z; qr[k][7] = b.w;
        mrow[k] = mask + (long long)qi * mask_stride;
    }
    int d1[kMaskQ], d2[kMaskQ], i1[kMaskQ], i2[kMaskQ];
#pragma unroll
    for (int k = 0; k < kMaskQ; ++k) { d1[k] = 1 << 20; d2[k] = 1 << 20; i1[k] = -1; i2[k] = -1; }

    for (int r0 = r_begin + lane; r0 < r_end; r0 += 32 * kMaskRows) {
        uint4 lo[kMaskRows], hi[kMaskRows];
        unsigned char m[kMaskRows][kMaskQ];
#pragma unroll
        for (int u = 0; u < kMaskRows; ++u) {
            const int r = min(r0 + u * 32, r_end - 1);           // clamped: the row is ignored below when out of range
            lo[u] = __ldg(t + 2ll * r);
            hi[u] = __ldg(t + 2ll * r + 1);
#pragma unroll
            for (int k = 0; k < kMaskQ; ++k) m[u][k] = __ldg(mrow[k] + r);
        }
#pragma unroll
        for (int u = 0; u < kMaskRows; ++u) {
            const int r = r0 + u * 32;
            if (r < r_end) {
#pragma unroll
                for (int k = 0; k < kMaskQ; ++k) {
                    if (m[u][k]) {
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
    }
#pragma unroll
    for (int k = 0; k < kMaskQ; ++k) {
        unsigned long long k1 = i1[k] < 0 ? kKeyNone : ((unsigned long long)d1[k] << 32) | (unsigned long long)(base + i1[k]);
        unsigned long long k2 = i2[k] < 0 ? kKeyNone : ((unsigned long long)d2[k] << 32) | (unsigned long long)(base + i2[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, k2, o);
            top2_min(k1, k2, o1);
            top2_min(k1, k2, o2);
        }
        // layout of the partials = gathered keys of merge_finalize_kernel: [slice][nq][2]
        if (lane == 0 && q0 + k < nq)
            reinterpret_cast<ulonglong2 *>(part)[(long long)blockIdx.y * nq + q0 + k] = make_ulonglong2(k1, k2);
    }
}

}  // namespace

int slm_masked_knn2(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                    const uint8_t *mask, int64_t mask_stride, int32_t ratio_num, int32_t ratio_den, int32_t *idx_out,
                    int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream)
{
    ctx->last_variant = SLM_VARIANT_POPC;
    ctx->last_kernel = "knn2_masked_kernel";
    const long long q_blocks = (nq + kMaskQPB - 1) / kMaskQPB;
    // (query block, train slice) pairs: at most two full waves of the 2 resident blocks per SM (102 registers x 256 threads) --
    // rounding the slice count UP left a third wave that was 13 % full on the 2000 x 20000 shape; a slice is at least 256 rows
    long long slices = (4ll * ctx->sm_count) / q_blocks;
    const long long max_slices = (nt + 255) / 256;
    if (slices > max_slices) slices = max_slices;
    if (slices > 4096) slices = 4096;
    if (slices < 1) slices = 1;
    long long rows_per_slice = (nt + slices - 1) / slices;
    rows_per_slice = (rows_per_slice + 31) / 32 * 32;
    slices = (nt + rows_per_slice - 1) / rows_per_slice;
    SLM_TRY(slm_buf_reserve(ctx, &ctx->scratch, (size_t)slices * nq * 16));
    unsigned long long *part = reinterpret_cast<unsigned long long *>(ctx->scratch.p);
    SLM_TRY(slm_prof_begin(ctx, stream));
    knn2_masked_kernel<<<dim3((unsigned)q_blocks, (unsigned)slices), kMaskThreads, 0, stream>>>(
        q, (int)nq, reinterpret_cast<const uint4 *>(t), (int)nt, base, mask, mask_stride, (int)rows_per_slice, part);
    SLM_CUDA(cudaGetLastError());
    SLM_TRY(slm_prof_end(ctx, stream));
    ctx->launches += 1;
    return slm_merge_finalize(ctx, reinterpret_cast<const uint64_t *>(part), (int32_t)slices, nq, ratio_num, ratio_den, idx_out,
                              dist_out, accept_out, stream);
}
