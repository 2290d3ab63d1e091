// bow.cu -- the step AFTER word assignment: per-image word histogram and the
// chi-square scan of a query histogram against every stored keyframe histogram.
//
// Reference (bag_of_words.py):
//   hist  :23-26   labels = kmeans.predict(descriptors); np.histogram(labels, bins=k, range=(0, k-1))
//                  (for integer labels in [0, k) that binning is the identity: floor(l*k/(k-1)) = l, the last
//                  bin is closed on the right) -- here labels are the Hamming word assignment (config 4).
//   chi2  :30-31   np.sum(2 * (x - y)**2 / np.maximum(1, x + y))      integer histograms, float64 result
//   scan  :38-42   dist over db[0 : i+1-threshold]; (np.argmin(dist), np.min(dist))
// The scan is the one HBM-bound loop of the pipeline (4*k bytes per stored keyframe, no reuse).
// Results are BIT-exact with numpy: the terms are IEEE double divisions of exactly representable integers
// and the sum follows numpy's pairwise_sum order (8 accumulators for n <= 128, a binary tree of such leaves above).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <numeric>
#include <vector>
#include "slm_internal.cuh"
#include "chi2_plan.h"

namespace {

__global__ void bow_hist_kernel(const int *idx, long long n, int idx_stride, int n_words, int *hist)
{
    extern __shared__ int sh[];
    const bool use_smem = n_words <= 8192;
    if (use_smem) {
        for (int w = threadIdx.x; w < n_words; w += blockDim.x) sh[w] = 0;
        __syncthreads();
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int w = idx[i * idx_stride];
        if (w >= 0 && w < n_words) atomicAdd(use_smem ? &sh[w] : &hist[w], 1);
    }
    if (use_smem) {
        __syncthreads();
        for (int w = threadIdx.x; w < n_words; w += blockDim.x)
            if (sh[w]) atomicAdd(&hist[w], sh[w]);
    }
}

__device__ __forceinline__ double chi2_term(int x, int y)
{
    // numpy evaluates 2 * (x - y)**2 and max(1, x + y) in int64 and divides in float64.  Same values here, with the two cases
    // that make up almost every term of a sparse histogram (config 4: 2000 descriptors over 65 536 words) taken out of the
    // division: a zero numerator gives +0.0 and a denominator of 1 gives the numerator itself, both exactly what the IEEE
    // division returns -- and a zero operand sent every call of the double-precision division down its slow path (ncu:
    // 123 warp instructions per 32 terms, issue-bound).  Small counts use 32-bit arithmetic and conversions.
    if (((x | y) >= 0) && x < 32768 && y < 32768) {
        const int d = x - y;
        const unsigned num = 2u * (unsigned)(d * d), den = (unsigned)(x + y);        // < 2^31, < 2^16
        if (num == 0u) return 0.0;
        if (den <= 1u) return (double)num;
        return __ddiv_rn((double)num, (double)den);
    }
    const long long d = (long long)x - (long long)y;
    const long long num = 2 * d * d;                 // exact in int64, like numpy's integer arithmetic
    const long long den = max(1ll, (long long)x + (long long)y);
    return __ddiv_rn((double)num, (double)den);      // true divide: both converted to float64
}

// numpy's pairwise_sum of the float64 vector whose i-th element is chi2_term(hq[i], row[i]):
//   n < 8     sequential;
//   n <= 128  ("leaf") eight accumulators r[j] += a[i + j] for i = 0, 8, 16, ...; ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); the
//             n % 8 trailing elements are added one by one;
//   n > 128   sum(first n2) + sum(rest), n2 = n / 2 rounded down to a multiple of 8 -- a binary tree over leaves.
// A leaf is summed by an 8-LANE GROUP, lane j owning accumulator r[j]: every step of a group reads 8 consecutive words of
// the stored histogram (one 32-byte sector) and of the query histogram, so a warp request covers four full sectors instead
// of 32 scattered words (one thread per stored histogram -- the first version of this scan -- reached 4 % of the HBM
// bandwidth and needed device recursion); the three XOR-shuffle additions reproduce the association above exactly
// (IEEE addition is commutative, so lane 1's r1 + r0 is lane 0's r0 + r1).  The result is valid in every lane of the group.
// Memory-level parallelism: a leaf is at most 128 words = 16 steps of a group, and ALL of a leaf's loads are issued before the
// first term is evaluated (the first form of this loop loaded, divided and added step by step: one 32-byte sector in flight per
// group, 0.16-0.22 of the HBM bandwidth however sparse the histograms were).  Steps past the end of the leaf load nothing
// and add nothing; every accumulator starts at +0.0 instead of at its first term, which is the same value bit for bit
// (+0.0 + t = t for every t >= +0.0).
constexpr int kChi2LeafSteps = 16;                       // kChi2LeafWords / 8

__device__ __forceinline__ double chi2_leaf8(const int *hq, const int *row, int n, int j, unsigned group_mask)
{
    double r = 0.0;
    const int steps = n >> 3;                            // full steps of the eight accumulators (0 when n < 8)
    if (steps > 0) {
        int x[kChi2LeafSteps], y[kChi2LeafSteps];
#pragma unroll
        for (int s = 0; s < kChi2LeafSteps; ++s) {
            const bool in = s < steps;
            x[s] = in ? __ldg(hq + s * 8 + j) : 0;
            y[s] = in ? __ldg(row + s * 8 + j) : 0;
        }
#pragma unroll
        for (int s = 0; s < kChi2LeafSteps; ++s)
            if (s < steps) r = __dadd_rn(r, chi2_term(x[s], y[s]));      // (+0.0 + first term = first term: numpy starts there)
        r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 1));
        r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 2));
        r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 4));
    }
    // the n % 8 trailing elements (all of them when n < 8) are added one by one, by every lane of the group
    const int n8 = steps << 3;
    if (n8 < n) {
        int tx[7], ty[7];
#pragma unroll
        for (int b = 0; b < 7; ++b) {
            const bool in = n8 + b < n;
            tx[b] = in ? __ldg(hq + n8 + b) : 0;
            ty[b] = in ? __ldg(row + n8 + b) : 0;
        }
#pragma unroll
        for (int b = 0; b < 7; ++b)
            if (n8 + b < n) r = __dadd_rn(r, chi2_term(tx[b], ty[b]));
    }
    return r;
}

// k <= 128 words (the reference's 50-bin histograms): the whole histogram is one leaf -- one 8-lane group per stored
// histogram, 16 stored histograms per block.
__global__ void __launch_bounds__(128) chi2_scan_kernel(const int *hq, const int *db, long long n_db, int k, double *dist)
{
    const int j = threadIdx.x & 7;
    const unsigned group_mask = 0xFFu << (threadIdx.x & 24);
    const long long i = (long long)blockIdx.x * 16 + (threadIdx.x >> 3);
    if (i >= n_db) return;                               // whole groups leave
    const double r = chi2_leaf8(hq, db + i * k, k, j, group_mask);
    if (j == 0) dist[i] = r;
}

// k > 128 words (config 4's vocabulary has 65 536): ONE BLOCK per stored histogram.  The leaves (offset, length) and the
// inner nodes of the recursion  sum(n) = sum(n2) + sum(n - n2)  are laid out on the host for this k: value slots
// [0, n_leaves) are the leaf sums, slot n_leaves + i is inner node i = slot[left] + slot[right] (left operand first, as the
// recursion adds them), inner nodes sorted by height so that every level only reads finished slots.  The block's sixteen
// 8-lane groups sum the leaves, then the levels are added in parallel (9 levels for 65 536 words; a single thread walking
// the tree was most of the first wide kernel's time).  Bit-identical with np.sum.

__global__ void __launch_bounds__(128) chi2_scan_wide_kernel(const int *hq, const int *db, long long n_db, int k,
                                                             const int2 *leaves, int n_leaves, const int2 *nodes,
                                                             Chi2Levels lv, double *dist)
{
    extern __shared__ double slot[];
    const int j = threadIdx.x & 7, group = threadIdx.x >> 3;
    const unsigned group_mask = 0xFFu << (threadIdx.x & 24);
    const int *row = db + (long long)blockIdx.x * k;
    for (int l = group; l < n_leaves; l += 16) {
        const int2 lf = leaves[l];
        const double r = chi2_leaf8(hq + lf.x, row + lf.x, lf.y, j, group_mask);
        if (j == 0) slot[l] = r;
    }
    for (int h = 0; h < lv.n_levels; ++h) {
        __syncthreads();
        for (int i = lv.start[h] + threadIdx.x; i < lv.start[h + 1]; i += blockDim.x) {
            const int2 c = nodes[i];
            slot[n_leaves + i] = __dadd_rn(slot[c.x], slot[c.y]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) dist[blockIdx.x] = slot[n_leaves + lv.start[lv.n_levels] - 1];      // the root is the last inner node
}

// np.argmin / np.min: first index of the smallest value.  One CTA.
__global__ void argmin_kernel(const double *dist, long long n, int *best_idx, double *best_val)
{
    __shared__ double sv[32];
    __shared__ long long si[32];
    double v = 1.0 / 0.0;
    long long ix = -1;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = dist[i];
        if (ix < 0 || d < v) { v = d; ix = i; }      // strided, increasing i: keeps the first minimum per thread
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        const long long oi = __shfl_xor_sync(0xFFFFFFFFu, ix, o);
        if (oi >= 0 && (ix < 0 || ov < v || (ov == v && oi < ix))) { v = ov; ix = oi; }
    }
    if (lane == 0) { sv[warp] = v; si[warp] = ix; }
    __syncthreads();
    if (warp == 0) {
        v = lane < (int)(blockDim.x >> 5) ? sv[lane] : 1.0 / 0.0;
        ix = lane < (int)(blockDim.x >> 5) ? si[lane] : -1;
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xFFFFFFFFu, v, o);
            const long long oi = __shfl_xor_sync(0xFFFFFFFFu, ix, o);
            if (oi >= 0 && (ix < 0 || ov < v || (ov == v && oi < ix))) { v = ov; ix = oi; }
        }
        if (lane == 0) { *best_idx = (int)ix; *best_val = v; }
    }
}

}  // namespace

int slm_bow_hist_impl(slm_ctx *ctx, const int32_t *idx, int64_t n, int32_t idx_stride, int32_t n_words, int32_t *hist,
                      cudaStream_t stream)
{
    SLM_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_words * sizeof(int32_t), stream));
    if (n <= 0) return SLM_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 4LL * ctx->sm_count) blocks = 4LL * ctx->sm_count;
    const size_t smem = n_words <= 8192 ? (size_t)n_words * sizeof(int) : 0;
    bow_hist_kernel<<<(unsigned)blocks, 256, smem, stream>>>(idx, n, idx_stride, n_words, hist);
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SLM_OK;
}

int slm_chi2_scan_impl(slm_ctx *ctx, const int32_t *hq, const int32_t *db, int64_t n_db, int32_t k, double *dist,
                       int32_t *best_idx, double *best_val, cudaStream_t stream)
{
    if (n_db <= 0) return SLM_OK;
    if (k > kChi2LeafWords) {
        // leaves and inner nodes of numpy's pairwise-sum tree for this k (chi2_plan.h; cached on the ctx while k stays the same)
        if (ctx->chi2_leaves_k != k) {
            const Chi2Plan plan = chi2_plan(k);
            if (!plan.ok) return slm_fail(SLM_ERR_UNSUPPORTED, "chi-square scan: vocabulary too large (%d words)", k);
            SLM_TRY(slm_buf_reserve(ctx, &ctx->chi2_leaves, plan.table.size() * sizeof(Chi2Pair)));
            // pageable source: the driver has staged it when the call returns, so `plan` may go out of scope
            SLM_CUDA(cudaMemcpyAsync(ctx->chi2_leaves.p, plan.table.data(), plan.table.size() * sizeof(Chi2Pair),
                                     cudaMemcpyHostToDevice, stream));
            ctx->chi2_leaves_k = k;
            ctx->chi2_n_leaves = plan.n_leaves;
            ctx->chi2_n_inner = plan.n_inner;
            static_assert(sizeof(ctx->chi2_levels) == sizeof(Chi2Levels), "Chi2Levels mirror");
            static_assert(sizeof(Chi2Pair) == sizeof(int2), "Chi2Pair mirrors int2");
            memcpy(ctx->chi2_levels, &plan.levels, sizeof(Chi2Levels));
        }
        Chi2Levels lv;
        memcpy(&lv, ctx->chi2_levels, sizeof(lv));
        const size_t smem = (size_t)(ctx->chi2_n_leaves + ctx->chi2_n_inner) * sizeof(double);
        static std::atomic<bool> configured[64];
        if (smem > 48 * 1024 && !configured[ctx->device & 63].load()) {
            SLM_CUDA(cudaFuncSetAttribute(chi2_scan_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            configured[ctx->device & 63].store(true);
        }
        const int2 *table_dev = reinterpret_cast<const int2 *>(ctx->chi2_leaves.p);
        chi2_scan_wide_kernel<<<(unsigned)n_db, 128, smem, stream>>>(hq, db, n_db, k, table_dev, ctx->chi2_n_leaves,
                                                                    table_dev + ctx->chi2_n_leaves, lv, dist);
    } else {
        chi2_scan_kernel<<<(unsigned)((n_db + 15) / 16), 128, 0, stream>>>(hq, db, n_db, k, dist);
    }
    SLM_CUDA(cudaGetLastError());
    argmin_kernel<<<1, 1024, 0, stream>>>(dist, n_db, best_idx, best_val);
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return SLM_OK;
}
