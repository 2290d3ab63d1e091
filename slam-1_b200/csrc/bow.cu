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
// and the sum follows numpy's pairwise_sum order (8 accumulators for n <= 128, recursive halves above).
#include <atomic>
#include <vector>
#include "slm_internal.cuh"

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
    const long long d = (long long)x - (long long)y;
    const long long num = 2 * d * d;                 // exact in int64, like numpy's integer arithmetic
    const long long den = max(1ll, (long long)x + (long long)y);
    return __ddiv_rn((double)num, (double)den);      // true divide: both converted to float64
}

// numpy's pairwise_sum for a contiguous float64 vector whose i-th element is chi2_term(hq[i], row[i]).
__device__ double pairwise_chi2(const int *hq, const int *row, int n)
{
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, chi2_term(hq[i], row[i]));
        return r;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = chi2_term(hq[j], row[j]);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], chi2_term(hq[i + j], row[i + j]));
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, chi2_term(hq[i], row[i]));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(pairwise_chi2(hq, row, n2), pairwise_chi2(hq + n2, row + n2, n - n2));
}

__global__ void chi2_scan_kernel(const int *hq, const int *db, long long n_db, int k, double *dist)
{
    extern __shared__ int shq[];
    for (int w = threadIdx.x; w < k; w += blockDim.x) shq[w] = hq[w];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_db) dist[i] = pairwise_chi2(shq, db + i * k, k);
}

// Larger vocabularies (k > kChi2SmemWords: one thread per stored histogram would walk a 4 * k-byte row alone, and the
// recursive pairwise_chi2 above would outgrow the 1 KB per-thread stack beyond 8192 words -- a latent fault of the first
// version of this scan, which accepted 12288): ONE BLOCK per stored histogram.  numpy's pairwise sum is a binary tree whose
// leaves are runs of <= 128 elements (8 accumulators each) -- the leaves (offset, length; computed on the host for this k)
// are summed by the threads in parallel, every thread reading whole 128-byte lines of its runs, and thread 0 then adds the
// leaf sums in the recursion's own order, so the result is still bit-identical with np.sum.
// The recursion  sum(n) = sum(n2) + sum(n - n2),  n2 = n / 2 rounded down to a multiple of 8,  over the leaf sums in leaf
// order -- written as a loop with an explicit stack (depth <= 14 for 2^20 words): device recursion would live on the
// 1 KB default per-thread stack.
__device__ double chi2_tree(int k, const double *leaf_sum)
{
    int n_of[24];
    double left_of[24];
    bool has_left[24];
    int sp = 0, next = 0;
    n_of[0] = k;
    has_left[0] = false;
    double ret = 0.0;
    bool returning = false;
    while (sp >= 0) {
        const int n = n_of[sp];
        int n2 = n / 2;
        n2 -= n2 % 8;
        if (!returning) {
            if (n <= 128) {                      // a leaf: return its sum to the parent
                ret = leaf_sum[next++];
                returning = true;
                --sp;
            } else {                             // descend into the left child
                has_left[sp] = false;
                n_of[++sp] = n2;
            }
        } else if (!has_left[sp]) {              // back from the left child: keep it, descend into the right child
            left_of[sp] = ret;
            has_left[sp] = true;
            returning = false;
            n_of[++sp] = n - n2;
        } else {                                 // back from the right child
            ret = __dadd_rn(left_of[sp], ret);
            --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(128) chi2_scan_wide_kernel(const int *hq, const int *db, long long n_db, int k,
                                                             const int2 *leaves, int n_leaves, double *dist)
{
    extern __shared__ double leaf_sum[];
    const int *row = db + (long long)blockIdx.x * k;
    for (int l = threadIdx.x; l < n_leaves; l += blockDim.x) {
        const int2 lf = leaves[l];
        leaf_sum[l] = pairwise_chi2(hq + lf.x, row + lf.x, lf.y);       // lf.y <= 128: the non-recursive branches
    }
    __syncthreads();
    if (threadIdx.x == 0) dist[blockIdx.x] = chi2_tree(k, leaf_sum);
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
    if (k > kChi2SmemWords) {
        // leaves of numpy's pairwise-sum tree for this k (cached on the ctx while k stays the same)
        if (ctx->chi2_leaves_k != k) {
            std::vector<int2> lv;
            std::vector<int2> stack{make_int2(0, k)};
            while (!stack.empty()) {                         // depth-first, left child first = the recursion's leaf order
                const int2 r = stack.back();
                stack.pop_back();
                if (r.y <= 128) { lv.push_back(r); continue; }
                int n2 = r.y / 2;
                n2 -= n2 % 8;
                stack.push_back(make_int2(r.x + n2, r.y - n2));
                stack.push_back(make_int2(r.x, n2));
            }
            SLM_TRY(slm_buf_reserve(ctx, &ctx->chi2_leaves, lv.size() * sizeof(int2)));
            // pageable source: the driver has staged it when the call returns, so `lv` may go out of scope
            SLM_CUDA(cudaMemcpyAsync(ctx->chi2_leaves.p, lv.data(), lv.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
            ctx->chi2_leaves_k = k;
            ctx->chi2_n_leaves = (int)lv.size();
        }
        const size_t smem = (size_t)ctx->chi2_n_leaves * sizeof(double);
        static std::atomic<bool> configured[64];
        if (smem > 48 * 1024 && !configured[ctx->device & 63].load()) {
            SLM_CUDA(cudaFuncSetAttribute(chi2_scan_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            configured[ctx->device & 63].store(true);
        }
        chi2_scan_wide_kernel<<<(unsigned)n_db, 128, smem, stream>>>(hq, db, n_db, k, reinterpret_cast<const int2 *>(ctx->chi2_leaves.p),
                                                                    ctx->chi2_n_leaves, dist);
    } else {
        chi2_scan_kernel<<<(unsigned)((n_db + 127) / 128), 128, (size_t)k * sizeof(int), stream>>>(hq, db, n_db, k, dist);
    }
    SLM_CUDA(cudaGetLastError());
    argmin_kernel<<<1, 1024, 0, stream>>>(dist, n_db, best_idx, best_val);
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return SLM_OK;
}
