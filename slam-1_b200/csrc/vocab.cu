// vocab.cu -- binary vocabulary training, the UPDATE half of one Lloyd iteration in Hamming space (k-majority).
//
// Reference: bag_of_words.py:14,20 -- `KMeans(n_clusters).fit(dpool)` clusters the pooled ORB descriptors as float
// vectors (and cannot be constructed on a current sklearn: n_jobs).  BASELINE config 4 assumes a
// BINARY vocabulary (uint8[k][32]) searched with the Hamming kernel, so training is re-specified the same way
// (SURVEY.md section 8(f) rank 3): assignment = the kNN kernel (word = nearest vocabulary row, lowest index on
// ties), update = per-word bitwise majority vote of its members:
//     bit b of word w  <-  1 if 2 * |{members with bit b set}| > |members|,
//                          unchanged if the vote is tied or the word has no members.
// The vote is a sum, so the result does not depend on the order members are visited in: the scatter below uses
// atomics for placement, yet the new vocabulary is bit-exact and deterministic.
//
// Pipeline (4 launches): word histogram (bow.cu) -> exclusive scan -> scatter of descriptor rows into per-word
// segments -> one CTA per word accumulates 256 bit counters over its segment and rewrites the centroid.
#include "slm_internal.cuh"

namespace {

constexpr int kScanThreads = 1024;

// offsets[w] = sum of counts[0..w); offsets[n_words] = total; cursor[w] = 0.  One CTA.
__global__ void __launch_bounds__(kScanThreads) vocab_scan_kernel(const int *counts, int n_words, int *offsets, int *cursor)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int w0 = 0; w0 < n_words; w0 += kScanThreads) {
        const int w = w0 + tid;
        const int c = w < n_words ? counts[w] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int s = warp_sums[lane];
            int si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, si, o);
                if (lane >= o) si += v;
            }
            warp_sums[lane] = si - s;
        }
        __syncthreads();
        const int excl = carry + warp_sums[warp] + incl - c;
        if (w < n_words) { offsets[w] = excl; cursor[w] = 0; }
        __syncthreads();
        if (tid == kScanThreads - 1) carry = excl + c;
        __syncthreads();
    }
    if (tid == 0) offsets[n_words] = carry;
}

__global__ void vocab_scatter_kernel(const int *words, long long n, int stride, int n_words, const int *offsets,
                                     int *cursor, int *order)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int w = words[i * stride];
        if (w >= 0 && w < n_words) order[offsets[w] + atomicAdd(&cursor[w], 1)] = (int)i;
    }
}

constexpr int kMajThreads = 128;

// One CTA per word.  Lane l of every warp owns descriptor word (l & 7), bits 8*(l >> 3) .. +7; the warps stride over
// the word's members.  256 counters in shared memory, then threads 0..7 rebuild the eight 32-bit centroid words.
__global__ void __launch_bounds__(kMajThreads) vocab_majority_kernel(const uint32_t *desc, const int *order, const int *offsets,
                                                                     uint32_t *vocab, int *changed)
{
    __shared__ int cnt[256];
    const int w = blockIdx.x;
    const int begin = offsets[w], end = offsets[w + 1], members = end - begin;
    if (members == 0) return;                     // empty word: the centroid stays
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cnt[tid] = 0;
    cnt[tid + kMajThreads] = 0;
    __syncthreads();
    const int wi = lane & 7, sh = 8 * (lane >> 3);
    int c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = 0;
#pragma unroll 4
    for (int k = begin + warp; k < end; k += kMajThreads / 32) {
        const unsigned byte = (__ldg(desc + (long long)order[k] * 8 + wi) >> sh) & 0xFFu;
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] += (byte >> j) & 1u;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&cnt[wi * 32 + sh + j], c[j]);
    __syncthreads();
    bool diff = false;
    if (tid < 8) {
        const uint32_t old = vocab[(long long)w * 8 + tid];
        uint32_t nw = 0;
        for (int b = 0; b < 32; ++b) {
            const int twice = 2 * cnt[tid * 32 + b];
            const uint32_t bit = twice > members ? 1u : (twice == members ? (old >> b) & 1u : 0u);
            nw |= bit << b;
        }
        if (nw != old) {
            vocab[(long long)w * 8 + tid] = nw;
            diff = true;
        }
    }
    if (warp == 0) {
        const unsigned any = __ballot_sync(0xFFFFFFFFu, diff);
        if (lane == 0 && any && changed) atomicAdd(changed, 1);
    }
}

}  // namespace

int slm_vocab_update_impl(slm_ctx *ctx, const uint32_t *desc, int64_t n, const int32_t *words, int32_t stride,
                          uint32_t *vocab, int32_t n_words, int32_t *counts_out, int32_t *changed_out, cudaStream_t stream)
{
    // workspace: counts[n_words] | offsets[n_words + 1] | cursor[n_words] | order[n]
    const size_t ints = (size_t)3 * n_words + 1 + (size_t)n;
    SLM_TRY(slm_buf_reserve(ctx, &ctx->misc, ints * sizeof(int)));
    int *counts = reinterpret_cast<int *>(ctx->misc.p);
    int *offsets = counts + n_words;
    int *cursor = offsets + n_words + 1;
    int *order = cursor + n_words;
    if (changed_out) SLM_CUDA(cudaMemsetAsync(changed_out, 0, sizeof(int32_t), stream));
    SLM_TRY(slm_bow_hist_impl(ctx, words, n, stride, n_words, counts, stream));
    if (counts_out)
        SLM_CUDA(cudaMemcpyAsync(counts_out, counts, (size_t)n_words * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    if (n <= 0) return SLM_OK;
    vocab_scan_kernel<<<1, kScanThreads, 0, stream>>>(counts, n_words, offsets, cursor);
    SLM_CUDA(cudaGetLastError());
    long long blocks = (n + 255) / 256;
    if (blocks > 8LL * ctx->sm_count) blocks = 8LL * ctx->sm_count;
    vocab_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(words, n, stride, n_words, offsets, cursor, order);
    SLM_CUDA(cudaGetLastError());
    vocab_majority_kernel<<<(unsigned)n_words, kMajThreads, 0, stream>>>(desc, order, offsets, vocab, changed_out);
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 3;
    return SLM_OK;
}
