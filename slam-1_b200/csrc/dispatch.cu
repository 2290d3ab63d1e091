// dispatch.cu -- shape policy for SLM_VARIANT_AUTO and the batched (config 3) entry point.
#include "slm_internal.cuh"

static constexpr int64_t kAutoTensorMinCmp = 1ll << 18;

int slm_auto_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt,
                       int64_t base, uint64_t *keys_out, cudaStream_t stream)
{
    // The tensor-pipe variant is ~9x faster at scale; the integer-pipe variant has the smaller
    // fixed cost, which wins only for tiny problems.
    // a handful of queries against a long train set is HBM-bound: stream the train rows once
    // frame-to-frame shapes: one launch instead of three
    if (slm_frame_eligible(ctx, nq, nt, false))
        return slm_frame_knn2(ctx, q, nq, t, nt, base, 0, 1, 0, keys_out, nullptr, nullptr, nullptr, stream);
    if (nq <= 8) return slm_stream_knn2_keys(ctx, q, nq, t, nt, base, keys_out, stream);
    if (nq * nt < kAutoTensorMinCmp) return slm_popc_knn2_keys(ctx, q, nq, t, nt, base, keys_out, stream);
    return slm_tc_knn2_keys(ctx, q, nq, t, nt, base, keys_out, stream, ctx->tc_fp4 != 0);
}

int slm_batched_knn2_keys(slm_ctx *ctx, const uint32_t *desc, int64_t n_per_frame, const int32_t *pairs_dev,
                          int64_t n_pairs, uint64_t *keys_out, cudaStream_t stream, const slm_chain *chain)
{
    // grid.y / grid.z carry the pair index: chunk long pair lists
    const int64_t kMaxPairs = 32768;
    if (n_pairs > kMaxPairs) chain = nullptr;     // chain units index the whole (sorted) pair list
    for (int64_t p0 = 0; p0 < n_pairs; p0 += kMaxPairs) {
        int64_t n = n_pairs - p0 < kMaxPairs ? n_pairs - p0 : kMaxPairs;
        if (ctx->variant == SLM_VARIANT_TENSOR || ctx->variant == SLM_VARIANT_TENSOR4 ||
            (ctx->variant == SLM_VARIANT_AUTO && n_per_frame * n_per_frame >= kAutoTensorMinCmp))
            SLM_TRY(slm_tc_knn2_keys_batched(ctx, desc, n_per_frame, pairs_dev + 2 * p0, n,
                                             keys_out + 2 * p0 * n_per_frame, stream, chain,
                                             ctx->variant == SLM_VARIANT_TENSOR4 || (ctx->variant == SLM_VARIANT_AUTO && ctx->tc_fp4)));
        else
            SLM_TRY(slm_popc_knn2_keys_batched(ctx, desc, n_per_frame, pairs_dev + 2 * p0, n,
                                               keys_out + 2 * p0 * n_per_frame, stream));
    }
    return SLM_OK;
}

