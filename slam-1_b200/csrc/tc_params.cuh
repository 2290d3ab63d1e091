// tc_params.cuh -- launch parameters and candidate-key constants shared by the tensor-pipe kernels
// (knn2_tc.cu: fp8 kind::f8f6f4, knn2_tc4.cu: fp4 kind::mxf4) and their refine kernels.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace tcp {

constexpr int kTileM = 128;                        // query rows per MMA half (one CTA)
constexpr int kChunkBits = 15;                     // chunk counter bits in the packed fp32 candidate key
constexpr int kChunkMask = (1 << kChunkBits) - 1;  // 32767
constexpr float kKeyScale = (float)(1 << kChunkBits);
constexpr int kKeyBias = 256 << kChunkBits;        // makes (int)key non-negative: |dot| <= 256

// Candidate key of a chunk of train rows, as the epilogues write it:
//   key = max dot of the chunk * 2^15 + (32767 - chunk counter inside the unit's candidate epoch)       (exact, < 2^24)
// so a plain fp32 max prefers the larger dot (smaller distance) and, on ties, the earlier chunk.

struct TcParams {
    const uint32_t *q, *t;        // single problem
    const uint32_t *desc;         // batched: uint32[n_frames][n_per_frame][8] (else nullptr)
    const int32_t *pairs;         // batched: device int32[n_prob][2]
    long long frame_words;
    int nq, nt;
    int n_prob;
    int mt;                       // query tiles per group
    int n_groups;                 // ceil(nq / (128 * mt))
    int tile_n;                   // train rows per tile: 256 (fp8 kernels) or 240 (fp4 kernel)
    int range_tiles, n_ranges;    // the train set is cut into n_ranges ranges of range_tiles tiles
    int cpg;                      // CTAs (1-CTA kernel) / clusters (2-CTA kernels) per query group (pair):
                                  // unit c walks ranges c, c + cpg, c + 2 cpg, ... of its group
    int chunk;                    // candidate chunk width in train rows (tile_n is a multiple of it)
    int rpe, n_epochs;            // ranges per candidate epoch, epochs per unit
    float2 *cand;                 // [n_prob][nq][cpg * n_epochs][2 sets]: best two chunk keys per epilogue set
    // Chained batches (2-CTA kernels, config 3): grid.y indexes UNITS = runs of pairs that share the query frame.  The
    // cluster expands that frame's query tiles once and walks the train frames of the run back to back (one
    // candidate flush per pair), instead of paying the cluster start-up once per pair.
    const int32_t *chain_pairs;   // device int32[n_prob][2], pairs sorted by query frame (nullptr = not chained)
    const int32_t *chain_prob;    // device int32[n_prob]: caller's pair index of every sorted pair (output slot)
    const int32_t *chain_units;   // device int32[n_units][2] = (first sorted pair, number of pairs)
    long long *timing;            // diagnostic build of the fp4 kernel only (SLM_TC4_TIMING): per-phase cycle sums
};

}  // namespace tcp

// fp4 kernel (knn2_tc4.cu): grid_y = problems (or chain units); p.mt query tiles per CTA (<= kMaxMT4), p.chunk in {120, 40, 24}
struct slm_ctx;
int slm_tc4_launch(slm_ctx *ctx, const tcp::TcParams &p, int grid_y, cudaStream_t stream);
constexpr int kMaxMT4 = 8;
