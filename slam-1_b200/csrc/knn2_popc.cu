// knn2_popc.cu -- variant P: exact Hamming 2-NN with LOP3(XOR) + POPC on the integer pipe.
//
// Replaces the O(nq*nt) distance evaluation inside matcher.knnMatch(des1, des2, k=2)
// (reference call sites tracking.py:22, keypoint.py:44, Point3D.py:40; semantics = the exhaustive
// cv2.BFMatcher(NORM_HAMMING) form).
//
// Layout / mapping
//   * Each thread keeps kQPT query descriptors (8 x uint32 each) in registers for the whole kernel.
//   * The train set is split in `n_splits` contiguous row ranges (grid.y) so that small query sets still
//     fill 148 SMs; every CTA streams its range through shared memory in kTileRows-row tiles,
//     double-buffered with 16-byte cp.async (coalesced: consecutive threads copy consecutive 16 B).
//   * Inner loop: all lanes read the same train row (two broadcast LDS.128), 8 XOR + 8 POPC per
//     query, then a branch-free top-2 update on a packed 32-bit key (distance << 23 | local row):
//     b2 = min(b2, max(b1, key)); b1 = min(b1, key).  Unsigned key order == (distance, row), so the
//     lowest train index wins ties exactly as OpenCV's brute-force matcher does.
//   * Per-split partial keys go to workspace; a second tiny kernel merges the splits into the 64-bit
//     (distance << 32 | global index) keys shared by all variants.
//
// Roofline: 8 POPC32 per comparison on the quarter-rate pipe bounds this variant.
#include "slm_internal.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kQPT = 4;                  // queries per thread
constexpr int kQueriesPerCta = kThreads * kQPT;
constexpr int kTileRows = 128;           // train rows per shared-memory stage (4 KB)
constexpr int kIdxBits = 23;             // local row index bits in the packed 32-bit key
constexpr unsigned kLocalNone = 0xFFFFFFFFu;

struct PopcParams {
    const uint32_t *q;          // single problem: queries
    const uint32_t *t;          // single problem: train rows
    const uint32_t *desc;       // batched: uint32[n_frames][n_per_frame][8] (else nullptr)
    const int32_t *pairs;       // batched: device int32[n_prob][2] = (query frame, train frame)
    long long frame_words;      // batched: n_per_frame * 8
    int nq, nt;
    int rows_per_split, n_splits;
    unsigned *part;             // uint32[n_prob][n_splits][nq][2]
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__global__ void __launch_bounds__(kThreads) knn2_popc_kernel(PopcParams p)
{
    __shared__ __align__(16) uint4 tile[2][kTileRows * 2];

    const uint32_t *q = p.q;
    const uint32_t *t = p.t;
    if (p.desc != nullptr) {
        int2 pr = reinterpret_cast<const int2 *>(p.pairs)[blockIdx.z];
        q = p.desc + (long long)pr.x * p.frame_words;
        t = p.desc + (long long)pr.y * p.frame_words;
    }

    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * kQueriesPerCta;
    const int split = blockIdx.y;
    const int row_begin = split * p.rows_per_split;
    const int row_end = min(p.nt, row_begin + p.rows_per_split);
    const int n_rows = row_end - row_begin;  // > 0 by construction

    // register-resident queries (rows past nq are clamped; their results are never stored)
    uint32_t qr[kQPT][8];
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        int qi = min(q0 + k * kThreads + tid, p.nq - 1);
        const uint4 *src = reinterpret_cast<const uint4 *>(q + (long long)qi * 8);
        uint4 lo = __ldg(src), hi = __ldg(src + 1);
        qr[k][0] = lo.x; qr[k][1] = lo.y; qr[k][2] = lo.z; qr[k][3] = lo.w;
        qr[k][4] = hi.x; qr[k][5] = hi.y; qr[k][6] = hi.z; qr[k][7] = hi.w;
    }
    unsigned b1[kQPT], b2[kQPT];
#pragma unroll
    for (int k = 0; k < kQPT; ++k) { b1[k] = kLocalNone; b2[k] = kLocalNone; }

    const uint4 *tsrc = reinterpret_cast<const uint4 *>(t) + (long long)row_begin * 2;
    const int n_tiles = (n_rows + kTileRows - 1) / kTileRows;
    const int last_piece = n_rows * 2 - 1;

    auto load_tile = [&](int it, int buf) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            int piece = it * (kTileRows * 2) + k * kThreads + tid;
            cp_async16(&tile[buf][k * kThreads + tid], tsrc + min(piece, last_piece));
        }
        cp_async_commit();
    };

    load_tile(0, 0);
    for (int it = 0; it < n_tiles; ++it) {
        if (it + 1 < n_tiles) {
            load_tile(it + 1, (it + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint4 *s = tile[it & 1];
        const int rows = min(kTileRows, n_rows - it * kTileRows);
        const unsigned j0 = (unsigned)(it * kTileRows);
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            uint4 a = s[2 * r], b = s[2 * r + 1];
#pragma unroll
            for (int k = 0; k < kQPT; ++k) {
                unsigned d = __popc(qr[k][0] ^ a.x) + __popc(qr[k][1] ^ a.y) + __popc(qr[k][2] ^ a.z) +
                             __popc(qr[k][3] ^ a.w) + __popc(qr[k][4] ^ b.x) + __popc(qr[k][5] ^ b.y) +
                             __popc(qr[k][6] ^ b.z) + __popc(qr[k][7] ^ b.w);
                unsigned key = (d << kIdxBits) + (j0 + (unsigned)r);
                unsigned m = max(b1[k], key);
                b1[k] = min(b1[k], key);
                b2[k] = min(b2[k], m);
            }
        }
        __syncthreads();
    }

    unsigned *part = p.part + ((long long)blockIdx.z * p.n_splits + split) * (long long)p.nq * 2;
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        int qi = q0 + k * kThreads + tid;
        if (qi < p.nq) reinterpret_cast<uint2 *>(part)[qi] = make_uint2(b1[k], b2[k]);
    }
}


// =====================================================================================================
// Variant B: b1 AND.POPC mma.sync tiles (north_star part (2), kept for the A/B against LOP3+POPC).
//   Hamming(a, b) = popc(a) + popc(b) - 2 * popc(a & b);  popc(a & b) comes from
//   mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc.
// On sm_100a ptxas has no native binary MMA: every such instruction is lowered to 8 IMMA.16832 plus ~100
// logic / move instructions that expand bits to bytes per instruction, so this arm measures
// "legacy int8 tensor path behind a per-MMA bit expansion".  Same partial-key format and merge kernel as
// variant P.  Each warp owns 32 queries (two m16 tiles, A fragments register-resident); the 4 lanes of a
// quad hold different train columns, so the per-row top-2 is finished with two quad shuffles.
// =====================================================================================================
constexpr int kBmmaQueriesPerCta = 128;      // 4 warps x 32 queries
constexpr int kBmmaRowWords = 12;            // 48-byte padded train rows: conflict-free B-fragment loads

__device__ __forceinline__ void bmma_and_popc(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc "
                 "{%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kThreads) knn2_bmma_kernel(PopcParams p)
{
    __shared__ __align__(16) uint32_t tile[2][kTileRows * kBmmaRowWords];
    __shared__ int tile_popc[2][kTileRows];

    const uint32_t *q = p.q;
    const uint32_t *t = p.t;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
    const int q0 = blockIdx.x * kBmmaQueriesPerCta + warp * 32;
    const int split = blockIdx.y;
    const int row_begin = split * p.rows_per_split;
    const int row_end = min(p.nt, row_begin + p.rows_per_split);
    const int n_rows = row_end - row_begin;

    // A fragments: tile m covers query rows q0 + 16 m + {g, g + 8}; register i holds words {tg, 4 + tg}
    uint32_t a[2][4];
    int pa[2][2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int qi = min(q0 + 16 * m + g + 8 * h, p.nq - 1);
            const uint32_t *row = q + (long long)qi * 8;
            a[m][h] = __ldg(row + tg);
            a[m][2 + h] = __ldg(row + 4 + tg);
            int pc = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) pc += __popc(__ldg(row + w));
            pa[m][h] = pc;
        }
    }
    unsigned b1[2][2], b2[2][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) { b1[m][h] = kLocalNone; b2[m][h] = kLocalNone; }

    const uint4 *tsrc = reinterpret_cast<const uint4 *>(t) + (long long)row_begin * 2;
    const int n_tiles = (n_rows + kTileRows - 1) / kTileRows;
    const int last_row = n_rows - 1;
    auto load_tile = [&](int it, int buf) {
        // thread r copies train row r of the tile (two 16-byte pieces) and records its popcount
        const int r = min(it * kTileRows + tid, last_row);
        const uint4 lo = __ldg(tsrc + 2 * r), hi = __ldg(tsrc + 2 * r + 1);
        uint4 *dst = reinterpret_cast<uint4 *>(&tile[buf][tid * kBmmaRowWords]);
        dst[0] = lo;
        dst[1] = hi;
        tile_popc[buf][tid] = __popc(lo.x) + __popc(lo.y) + __popc(lo.z) + __popc(lo.w) + __popc(hi.x) + __popc(hi.y) +
                              __popc(hi.z) + __popc(hi.w);
    };

    load_tile(0, 0);
    __syncthreads();
    for (int it = 0; it < n_tiles; ++it) {
        const int buf = it & 1;
        if (it + 1 < n_tiles) load_tile(it + 1, buf ^ 1);
        const int rows = min(kTileRows, n_rows - it * kTileRows);
        const unsigned j0 = (unsigned)(it * kTileRows);
        for (int n0 = 0; n0 < rows; n0 += 8) {
            // B fragment: train column n0 + g, words {tg, 4 + tg}
            const uint32_t *brow = &tile[buf][(n0 + g) * kBmmaRowWords];
            const uint32_t bf0 = brow[tg], bf1 = brow[4 + tg];
            const int2 pb = *reinterpret_cast<const int2 *>(&tile_popc[buf][n0 + 2 * tg]);
            const bool v0 = n0 + 2 * tg < rows, v1 = n0 + 2 * tg + 1 < rows;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                int c[4] = {0, 0, 0, 0};
                bmma_and_popc(c, a[m], bf0, bf1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const unsigned d = (unsigned)(pa[m][h] + (e ? pb.y : pb.x) - 2 * c[2 * h + e]);
                        unsigned key = (d << kIdxBits) + (j0 + (unsigned)(n0 + 2 * tg + e));
                        key = (e ? v1 : v0) ? key : kLocalNone;
                        const unsigned mx = max(b1[m][h], key);
                        b1[m][h] = min(b1[m][h], key);
                        b2[m][h] = min(b2[m][h], mx);
                    }
                }
            }
        }
        __syncthreads();
    }
    // finish the per-row top-2 across the 4 lanes of a quad
    unsigned *part = p.part + ((long long)blockIdx.z * p.n_splits + split) * (long long)p.nq * 2;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            unsigned k1 = b1[m][h], k2 = b2[m][h];
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const unsigned o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, k2, o);
                const unsigned mx = max(k1, o1);
                k1 = min(k1, o1);
                k2 = min(min(k2, o2), mx);
            }
            const int qi = q0 + 16 * m + g + 8 * h;
            if (tg == 0 && qi < p.nq) reinterpret_cast<uint2 *>(part)[qi] = make_uint2(k1, k2);
        }
}

// Merge the per-split 32-bit local keys into 64-bit global keys.  One thread per (problem, query).
__global__ void popc_merge_kernel(const unsigned *part, int n_prob, int nq, int n_splits, int rows_per_split,
                                  long long base, unsigned long long *keys_out)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_prob * nq) return;
    int prob = (int)(gid / nq), i = (int)(gid % nq);
    unsigned long long k1 = kKeyNone, k2 = kKeyNone;
    const uint2 *pp = reinterpret_cast<const uint2 *>(part) + (long long)prob * n_splits * nq + i;
    for (int s = 0; s < n_splits; ++s) {
        uint2 v = pp[(long long)s * nq];
        unsigned long long off = (unsigned long long)(base + (long long)s * rows_per_split);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            unsigned lk = c == 0 ? v.x : v.y;
            if (lk == kLocalNone) continue;
            unsigned long long key = ((unsigned long long)(lk >> kIdxBits) << 32) |
                                     (off + (unsigned long long)(lk & ((1u << kIdxBits) - 1u)));
            unsigned long long m = max(k1, key);
            k1 = min(k1, key);
            k2 = min(k2, m);
        }
    }
    reinterpret_cast<ulonglong2 *>(keys_out)[gid] = make_ulonglong2(k1, k2);
}

int popc_launch(slm_ctx *ctx, PopcParams p, int n_prob, long long base, uint64_t *keys_out, cudaStream_t stream,
                bool bmma = false)
{
    ctx->last_variant = bmma ? SLM_VARIANT_BMMA : SLM_VARIANT_POPC;
    ctx->last_kernel = bmma ? "knn2_bmma_kernel" : "knn2_popc_kernel";
    const int per_cta = bmma ? kBmmaQueriesPerCta : kQueriesPerCta;
    const int qblocks = (p.nq + per_cta - 1) / per_cta;
    // enough CTAs for ~8 resident per SM; never split below one tile; local index must fit kIdxBits
    long long target = (long long)ctx->sm_count * 8;
    long long splits = (target + (long long)qblocks * n_prob - 1) / ((long long)qblocks * n_prob);
    long long max_splits = (p.nt + kTileRows - 1) / kTileRows;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long rps = (p.nt + splits - 1) / splits;
    rps = (rps + kTileRows - 1) / kTileRows * kTileRows;
    const long long max_rps = (1ll << kIdxBits);
    if (rps > max_rps) rps = max_rps;
    splits = (p.nt + rps - 1) / rps;
    if (splits > 65535) return slm_fail(SLM_ERR_UNSUPPORTED, "train set too large for one call (%d rows)", p.nt);
    p.rows_per_split = (int)rps;
    p.n_splits = (int)splits;

    size_t part_bytes = (size_t)n_prob * (size_t)splits * (size_t)p.nq * 2 * sizeof(unsigned);
    SLM_TRY(slm_buf_reserve(ctx, &ctx->scratch, part_bytes));
    p.part = reinterpret_cast<unsigned *>(ctx->scratch.p);

    dim3 grid(qblocks, (unsigned)splits, n_prob);
    SLM_TRY(slm_prof_begin(ctx, stream));
    if (bmma)
        knn2_bmma_kernel<<<grid, kThreads, 0, stream>>>(p);
    else
        knn2_popc_kernel<<<grid, kThreads, 0, stream>>>(p);
    SLM_CUDA(cudaGetLastError());
    SLM_TRY(slm_prof_end(ctx, stream));
    long long total = (long long)n_prob * p.nq;
    popc_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        p.part, n_prob, p.nq, p.n_splits, p.rows_per_split, base,
        reinterpret_cast<unsigned long long *>(keys_out));
    SLM_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return SLM_OK;
}

}  // namespace

int slm_popc_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt,
                       int64_t base, uint64_t *keys_out, cudaStream_t stream)
{
    PopcParams p{};
    p.q = q; p.t = t; p.desc = nullptr; p.pairs = nullptr; p.frame_words = 0;
    p.nq = (int)nq; p.nt = (int)nt;
    return popc_launch(ctx, p, 1, base, keys_out, stream);
}

int slm_bmma_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                       uint64_t *keys_out, cudaStream_t stream)
{
    PopcParams p{};
    p.q = q; p.t = t; p.desc = nullptr; p.pairs = nullptr; p.frame_words = 0;
    p.nq = (int)nq; p.nt = (int)nt;
    return popc_launch(ctx, p, 1, base, keys_out, stream, true);
}

int slm_popc_knn2_keys_batched(slm_ctx *ctx, const uint32_t *desc, int64_t n_per_frame,
                               const int32_t *pairs_dev, int64_t n_pairs, uint64_t *keys_out,
                               cudaStream_t stream)
{
    if (n_pairs > 65535) return slm_fail(SLM_ERR_UNSUPPORTED, "at most 65535 pairs per batched call");
    PopcParams p{};
    p.q = nullptr; p.t = nullptr; p.desc = desc; p.pairs = pairs_dev;
    p.frame_words = n_per_frame * 8;
    p.nq = (int)n_per_frame; p.nt = (int)n_per_frame;
    return popc_launch(ctx, p, (int)n_pairs, 0, keys_out, stream);
}
