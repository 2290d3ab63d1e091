// knn2_tc.cu -- variant T: exact Hamming 2-NN on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces the O(nq*nt) distance evaluation inside matcher.knnMatch(des1, des2, k=2)
// (reference call sites tracking.py:22, keypoint.py:44, Point3D.py:40; exhaustive semantics).
//
// Idea.  The all-pairs Hamming distance is a dense binary contraction.  Expanding every descriptor bit
// to fp8 +1.0 / -1.0 gives   dot(x_i, y_j) = 256 - 2 * Hamming(q_i, t_j)   exactly (small integers in an
// fp32 accumulator), so a 128 (queries) x 256 (train) x 256 (bits) "job" is 8 tcgen05.mma instructions.
// sm_100a has no native 1-bit tensor op (ptxas emulates mma.sync b1 with IMMA + bit twiddling,
// SURVEY.md H1); the expansion is instead done ONCE per operand tile, in shared memory.
//
// One CTA = one work item = (query group of up to 3 x 128 rows) x (contiguous train range), 1 CTA / SM:
//   warps 0-7   epilogue: TMEM -> registers (tcgen05.ld 32x32b.x32), branch-free candidate tracking
//   warps 8-15  expanders: packed bits (LDG.128, coalesced 32-B rows) -> +-1 fp8 rows in the K-major
//               no-swizzle UMMA layout (conflict-free STS.128), double-buffered train stages
//   warp  16    one elected thread issues tcgen05.mma (M=128, N=256, K=32) into two 256-column TMEM
//               accumulators and tcgen05.commit's onto mbarriers
// Pipelines (all mbarrier): a_full, b_full[2]/b_empty[2] (expanders <-> MMA), acc_full[2]/acc_empty[2]
// (MMA <-> epilogue).  HBM traffic is the packed descriptors only; the 8x expanded operands never
// leave the SM.
//
// Epilogue.  Thread = one query row (one TMEM lane); it sees its row's dot products 32 columns at a time.
// Tracking (distance, index) per element would cost more ALU than the MMA leaves room for, so the kernel
// keeps, per row, only the best two 32-column CHUNKS by (max dot desc, chunk index asc), packed into one
// fp32 (dot * 32768 + 32767 - chunk counter; exact, < 2^24) and updated with three FMNMX.  The exact top-2 rows
// (distance asc, index asc) of a range always lie inside its best two chunks, so a tiny second kernel
// re-scores just those candidate rows with XOR+POPC and applies the reference tie-break bit-exactly.
// Cost per element: ~0.5 FMNMX3; data-independent (no divergence on adversarial inputs).
#include <cfloat>
#include <atomic>
#include "slm_internal.cuh"
#include "exchange.cuh"
#include "tc_common.cuh"
#include "tc_params.cuh"

namespace {

using namespace tcp;

constexpr int kTileN = 256;
constexpr int kChunk = 32;
constexpr int kChunksPerTile = kTileN / kChunk;   // 8
constexpr int kMaxMT = 3;                          // query tiles resident per CTA
constexpr int kEpiWarps = 8;
constexpr int kExpWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;        // 256
constexpr int kExpThreads = kExpWarps * 32;        // 256
constexpr int kThreads = kEpiThreads + kExpThreads + 32;   // 544
constexpr uint32_t kATileBytes = kTileM * 256;     // 32 KB
constexpr uint32_t kBTileBytes = kTileN * 256;     // 64 KB
constexpr int kMaxEpochTiles = (1 << kChunkBits) / kChunksPerTile;   // 4096 tiles per candidate epoch

struct TcBarriers {
    uint64_t a_full;
    uint64_t b_full[2], b_empty[2];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float max32(const uint32_t (&v)[32])
{
    float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
    for (int j = 4; j < 32; j += 8) {   // 4 independent chains of 3-input max (FMNMX3)
        m0 = fmaxf(fmaxf(m0, __uint_as_float(v[j + 0])), __uint_as_float(v[j + 1]));
        m1 = fmaxf(fmaxf(m1, __uint_as_float(v[j + 2])), __uint_as_float(v[j + 3]));
        m2 = fmaxf(fmaxf(m2, __uint_as_float(v[j + 4])), __uint_as_float(v[j + 5]));
        m3 = fmaxf(fmaxf(m3, __uint_as_float(v[j + 6])), __uint_as_float(v[j + 7]));
    }
    m0 = fmaxf(fmaxf(m0, __uint_as_float(v[28])), __uint_as_float(v[29]));
    m1 = fmaxf(fmaxf(m1, __uint_as_float(v[30])), __uint_as_float(v[31]));
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// maximum of 16 of the 32 loaded columns (half-width candidate chunks of chained batches): 8 FMNMX3
template <int O>
__device__ __forceinline__ float max16(const uint32_t (&v)[32])
{
    float m0 = fmaxf(fmaxf(__uint_as_float(v[O + 0]), __uint_as_float(v[O + 1])), __uint_as_float(v[O + 2]));
    float m1 = fmaxf(fmaxf(__uint_as_float(v[O + 3]), __uint_as_float(v[O + 4])), __uint_as_float(v[O + 5]));
    float m2 = fmaxf(fmaxf(__uint_as_float(v[O + 6]), __uint_as_float(v[O + 7])), __uint_as_float(v[O + 8]));
    float m3 = fmaxf(fmaxf(__uint_as_float(v[O + 9]), __uint_as_float(v[O + 10])), __uint_as_float(v[O + 11]));
    m0 = fmaxf(fmaxf(m0, __uint_as_float(v[O + 12])), __uint_as_float(v[O + 13]));
    m1 = fmaxf(fmaxf(m1, __uint_as_float(v[O + 14])), __uint_as_float(v[O + 15]));
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

template <int MT>
__global__ void __launch_bounds__(kThreads, 1) knn2_tc_kernel(TcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                   // MT tiles of 32 KB
    uint8_t *sB = smem + MT * kATileBytes;                // 2 stages of 64 KB
    TcBarriers *bars = reinterpret_cast<TcBarriers *>(smem + MT * kATileBytes + 2 * kBTileBytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int range = blockIdx.x / p.n_groups;
    const int group = blockIdx.x % p.n_groups;

    const uint32_t *q = p.q;
    const uint32_t *t = p.t;
    if (p.desc != nullptr) {
        int2 pr = reinterpret_cast<const int2 *>(p.pairs)[blockIdx.y];
        q = p.desc + (long long)pr.x * p.frame_words;
        t = p.desc + (long long)pr.y * p.frame_words;
    }

    const int q_first = group * (MT * kTileM);
    const int mt_here = min(MT, (p.nq - q_first + kTileM - 1) / kTileM);     // >= 1
    const int col_first = range * p.range_tiles * kTileN;
    const int col_end = min(p.nt, col_first + p.range_tiles * kTileN);
    const int n_tiles = (col_end - col_first + kTileN - 1) / kTileN;         // >= 1

    if (tid == 0) {
        tc::mbar_init(&bars->a_full, kExpThreads);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&bars->b_full[s], kExpThreads);
            tc::mbar_init(&bars->b_empty[s], 1);
            tc::mbar_init(&bars->acc_full[s], 1);
            tc::mbar_init(&bars->acc_empty[s], kEpiThreads);
        }
        tc::fence_barrier_init();
    }
    if (warp == kEpiWarps + kExpWarps) tc::tmem_alloc(&bars->tmem_base, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp < kEpiWarps) {
        // ===================== epilogue: candidate chunks per query row =====================
        const int set = warp >> 2, quad = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        float b1[MT], b2[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) { b1[m] = -FLT_MAX; b2[m] = -FLT_MAX; }
        int job = 0;
        for (int bt = 0; bt < n_tiles; ++bt) {
            const int valid_cols = min(kTileN, col_end - (col_first + bt * kTileN));
            const int n_chunks = (valid_cols + kChunk - 1) / kChunk;
            const float chunk_bias = (float)(kChunkMask - bt * kChunksPerTile);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                if (m < mt_here) {
                    const int ab = job & 1;
                    tc::mbar_wait(&bars->acc_full[ab], (job >> 1) & 1, 10 + ab);
                    tc::tc_fence_after();
                    const uint32_t acc_addr = lane_addr + ab * kTileN + set * kChunk;
                    if (n_chunks == kChunksPerTile) {
                        // full tile: 4 chunks per warp, TMEM load of chunk i+1 in flight while chunk i is reduced
                        uint32_t va[32], vb[32];
                        tc::tmem_ld32_nowait(acc_addr, va);
#pragma unroll
                        for (int cc = 0; cc < kChunksPerTile / 2; ++cc) {
                            uint32_t (&cur)[32] = (cc & 1) ? vb : va;
                            uint32_t (&nxt)[32] = (cc & 1) ? va : vb;
                            tc::tmem_ld_wait(cur);
                            if (cc + 1 < kChunksPerTile / 2) {
                                tc::tmem_ld32_nowait(acc_addr + (cc + 1) * 2 * kChunk, nxt);
                            } else {
                                // every load of this accumulator has landed: hand it back before the last reduce
                                tc::tc_fence_before();
                                tc::mbar_arrive(&bars->acc_empty[ab]);
                            }
                            const float key = fmaf(max32(cur), kKeyScale, chunk_bias - (float)(2 * cc + set));
                            b2[m] = fmaxf(b2[m], fminf(b1[m], key));
                            b1[m] = fmaxf(b1[m], key);
                        }
                    } else {
                        // last, partial tile of the train set: only chunks that contain valid columns count
#pragma unroll
                        for (int cc = 0; cc < kChunksPerTile / 2; ++cc) {
                            const int c = 2 * cc + set;
                            if (c < n_chunks) {
                                uint32_t v[32];
                                tc::tmem_ld32(acc_addr + cc * 2 * kChunk, v);
                                const float key = fmaf(max32(v), kKeyScale, chunk_bias - (float)c);
                                b2[m] = fmaxf(b2[m], fminf(b1[m], key));
                                b1[m] = fmaxf(b1[m], key);
                            }
                        }
                        tc::tc_fence_before();
                        tc::mbar_arrive(&bars->acc_empty[ab]);
                    }
                    ++job;
                }
            }
        }
        // candidates of one query are contiguous: cand[prob][query][slot = range][set] (float2 = best two chunk keys)
        float2 *cand = p.cand + (long long)blockIdx.y * p.nq * p.n_ranges * 2 + (long long)range * 2 + set;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int qi = q_first + m * kTileM + quad * 32 + lane;
            if (m < mt_here && qi < p.nq) cand[(long long)qi * p.n_ranges * 2] = make_float2(b1[m], b2[m]);
        }
    } else if (warp < kEpiWarps + kExpWarps) {
        // ===================== expanders: packed bits -> +-1 fp8 operand tiles =====================
        const int et = tid - kEpiThreads;   // 0..255
        const uint32_t sA_addr = tc::smem_u32(sA), sB_addr = tc::smem_u32(sB);
        for (int r = et; r < mt_here * kTileM; r += kExpThreads) {
            const int qi = min(q_first + r, p.nq - 1);
            const uint4 *src = reinterpret_cast<const uint4 *>(q + (long long)qi * 8);
            uint4 d0 = __ldg(src), d1 = __ldg(src + 1);
            tc::expand_row_to_smem(sA_addr + (uint32_t)(r / kTileM) * kATileBytes, r % kTileM, d0, d1);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&bars->a_full);

        auto load_row = [&](int bt, uint4 &d0, uint4 &d1) {
            const int row = min(col_first + bt * kTileN + et, p.nt - 1);
            const uint4 *src = reinterpret_cast<const uint4 *>(t + (long long)row * 8);
            d0 = __ldg(src);
            d1 = __ldg(src + 1);
        };
        uint4 n0, n1;
        load_row(0, n0, n1);
        for (int bt = 0; bt < n_tiles; ++bt) {
            const int s = bt & 1;
            const uint4 c0 = n0, c1 = n1;
            if (bt + 1 < n_tiles) load_row(bt + 1, n0, n1);      // prefetch the next tile's row
            tc::mbar_wait(&bars->b_empty[s], ((bt >> 1) & 1) ^ 1, 20 + s);
            tc::expand_row_to_smem(sB_addr + (uint32_t)s * kBTileBytes, et, c0, c1);
            tc::fence_proxy_async();
            tc::mbar_arrive(&bars->b_full[s]);
        }
    } else {
        // ===================== MMA issuer: the warp stays converged, one elected lane issues =====================
        {
            const uint32_t idesc = tc::idesc_e4m3_f32(kTileM, kTileN);
            const uint32_t a_lo0 = tc::smem_desc_lo(tc::smem_u32(sA)), b_lo0 = tc::smem_desc_lo(tc::smem_u32(sB));
            tc::mbar_wait(&bars->a_full, 0, 30);
            tc::tc_fence_after();
            int job = 0;
            for (int bt = 0; bt < n_tiles; ++bt) {
                const int s = bt & 1;
                tc::mbar_wait(&bars->b_full[s], (bt >> 1) & 1, 31 + s);
                tc::tc_fence_after();
                for (int m = 0; m < mt_here; ++m) {
                    const int ab = job & 1;
                    tc::mbar_wait(&bars->acc_empty[ab], ((job >> 1) & 1) ^ 1, 33 + ab);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        tc::umma_job<1>(tmem + ab * kTileN, a_lo0 + m * (kATileBytes >> 4), b_lo0 + s * (kBTileBytes >> 4), idesc);
                        tc::umma_commit(&bars->acc_full[ab]);
                    }
                    __syncwarp();
                    ++job;
                }
                if (tc::elect_one()) tc::umma_commit(&bars->b_empty[s]);
                __syncwarp();
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + kExpWarps) tc::tmem_dealloc(tmem, 512);
}


// =====================================================================================================
// 2-CTA variant (cta_group::2).  A CTA pair (cluster of 2, same TPC) works on one train range; each CTA
// owns up to MT query tiles (its 128 rows of the M = 256 MMA) and expands / stores only HALF of every
// 256-row train tile (CTA rank r holds tile rows [128 r, 128 r + 128)).  The leader CTA's elected thread
// issues tcgen05.mma.cta_group::2; the hardware exchanges the B halves between the two SMs, so per SM the
// MMA reads 8 KB of shared memory per instruction instead of 12 KB and the expansion work per SM halves.
// (ncu on the 1-CTA kernel: tensor pipe 52 % active with the shared-memory data pipe ~85 % busy -- the
// SS-mode operand fetch of an fp8 K=256 job is shared-memory-bandwidth bound on one SM.)
// Barriers: "full" barriers live in the leader and collect arrivals from both CTAs (remote arrives);
// "empty"/"acc_full" barriers are signalled into both CTAs by multicast tcgen05.commit.
// =====================================================================================================
// 12 warps = 384 threads: register allocation is per 4 warps, so a 13th warp would cap the kernel at 128
// registers per thread; the 128 rows of a half tile are therefore spread over 3 expander warps.
constexpr int kExp2Warps = 3;
constexpr int kExp2Threads = kExp2Warps * 32;                 // 96
constexpr int kThreads2 = kEpiThreads + kExp2Threads + 32;    // 384
constexpr uint32_t kBHalfBytes = 128 * 256;                   // 32 KB
constexpr int kBStages2 = 3;
constexpr int kMaxMT2 = 4;

struct TcBarriers2 {
    uint64_t a_full;
    uint64_t b_full[kBStages2], b_empty[kBStages2];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

template <int MT, bool CHAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1) knn2_tc2_kernel(TcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                   // MT tiles of 32 KB
    uint8_t *sB = smem + MT * kATileBytes;                // kBStages2 half tiles of 32 KB
    TcBarriers2 *bars = reinterpret_cast<TcBarriers2 *>(smem + MT * kATileBytes + kBStages2 * kBHalfBytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = tc::cluster_ctarank();          // 0 = leader
    const int item = blockIdx.x >> 1;                     // cluster index
    const int gpair = item / p.cpg;                       // pair of query groups served by this cluster
    const int unit = item % p.cpg;                        // this cluster walks ranges unit, unit + cpg, ...
    const int group = gpair * 2 + (int)rank;              // may be == n_groups (idle half of an odd pair)

    const uint32_t *q = p.q;
    const uint32_t *t = p.t;
    int link_first = 0, n_links = 1;              // chained batch: sorted pairs [link_first, link_first + n_links)
    if constexpr (CHAIN) {
        const int2 cu = reinterpret_cast<const int2 *>(p.chain_units)[blockIdx.y];
        link_first = cu.x;
        n_links = cu.y;
        q = p.desc + (long long)p.chain_pairs[2 * link_first] * p.frame_words;
    } else if (p.desc != nullptr) {
        int2 pr = reinterpret_cast<const int2 *>(p.pairs)[blockIdx.y];
        q = p.desc + (long long)pr.x * p.frame_words;
        t = p.desc + (long long)pr.y * p.frame_words;
    }
    // train rows / output slot of link l (identity when not chained)
    auto link_train = [&](int l) -> const uint32_t * {
        if constexpr (CHAIN) return p.desc + (long long)p.chain_pairs[2 * (link_first + l) + 1] * p.frame_words;
        else return t;
    };
    auto link_prob = [&](int l) -> int {
        if constexpr (CHAIN) return p.chain_prob[link_first + l];
        else return (int)blockIdx.y;
    };
    if constexpr (!CHAIN) n_links = 1;

    const int q_first = group * (MT * kTileM);
    // query tiles this CTA really owns (0 for the idle half) and the number of MMA chains per train tile,
    // which is the pair's maximum: both CTAs must take part in every chain
    const int mt_mine = max(0, min(MT, (p.nq - q_first + kTileM - 1) / kTileM));
    const int mt_pair = min(MT, (p.nq - gpair * 2 * (MT * kTileM) + kTileM - 1) / kTileM);   // leader's count
    const int total_tiles = (p.nt + kTileN - 1) / kTileN;
    auto tiles_in_range = [&](int r) { return min(p.range_tiles, total_tiles - r * p.range_tiles); };

    if (tid == 0) {
        tc::mbar_init(&bars->a_full, 2 * kThreads2);
        for (int s = 0; s < kBStages2; ++s) {
            tc::mbar_init(&bars->b_full[s], 2 * kExp2Threads);
            tc::mbar_init(&bars->b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&bars->acc_full[s], 1);
            tc::mbar_init(&bars->acc_empty[s], 2 * kEpiThreads);
        }
        tc::fence_barrier_init();
    }
    const int mma_warp = kEpiWarps + kExp2Warps;
    if (warp == mma_warp) tc::tmem_alloc_2cta(&bars->tmem_base, 512);
    tc::tc_fence_before();
    tc::cluster_sync();          // barrier inits + TMEM allocation visible to both CTAs
    tc::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    // Query tiles: expanded once per cluster by ALL threads (loads issued first, then the expansion), so the
    // start-up phase is short -- it is pure overhead on the sharded path, where one launch lasts ~0.3 ms.
    {
        const uint32_t sA_addr = tc::smem_u32(sA);
        const int n_rows = mt_mine * kTileM;
        constexpr int kPer = (kMaxMT2 * kTileM + kThreads2 - 1) / kThreads2;   // 2 rows per thread at most
        uint4 d0[kPer], d1[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int r = tid + i * kThreads2;
            if (r < n_rows) {
                const uint4 *src = reinterpret_cast<const uint4 *>(q + (long long)min(q_first + r, p.nq - 1) * 8);
                d0[i] = __ldg(src);
                d1[i] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int r = tid + i * kThreads2;
            if (r < n_rows) tc::expand_row_to_smem(sA_addr + (uint32_t)(r / kTileM) * kATileBytes, r % kTileM, d0[i], d1[i]);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive_cluster(&bars->a_full, 0);
    }

    if (warp < kEpiWarps) {
        // ===================== epilogue (own TMEM: own 128 query rows x 256 train columns) =====================
        const int set = warp >> 2, quad = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        const int n_slots = p.cpg * p.n_epochs;
        float2 *cand = nullptr;
        float b1[MT], b2[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) { b1[m] = -FLT_MAX; b2[m] = -FLT_MAX; }
        auto flush = [&](int epoch) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int qi = q_first + m * kTileM + quad * 32 + lane;
                if (m < mt_mine && qi < p.nq) cand[(long long)qi * n_slots * 2 + epoch * 2] = make_float2(b1[m], b2[m]);
                b1[m] = -FLT_MAX;
                b2[m] = -FLT_MAX;
            }
        };
        int job = 0;
        for (int link = 0; link < n_links; ++link) {
        cand = p.cand + ((long long)link_prob(link) * p.nq) * n_slots * 2 + (long long)(unit * p.n_epochs) * 2 + set;
        int epoch = 0, lt = 0, j = 0;     // lt = tiles seen in this epoch, j = ranges walked
        for (int r = unit; r < p.n_ranges; r += p.cpg, ++j) {
            if (j > 0 && j % p.rpe == 0) { flush(epoch); ++epoch; lt = 0; }
            const int col_first = r * p.range_tiles * kTileN;
            const int col_end = min(p.nt, col_first + p.range_tiles * kTileN);
            const int n_tiles = (col_end - col_first + kTileN - 1) / kTileN;
            for (int bt = 0; bt < n_tiles; ++bt, ++lt) {
                const int valid_cols = min(kTileN, col_end - (col_first + bt * kTileN));
                // Chained batches track 16-column candidate chunks (CH = 16): the refine pass re-scores two chunks per
                // query and is the second-largest cost of a batch of frame-sized problems, so halving the chunk
                // halves it for ~20 % more FMNMX here; long single problems keep 32 (their refine is negligible).
                constexpr int CH = CHAIN ? 16 : kChunk;
                constexpr int kCPT = kTileN / CH;
                const int n_chunks = (valid_cols + CH - 1) / CH;
                const float chunk_bias = (float)(kChunkMask - lt * kCPT);
                auto track = [&](float &t1, float &t2, float key) {
                    t2 = fmaxf(t2, fminf(t1, key));
                    t1 = fmaxf(t1, key);
                };
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    if (m < mt_pair) {
                        const int ab = job & 1;
                        tc::mbar_wait(&bars->acc_full[ab], (job >> 1) & 1, 10 + ab);
                        tc::tc_fence_after();
                        const uint32_t acc_addr = lane_addr + ab * kTileN + set * kChunk;
                        if (m < mt_mine && n_chunks == kCPT) {
                            // full tile: all four TMEM loads of this warp in flight at once; the accumulator is
                            // handed back as soon as they have landed, BEFORE the reduction -- with two accumulators
                            // the MMA may only run ahead by one job, so the hand-back latency is on the critical path
                            uint32_t v0[32], v1[32], v2[32], v3[32];
                            tc::tmem_ld32_nowait(acc_addr + 0 * kChunk, v0);
                            tc::tmem_ld32_nowait(acc_addr + 2 * kChunk, v1);
                            tc::tmem_ld32_nowait(acc_addr + 4 * kChunk, v2);
                            tc::tmem_ld32_nowait(acc_addr + 6 * kChunk, v3);
                            tc::tmem_ld_wait(v0);
                            tc::pin_regs(v1);
                            tc::pin_regs(v2);
                            tc::pin_regs(v3);
                            tc::tc_fence_before();
                            tc::mbar_arrive_cluster_relaxed(&bars->acc_empty[ab], 0);
                            if constexpr (CH == kChunk) {
                                const float base_bias = chunk_bias - (float)set;
                                track(b1[m], b2[m], fmaf(max32(v0), kKeyScale, base_bias));
                                track(b1[m], b2[m], fmaf(max32(v1), kKeyScale, base_bias - 2.0f));
                                track(b1[m], b2[m], fmaf(max32(v2), kKeyScale, base_bias - 4.0f));
                                track(b1[m], b2[m], fmaf(max32(v3), kKeyScale, base_bias - 6.0f));
                            } else {
                                // 32-column block b = set + 2 cc holds the 16-column chunks 2 b and 2 b + 1
                                const float base_bias = chunk_bias - (float)(2 * set);
                                track(b1[m], b2[m], fmaf(max16<0>(v0), kKeyScale, base_bias));
                                track(b1[m], b2[m], fmaf(max16<16>(v0), kKeyScale, base_bias - 1.0f));
                                track(b1[m], b2[m], fmaf(max16<0>(v1), kKeyScale, base_bias - 4.0f));
                                track(b1[m], b2[m], fmaf(max16<16>(v1), kKeyScale, base_bias - 5.0f));
                                track(b1[m], b2[m], fmaf(max16<0>(v2), kKeyScale, base_bias - 8.0f));
                                track(b1[m], b2[m], fmaf(max16<16>(v2), kKeyScale, base_bias - 9.0f));
                                track(b1[m], b2[m], fmaf(max16<0>(v3), kKeyScale, base_bias - 12.0f));
                                track(b1[m], b2[m], fmaf(max16<16>(v3), kKeyScale, base_bias - 13.0f));
                            }
                        } else {
                            if (m < mt_mine) {
                                // last, partial tile of the train set: only chunks that contain valid columns count
#pragma unroll
                                for (int cc = 0; cc < kChunksPerTile / 2; ++cc) {
                                    const int blk = 2 * cc + set;          // 32-column block of the accumulator
                                    if (blk * kChunk < valid_cols) {
                                        uint32_t v[32];
                                        tc::tmem_ld32(acc_addr + cc * 2 * kChunk, v);
                                        if constexpr (CH == kChunk) {
                                            track(b1[m], b2[m], fmaf(max32(v), kKeyScale, chunk_bias - (float)blk));
                                        } else {
                                            track(b1[m], b2[m], fmaf(max16<0>(v), kKeyScale, chunk_bias - (float)(2 * blk)));
                                            if (2 * blk + 1 < n_chunks)
                                                track(b1[m], b2[m], fmaf(max16<16>(v), kKeyScale, chunk_bias - (float)(2 * blk + 1)));
                                        }
                                    }
                                }
                            }
                            tc::tc_fence_before();
                            tc::mbar_arrive_cluster_relaxed(&bars->acc_empty[ab], 0);
                        }
                        ++job;
                    }
                }
            }
        }
        for (; epoch < p.n_epochs; ++epoch) flush(epoch);     // remaining epochs are written as "none"
        }
    } else if (warp < mma_warp) {
        // ===================== expanders: this CTA's query tiles + its half of every train tile =====================
        const int et = tid - kEpiThreads;   // 0..95
        const uint32_t sB_addr = tc::smem_u32(sB);
        // 128 rows per half tile over 96 threads: thread et expands row et, threads 0..31 also row 96 + et
        const bool two_rows = et < 128 - kExp2Threads;
        auto load_row = [&](const uint32_t *tl, int r, int bt, int row_in_half, uint4 &d0, uint4 &d1) {
            const int row = min((r * p.range_tiles + bt) * kTileN + (int)rank * 128 + row_in_half, p.nt - 1);
            const uint4 *src = reinterpret_cast<const uint4 *>(tl + (long long)row * 8);
            d0 = __ldg(src);
            d1 = __ldg(src + 1);
        };
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0, m0 = n0, m1 = n0;
        const uint32_t *t_link = link_train(0);
        if (unit < p.n_ranges) {
            load_row(t_link, unit, 0, et, n0, n1);
            if (two_rows) load_row(t_link, unit, 0, et + kExp2Threads, m0, m1);
        }
        int s = 0, ph = 0;
        for (int link = 0; link < n_links; ++link) {
        const uint32_t *t_next = link + 1 < n_links ? link_train(link + 1) : nullptr;
        for (int r = unit; r < p.n_ranges; r += p.cpg) {
            const int n_tiles = tiles_in_range(r);
            for (int bt = 0; bt < n_tiles; ++bt) {
                const uint4 c0 = n0, c1 = n1, e0 = m0, e1 = m1;
                // prefetch the rows of the next tile this cluster will see (possibly in its next range, or in the
                // next train frame of a chained batch)
                int r2 = r, bt2 = bt + 1;
                const uint32_t *t2 = t_link;
                if (bt2 == n_tiles) { r2 = r + p.cpg; bt2 = 0; }
                if (r2 >= p.n_ranges && t_next != nullptr) { t2 = t_next; r2 = unit; }
                if (r2 < p.n_ranges) {
                    load_row(t2, r2, bt2, et, n0, n1);
                    if (two_rows) load_row(t2, r2, bt2, et + kExp2Threads, m0, m1);
                }
                tc::mbar_wait_backoff(&bars->b_empty[s], ph ^ 1, 20 + s, 200);
                tc::expand_row_to_smem(sB_addr + (uint32_t)s * kBHalfBytes, et, c0, c1);
                if (two_rows) tc::expand_row_to_smem(sB_addr + (uint32_t)s * kBHalfBytes, et + kExp2Threads, e0, e1);
                tc::fence_proxy_async();
                tc::mbar_arrive_cluster(&bars->b_full[s], 0);
                if (++s == kBStages2) { s = 0; ph ^= 1; }
            }
        }
        t_link = t_next;
        }
    } else {
        // ===================== MMA issuer: leader CTA; the warp stays converged, one elected lane issues ==========
        if (rank == 0) {
            const uint32_t idesc = tc::idesc_e4m3_f32(2 * kTileM, kTileN);
            const uint32_t a_lo0 = tc::smem_desc_lo(tc::smem_u32(sA)), b_lo0 = tc::smem_desc_lo(tc::smem_u32(sB));
            tc::mbar_wait_cluster(&bars->a_full, 0, 30);
            tc::tc_fence_after();
            int job = 0, s = 0, ph = 0;
            for (int link = 0; link < n_links; ++link)
            for (int r = unit; r < p.n_ranges; r += p.cpg) {
                const int n_tiles = tiles_in_range(r);
                for (int bt = 0; bt < n_tiles; ++bt) {
                    tc::mbar_wait_cluster(&bars->b_full[s], ph, 31 + s);
                    tc::tc_fence_after();
                    const uint32_t b_lo = b_lo0 + s * (kBHalfBytes >> 4);
                    for (int m = 0; m < mt_pair; ++m) {
                        const int ab = job & 1;
                        tc::mbar_wait_cluster(&bars->acc_empty[ab], ((job >> 1) & 1) ^ 1, 40 + ab);
                        tc::tc_fence_after();
                        if (tc::elect_one()) {
                            tc::umma_job<2>(tmem + ab * kTileN, a_lo0 + m * (kATileBytes >> 4), b_lo, idesc);
                            tc::umma_commit_2cta(&bars->acc_full[ab], 3);
                        }
                        __syncwarp();
                        ++job;
                    }
                    if (tc::elect_one()) tc::umma_commit_2cta(&bars->b_empty[s], 3);
                    __syncwarp();
                    if (++s == kBStages2) { s = 0; ph ^= 1; }
                }
            }
        }
    }

    tc::tc_fence_before();
    tc::cluster_sync();          // nobody exits while the peer may still signal its barriers / read its smem
    if (warp == mma_warp) tc::tmem_dealloc_2cta(tmem, 512);
}

// ---- refine: exact re-scoring of the candidate chunks -------------------------------------------------
__device__ __forceinline__ void top2_insert(unsigned long long &k1, unsigned long long &k2, unsigned long long key)
{
    unsigned long long m = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, m);
}

__device__ __forceinline__ void top2_insert_max(unsigned long long &k1, unsigned long long &k2, unsigned long long key)
{
    unsigned long long m = min(k1, key);
    k1 = max(k1, key);
    k2 = max(k2, m);
}

// Candidate key -> (dot + 257) << 32 | ~global chunk, so that a 64-bit max prefers the larger dot and, on ties,
// the chunk with the lower global index.  Undoes the unit's strided range walk: the key carries the chunk counter
// inside the unit's candidate epoch.
// x / d and x % d for 0 <= x < 2^17, d >= 1 through a float reciprocal: (x + 0.5) / d stays at least 0.5 / d away from every
// integer while the float error is below 2^-5 / d, so the truncation is exact -- ~8 instructions instead of ~25 per integer
// division (this function runs for every candidate of every query: 296 per query on config 5).
__device__ __forceinline__ void small_divmod(int x, int d, int &q, int &r)
{
    q = __float2int_rz(((float)x + 0.5f) * __frcp_rn((float)d));
    r = x - q * d;
}
__device__ __forceinline__ unsigned long long cand_to_chunk_key(const TcParams &p, float key, int slot)
{
    const int ki = (int)key + kKeyBias;
    const int lc = kChunkMask - (ki & kChunkMask);
    int unit = slot, epoch = 0, lt, lcr, rq, rr;
    if (p.n_epochs > 1) small_divmod(slot, p.n_epochs, unit, epoch);
    small_divmod(lc, p.tile_n / p.chunk, lt, lcr);            // tile counter inside the epoch, chunk inside the tile
    small_divmod(lt, p.range_tiles, rq, rr);
    const int range = unit + (epoch * p.rpe + rq) * p.cpg;
    const unsigned gchunk = (unsigned)((range * p.range_tiles + rr) * (p.tile_n / p.chunk) + lcr);
    return ((unsigned long long)((ki >> kChunkBits) + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - gchunk);
}
// The same for a plan with one unit and one epoch per group (cpg * n_epochs == 1): the chunk counter is the global chunk --
// none of the five integer divisions above, which were half of the refine kernel's instructions on config 4.
__device__ __forceinline__ unsigned long long cand_to_chunk_key_1(float key)
{
    const int ki = (int)key + kKeyBias;
    return ((unsigned long long)((ki >> kChunkBits) + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(kChunkMask - (ki & kChunkMask)));
}

// Exact distance of one row pair, the arithmetic of the reference's matcher (XOR + POPC over 256 bits) -- with three
// carry-save adders in front of the POPCs: POPC is a quarter-rate (XU) instruction and eight per row bounded the refine
// kernels (XU 64 clk per warp and row against 36 on the ALU pipe).  csa(a, b, c) = (a ^ b ^ c, maj(a, b, c)) is two LOP3;
// 8 words -> 2 words of weight 1 + 3 of weight 2 = 5 POPC + 6 extra LOP3, which balances the two pipes at ~40 clk.
__device__ __forceinline__ void csa(unsigned a, unsigned b, unsigned c, unsigned &sum, unsigned &carry)
{
    sum = a ^ b ^ c;
    carry = (a & b) | (a & c) | (b & c);
}
__device__ __forceinline__ unsigned hamming256(const uint4 &qa, const uint4 &qb, const uint4 &ta, const uint4 &tb)
{
    unsigned s1, c1, s2, c2, s3, c3;
    csa(qa.x ^ ta.x, qa.y ^ ta.y, qa.z ^ ta.z, s1, c1);
    csa(qa.w ^ ta.w, qb.x ^ tb.x, qb.y ^ tb.y, s2, c2);
    csa(s1, s2, qb.z ^ tb.z, s3, c3);
    return __popc(s3) + __popc(qb.w ^ tb.w) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

// Exact key of a candidate row inside the two winning chunks: (distance << 8) | (chunk selector << 7) | position, so a
// 32-bit min is the (distance, global index) order -- the reference tie-break -- as long as the lower-index chunk has
// selector 0.  Chunks hold at most 128 rows.
__device__ __forceinline__ unsigned refine_key(unsigned d, int sel, int pos) { return (d << 8) | ((unsigned)sel << 7) | (unsigned)pos; }
__device__ __forceinline__ unsigned long long refine_widen(unsigned k, unsigned glo, unsigned ghi, int chunk, long long base)
{
    if (k == 0xFFFFFFFFu) return kKeyNone;
    const long long row = (long long)((k & 128u) ? ghi : glo) * chunk + (k & 127u);
    return ((unsigned long long)(k >> 8) << 32) | (unsigned long long)(base + row);
}

// Exact top-2 of the rows of the two winning chunks (glo < ghi; 0xFFFFFFFF = no such chunk) of a train set in global
// memory, for the G lanes of one query: the 2 x chunk positions as ONE run over the lanes, U rows per lane in flight (the rows
// come from L2 / HBM), branch-free -- a position outside the run or the train set, or of a chunk that does not exist
// (row < 0), reads a clamped row and its key is replaced by "none".
template <int G, int U>
__device__ __forceinline__ void refine_rows(const uint32_t *t, int nt, int chunk, unsigned glo, unsigned ghi, int sub,
                                            const uint4 &qa, const uint4 &qb, unsigned &k1, unsigned &k2)
{
    const uint4 *t4 = reinterpret_cast<const uint4 *>(t);
    const int lo0 = (int)glo * chunk, hi0 = (int)ghi * chunk - chunk, n_pos = 2 * chunk;
    const unsigned last = (unsigned)max(nt, 1) - 1u;
    for (int e0 = sub; e0 < n_pos; e0 += U * G) {
        uint4 ta[U], tb[U];
        unsigned row[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * G;
            const unsigned r = (unsigned)((e >= chunk ? hi0 : lo0) + e);
            row[u] = e < n_pos ? r : 0xFFFFFFFFu;
            const uint4 *ts = t4 + 2ull * min(r, last);
            ta[u] = __ldg(ts);
            tb[u] = __ldg(ts + 1);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * G;
            const unsigned code = (unsigned)e + (e >= chunk ? 128u - (unsigned)chunk : 0u);   // (selector << 7) | position
            unsigned key = (hamming256(qa, qb, ta[u], tb[u]) << 8) | code;
            key = row[u] < (unsigned)nt ? key : 0xFFFFFFFFu;
            const unsigned m = max(k1, key);
            k1 = min(k1, key);
            k2 = min(k2, m);
        }
    }
}

// G lanes per (problem, query); 32 / G queries per warp.
//   phase 1: reduce the query's candidate chunk keys (2 per epilogue set per range) to the best two chunks of
//            the whole train set by (max dot desc, global chunk index asc).  The exact top-2 rows lie inside
//            them: the nearest row's chunk has the largest chunk maximum (lowest chunk on ties, because the
//            row has the lowest index among its ties); the runner-up is either in the same chunk or is the
//            best row outside it, which by the same argument is in the second-best chunk.
//   phase 2: re-score those 2 x chunk rows with XOR+POPC and keep the exact top-2 by (distance, global index).
// Sharded path (ex.world > 0): the warp stores the query's keys straight into every peer's exchange buffer and the
// kernel's last block publishes this rank's flags (exchange.cuh); exchange_wait_merge_kernel follows.
constexpr int kRefineMaxIters = 8;
template <int G>
__global__ void __launch_bounds__(256) tc_refine_kernel(TcParams p, long long base, unsigned long long *keys_out,
                                                        slm_exchange ex, int iters)
{
    constexpr int kQPW = 32 / G;   // queries per warp
    constexpr int kQPB = 8 * kQPW; // queries per block and iteration (256 threads)
    // sharded path: the block's keys are staged here and leave as contiguous runs, one per peer
    __shared__ ulonglong2 s_keys[kQPB * kRefineMaxIters];
    slm_pdl_launch_dependents();   // the exchange's wait + merge kernel may become resident while this one drains
    const int lane = threadIdx.x & 31, sub = lane % G;
    const long long n_q = (long long)p.n_prob * p.nq;
    const long long q_block = (long long)blockIdx.x * kQPB * iters;           // first query of this block
    // `iters` consecutive groups of kQPB queries per block (> 1 only on the sharded path with many queries, where one
    // system-scope fence per block is then amortised over 8 x as many peer stores)
    for (int it = 0; it < iters; ++it) {
    const long long gq = q_block + (long long)it * kQPB + (threadIdx.x >> 5) * kQPW + lane / G;
    const bool live = gq < n_q;
    const long long gqc = live ? gq : 0;
    const int prob = p.n_prob == 1 ? 0 : (int)(gqc / p.nq);          // (a 64-bit division is ~100 instructions)
    const int qi = p.n_prob == 1 ? (int)gqc : (int)(gqc - (long long)prob * p.nq);
    const uint32_t *q = p.q;
    const uint32_t *t = p.t;
    if (p.desc != nullptr) {
        int2 pr = reinterpret_cast<const int2 *>(p.pairs)[prob];
        q = p.desc + (long long)pr.x * p.frame_words;
        t = p.desc + (long long)pr.y * p.frame_words;
    }
    const uint4 *qs = reinterpret_cast<const uint4 *>(q + (long long)qi * 8);
    const uint4 qa = __ldg(qs), qb = __ldg(qs + 1);        // inputs of the call: readable before the search has finished
    if (it == 0) slm_pdl_wait();                           // the candidate keys of the search kernel are visible from here
    // ---- phase 1 ----
    const int n_cand = p.cpg * p.n_epochs * 4;   // (unit, epoch, set, best/second)
    const float *cand = reinterpret_cast<const float *>(p.cand) + gqc * (long long)n_cand;
    unsigned long long c1 = 0, c2 = 0;   // 0 = none
    auto consider = [&](int ci) {
        const float key = cand[ci];
        if (key > -1.0e30f) top2_insert_max(c1, c2, cand_to_chunk_key(p, key, ci >> 2));
    };
    if (n_cand == 4) {
        // one unit, one epoch (the usual plan): the four packed keys are directly comparable (larger = larger dot, then lower
        // chunk) and name four different chunks -- a top-2 of four floats, and only the two winners are converted
        const float4 c4 = *reinterpret_cast<const float4 *>(cand);
        const float ca = fmaxf(c4.x, c4.y), cb = fminf(c4.x, c4.y), cc = fmaxf(c4.z, c4.w), cd = fminf(c4.z, c4.w);
        const float f1 = fmaxf(ca, cc), f2 = fmaxf(fminf(ca, cc), fmaxf(cb, cd));
        c1 = f1 > -1.0e30f ? cand_to_chunk_key_1(f1) : 0ull;
        c2 = f2 > -1.0e30f ? cand_to_chunk_key_1(f2) : 0ull;
    } else if (n_cand <= 8) {
        for (int ci = 0; ci < n_cand; ++ci) consider(ci);          // every lane of the group scans all: no shuffles
    } else {
#pragma unroll 4
        for (int ci = sub; ci < n_cand; ci += G) consider(ci);
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, c1, o);
            const unsigned long long o2 = __shfl_xor_sync(0xFFFFFFFFu, c2, o);
            top2_insert_max(c1, c2, o1);
            top2_insert_max(c1, c2, o2);
        }
    }
    // ---- phase 2 ----
    unsigned ga = c1 ? 0xFFFFFFFFu - (unsigned)(c1 & 0xFFFFFFFFull) : 0xFFFFFFFFu;
    unsigned gb = c2 ? 0xFFFFFFFFu - (unsigned)(c2 & 0xFFFFFFFFull) : 0xFFFFFFFFu;
    const unsigned glo = min(ga, gb), ghi = max(ga, gb);           // 0xFFFFFFFF = no such chunk
    unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
    // 2 x chunk / G positions per lane: 30 / 10 / 5 (chunks of 120 / 40 / 20 rows) in batches of five, 8 / 4 (32 / 16) of four
    if ((2 * p.chunk / G) % 5 == 0) refine_rows<G, 5>(t, p.nt, p.chunk, glo, ghi, sub, qa, qb, k1, k2);
    else refine_rows<G, 4>(t, p.nt, p.chunk, glo, ghi, sub, qa, qb, k1, k2);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const unsigned o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, o);
        const unsigned o2 = __shfl_xor_sync(0xFFFFFFFFu, k2, o);
        unsigned m = max(k1, o1);
        k1 = min(k1, o1);
        k2 = min(min(k2, o2), m);
    }
    if (live && sub == 0) {
        const unsigned long long w1 = refine_widen(k1, glo, ghi, p.chunk, base), w2 = refine_widen(k2, glo, ghi, p.chunk, base);
        if (keys_out) reinterpret_cast<ulonglong2 *>(keys_out)[gq] = make_ulonglong2(w1, w2);
        if (ex.world > 0) s_keys[it * kQPB + (threadIdx.x >> 5) * kQPW + lane / G] = make_ulonglong2(w1, w2);
    }
    }   // iterations
    if (ex.world > 0) {
        // a run of 32 queries is 256 B of compact keys: whole NVLink packets instead of scattered 8-byte stores; only the
        // threads that stored fence (a system-scope fence per thread of every block was 0.2 ms on config 4)
        __syncthreads();
        const int n_here = (int)max(0ll, min((long long)kQPB * iters, n_q - q_block));
        bool wrote = false;
        for (int i = threadIdx.x; i < n_here * ex.world; i += blockDim.x) {
            const int r = i / n_here, k = i - r * n_here;
            slm_exchange_store_to(ex, r, q_block + k, s_keys[k].x, s_keys[k].y);
            wrote = true;
        }
        slm_exchange_publish(ex, wrote);
    }
}

// ---- two-phase form of the sharded refine (many queries) ----------------------------------------
// tc_refine_kernel re-scores two chunks per query ON EVERY RANK: constant work per rank, however many ranks share the train
// set -- on config 4 (1M queries) that is 0.34 ms next to 0.6 ms of search at 8 ranks.  But the chunk keys already hold the
// exact best distance of every chunk, and the proof of the single-GPU refine holds for the union of all ranks' chunks: the
// global top-2 rows lie in the GLOBAL best two chunks by (max dot desc, first row asc).  So:
//   phase 0  tc_chunk_keys_kernel     every rank sends its best two chunk keys per query (dot, first global row) to all ranks
//   phase 1  tc_refine_owned_kernel   every rank picks the global best two chunks per query from all ranks' keys and
//                                     re-scores only those that lie in ITS row block (on average 2 / world per query),
//                                     then sends its exact top-2 of them; exchange_wait_merge_kernel merges as before.
template <int G>
__global__ void __launch_bounds__(256) tc_chunk_keys_kernel(TcParams p, long long base, slm_exchange ex, int iters)
{
    constexpr int kQPW = 32 / G, kQPB = 8 * kQPW;
    __shared__ ulonglong2 s_keys[kQPB * kRefineMaxIters];
    slm_pdl_launch_dependents();
    const int lane = threadIdx.x & 31, sub = lane % G;
    const long long n_q = p.nq;
    const long long q_block = (long long)blockIdx.x * kQPB * iters;
    slm_pdl_wait();
    const int n_cand = p.cpg * p.n_epochs * 4;
    for (int it = 0; it < iters; ++it) {
        const long long gq = q_block + (long long)it * kQPB + (threadIdx.x >> 5) * kQPW + lane / G;
        const bool live = gq < n_q;
        const float *cand = reinterpret_cast<const float *>(p.cand) + (live ? gq : 0) * (long long)n_cand;
        unsigned long long c1 = 0, c2 = 0;
        auto consider = [&](int ci) {
            const float key = cand[ci];
            if (key > -1.0e30f) top2_insert_max(c1, c2, cand_to_chunk_key(p, key, ci >> 2));
        };
        if (n_cand == 4) {                          // as in tc_refine_kernel
            const float4 c4 = *reinterpret_cast<const float4 *>(cand);
            const float ca = fmaxf(c4.x, c4.y), cb = fminf(c4.x, c4.y), cc = fmaxf(c4.z, c4.w), cd = fminf(c4.z, c4.w);
            const float f1 = fmaxf(ca, cc), f2 = fmaxf(fminf(ca, cc), fmaxf(cb, cd));
            c1 = f1 > -1.0e30f ? cand_to_chunk_key_1(f1) : 0ull;
            c2 = f2 > -1.0e30f ? cand_to_chunk_key_1(f2) : 0ull;
        } else if (n_cand <= 8) {
            for (int ci = 0; ci < n_cand; ++ci) consider(ci);
        } else {
            for (int ci = sub; ci < n_cand; ci += G) consider(ci);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, c1, o);
                const unsigned long long o2 = __shfl_xor_sync(0xFFFFFFFFu, c2, o);
                top2_insert_max(c1, c2, o1);
                top2_insert_max(c1, c2, o2);
            }
        }
        if (live && sub == 0) {
            // (dot + 257) << 32 | ~local chunk  ->  (dot + 257) << 32 | ~first GLOBAL row of the chunk
            auto globalise = [&](unsigned long long c) -> unsigned long long {
                if (c == 0) return 0ull;
                const unsigned gchunk = 0xFFFFFFFFu - (unsigned)(c & 0xFFFFFFFFull);
                const unsigned first_row = (unsigned)(base + (long long)gchunk * p.chunk);
                return (c & 0xFFFFFFFF00000000ull) | (unsigned long long)(0xFFFFFFFFu - first_row);
            };
            s_keys[it * kQPB + (threadIdx.x >> 5) * kQPW + lane / G] = make_ulonglong2(globalise(c1), globalise(c2));
        }
    }
    __syncthreads();
    const int n_here = (int)max(0ll, min((long long)kQPB * iters, n_q - q_block));
    bool wrote = false;
    for (int i = threadIdx.x; i < n_here * ex.world; i += blockDim.x) {
        const int r = i / n_here, k = i - r * n_here;
        slm_exchange_store_chunks_to(ex, r, q_block + k, s_keys[k].x, s_keys[k].y);
        wrote = true;
    }
    slm_exchange_publish(ex, wrote, 0);
}

// The usual plan (one unit, one epoch per group: four candidates per query): ONE THREAD per query -- a float top-2 of the four
// packed keys, two conversions, and a coalesced 8- or 16-byte store per peer (a warp's stores to one peer are one run).
__global__ void __launch_bounds__(256) tc_chunk_keys_flat_kernel(TcParams p, long long base, slm_exchange ex)
{
    slm_pdl_launch_dependents();
    slm_pdl_wait();                                        // the candidate keys of the search kernel are visible from here
    const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
    if (q < p.nq) {
        const float4 c4 = reinterpret_cast<const float4 *>(p.cand)[q];
        const float ca = fmaxf(c4.x, c4.y), cb = fminf(c4.x, c4.y), cc = fmaxf(c4.z, c4.w), cd = fminf(c4.z, c4.w);
        const float f1 = fmaxf(ca, cc), f2 = fmaxf(fminf(ca, cc), fmaxf(cb, cd));
        // (dot + 257) << 32 | ~first GLOBAL row of the chunk
        auto chunk_key = [&](float f) -> unsigned long long {
            if (!(f > -1.0e30f)) return 0ull;
            const int ki = (int)f + kKeyBias;
            const unsigned first_row = (unsigned)(base + (long long)(kChunkMask - (ki & kChunkMask)) * p.chunk);
            return ((unsigned long long)((ki >> kChunkBits) + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - first_row);
        };
        const unsigned long long k1 = chunk_key(f1), k2 = chunk_key(f2);
        for (int r = 0; r < ex.world; ++r) slm_exchange_store_chunks_to(ex, r, q, k1, k2);
    }
    slm_exchange_publish(ex, true, 0);
}

// Tiles of 256 queries per block.  (A) one thread per query reads every rank's two chunk keys (coalesced), keeps the global
// best two chunks and queues those that lie in THIS rank's row block as work items; (B) the items -- on average 2 / world per
// query -- are re-scored densely, 8 lanes per item; (C) one thread per query combines its items and stores the exact keys to
// every peer (coalesced).  Only owned chunks cost anything: the re-scoring is shared between the ranks, not repeated.
constexpr int kOwnedTile = 256;
__global__ void __launch_bounds__(kOwnedTile) tc_refine_owned_kernel(TcParams p, long long base, slm_exchange ex)
{
    constexpr int G = 8, U = 5;
    __shared__ uint2 s_first[kOwnedTile];          // first global rows of the query's best two chunks (lower rows first)
    __shared__ uint4 s_item[kOwnedTile];           // exact (k1, k2) of chunk 0 and of chunk 1, 0xFFFFFFFF = none
    __shared__ unsigned short s_queue[2 * kOwnedTile];
    __shared__ int s_n;
    slm_pdl_launch_dependents();
    if (!slm_exchange_wait_flags(ex, 0)) return;          // a lost peer: reported; this rank publishes nothing either
    const int tid = threadIdx.x, sub = tid % G, grp = tid / G;
    const long long n_q = p.nq;
    const uint4 *t4 = reinterpret_cast<const uint4 *>(p.t);
    const unsigned last = (unsigned)max(p.nt, 1) - 1u;
    for (long long q0 = (long long)blockIdx.x * kOwnedTile; q0 < n_q; q0 += (long long)gridDim.x * kOwnedTile) {
        if (tid == 0) s_n = 0;
        __syncthreads();
        // ---- (A) ----
        const long long gq = q0 + tid;
        const bool live = gq < n_q;
        unsigned flo = 0xFFFFFFFFu, fhi = 0xFFFFFFFFu;
        if (live) {
            unsigned long long c1 = 0, c2 = 0;
            for (int i = 0; i < ex.world; ++i) {
                const long long slot = slm_exchange_slot(ex, 0, i) + gq;
                if (ex.key_bytes == 4) {
                    const uint2 g = reinterpret_cast<const uint2 *>(ex.peer_keys[ex.rank])[slot];
                    top2_insert_max(c1, c2, slm_chunk_widen(g.x));
                    top2_insert_max(c1, c2, slm_chunk_widen(g.y));
                } else {
                    const ulonglong2 g = reinterpret_cast<const ulonglong2 *>(ex.peer_keys[ex.rank])[slot];
                    top2_insert_max(c1, c2, g.x);
                    top2_insert_max(c1, c2, g.y);
                }
            }
            const unsigned fa = c1 ? 0xFFFFFFFFu - (unsigned)(c1 & 0xFFFFFFFFull) : 0xFFFFFFFFu;
            const unsigned fb = c2 ? 0xFFFFFFFFu - (unsigned)(c2 & 0xFFFFFFFFull) : 0xFFFFFFFFu;
            flo = min(fa, fb);
            fhi = max(fa, fb);
#pragma unroll
            for (int sel = 0; sel < 2; ++sel) {
                const unsigned f = sel ? fhi : flo;
                if (f != 0xFFFFFFFFu && (long long)f >= base && (long long)f < base + p.nt)
                    s_queue[atomicAdd(&s_n, 1)] = (unsigned short)(tid * 2 + sel);
            }
        }
        s_first[tid] = make_uint2(flo, fhi);
        s_item[tid] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        __syncthreads();
        // ---- (B) ----
        const int n_items = s_n;
        for (int it = grp; it < n_items; it += kOwnedTile / G) {
            const int item = s_queue[it], ql = item >> 1, sel = item & 1;
            const unsigned f = sel ? s_first[ql].y : s_first[ql].x;
            const unsigned row0 = (unsigned)((long long)f - base);
            const uint4 *qs = reinterpret_cast<const uint4 *>(p.q + (q0 + ql) * 8);
            const uint4 qa = __ldg(qs), qb = __ldg(qs + 1);
            unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
            for (int pos0 = sub; pos0 < p.chunk; pos0 += U * G) {
                uint4 ta[U], tb[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint4 *ts = t4 + 2ull * min(row0 + (unsigned)(pos0 + u * G), last);
                    ta[u] = __ldg(ts);
                    tb[u] = __ldg(ts + 1);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pos = pos0 + u * G;
                    unsigned key = (hamming256(qa, qb, ta[u], tb[u]) << 8) | ((unsigned)sel << 7) | (unsigned)pos;
                    key = (pos < p.chunk && row0 + (unsigned)pos < (unsigned)p.nt) ? key : 0xFFFFFFFFu;
                    const unsigned m = max(k1, key);
                    k1 = min(k1, key);
                    k2 = min(k2, m);
                }
            }
            // the four groups of a warp run different numbers of items: the shuffles name the group's own 8 lanes only
            const unsigned gmask = 0xFFu << ((tid & 31) & ~(G - 1));
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                const unsigned o1 = __shfl_xor_sync(gmask, k1, o), o2 = __shfl_xor_sync(gmask, k2, o);
                const unsigned m = max(k1, o1);
                k1 = min(k1, o1);
                k2 = min(min(k2, o2), m);
            }
            if (sub == 0) {
                if (sel) { s_item[ql].z = k1; s_item[ql].w = k2; }
                else { s_item[ql].x = k1; s_item[ql].y = k2; }
            }
        }
        __syncthreads();
        // ---- (C) ----
        if (live) {
            const uint4 r = s_item[tid];
            const unsigned k1 = min(r.x, r.z), k2 = min(max(r.x, r.z), min(r.y, r.w));
            auto widen = [&](unsigned k) -> unsigned long long {
                if (k == 0xFFFFFFFFu) return kKeyNone;
                return ((unsigned long long)(k >> 8) << 32) | (unsigned long long)(((k & 128u) ? fhi : flo) + (k & 127u));
            };
            const unsigned long long w1 = widen(k1), w2 = widen(k2);
            for (int pr = 0; pr < ex.world; ++pr) slm_exchange_store_to(ex, pr, gq, w1, w2, 1);
        }
    }
    slm_exchange_publish(ex, true, 1);
}

// Refine for batches of frame-sized problems (config 3).  tc_refine_kernel reads every query's candidate rows
// straight from L2: 2 KB per query, 8 GB for 2016 pairs x 2000 queries -- L2-bandwidth bound (0.9 ms, as long as the
// tensor kernel itself).  Here one CTA owns one pair, stages the pair's whole train frame in shared memory once
// (64 KB for 2000 rows) and re-scores all of the pair's queries from there; what is left is the POPC pipe.
// Requires one candidate slot per query (cpg * n_epochs == 1) and nt * 32 bytes of shared memory.
constexpr int kRefineFrameThreads = 512;
__global__ void __launch_bounds__(kRefineFrameThreads) tc_refine_frame_kernel(TcParams p, unsigned long long *keys_out)
{
    extern __shared__ __align__(16) uint4 s_rows[];      // nt rows x 2
    constexpr int G = 8;
    const int prob = blockIdx.x;
    const int2 pr = reinterpret_cast<const int2 *>(p.pairs)[prob];
    const uint32_t *q = p.desc + (long long)pr.x * p.frame_words;
    const uint4 *t4 = reinterpret_cast<const uint4 *>(p.desc + (long long)pr.y * p.frame_words);
    // low halves of all rows first, then the high halves: the 8 lanes of a group read 8 consecutive rows, i.e.
    // 128 contiguous bytes per LDS.128 quarter-warp -- conflict-free (32-byte row stride would be 2-way)
    for (int i = threadIdx.x; i < 2 * p.nt; i += kRefineFrameThreads) s_rows[(i & 1) * p.nt + (i >> 1)] = __ldg(t4 + i);
    slm_pdl_wait();
    __syncthreads();
    const int sub = threadIdx.x % G, grp = threadIdx.x / G;
    const float4 *cand4 = reinterpret_cast<const float4 *>(p.cand) + (long long)prob * p.nq;
    for (int q0 = 0; q0 < p.nq; q0 += kRefineFrameThreads / G) {
        const int qi = q0 + grp;
        const bool live = qi < p.nq;
        const int qc = live ? qi : 0;
        // ---- phase 1: the best two chunks out of the four candidates (two epilogue sets x best / second) ----
        // one unit, one epoch: the chunk counter in a key IS the global chunk, so the packed keys of all four candidates are
        // directly comparable (larger = larger dot, then lower chunk) and the two sets cover disjoint chunks: a top-2 of four
        // floats (7 FMNMX) instead of four 64-bit insertions (a quarter of this kernel's issue slots)
        const float4 c4 = __ldg(cand4 + qc);
        const float ca = fmaxf(c4.x, c4.y), cb = fminf(c4.x, c4.y), cc = fmaxf(c4.z, c4.w), cd = fminf(c4.z, c4.w);
        const float f1 = fmaxf(ca, cc), f2 = fmaxf(fminf(ca, cc), fmaxf(cb, cd));
        const unsigned ga = f1 > -1.0e30f ? (unsigned)(kChunkMask - (((int)f1 + kKeyBias) & kChunkMask)) : 0xFFFFFFFFu;
        const unsigned gb = f2 > -1.0e30f ? (unsigned)(kChunkMask - (((int)f2 + kKeyBias) & kChunkMask)) : 0xFFFFFFFFu;
        const unsigned glo = min(ga, gb), ghi = max(ga, gb);
        // ---- phase 2: exact re-scoring of those rows from shared memory ----
        const uint4 *qs = reinterpret_cast<const uint4 *>(q + (long long)qc * 8);
        const uint4 qa = __ldg(qs), qb = __ldg(qs + 1);
        unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
        // the rows of both chunks as ONE run of 2 x chunk positions over the 8 lanes (20-row chunks: 5 trips, not 2 x 3),
        // branch-free: a position outside the frame (or of a chunk that does not exist: row < 0) reads a clamped row and
        // its key is replaced by "none" -- a branch per row cost a third of the loop's issue slots
        const int lo0 = (int)glo * p.chunk, hi0 = (int)ghi * p.chunk - p.chunk;
        for (int e = sub; e < 2 * p.chunk; e += G) {
            const bool s = e >= p.chunk;
            const unsigned row = (unsigned)((s ? hi0 : lo0) + e);
            const unsigned rc = min(row, (unsigned)p.nt - 1u);
            const unsigned code = (unsigned)e + (s ? 128u - (unsigned)p.chunk : 0u);       // (selector << 7) | position
            unsigned key = (hamming256(qa, qb, s_rows[rc], s_rows[p.nt + rc]) << 8) | code;
            key = row < (unsigned)p.nt ? key : 0xFFFFFFFFu;
            const unsigned m = max(k1, key);
            k1 = min(k1, key);
            k2 = min(k2, m);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const unsigned o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, k2, o);
            const unsigned m = max(k1, o1);
            k1 = min(k1, o1);
            k2 = min(min(k2, o2), m);
        }
        if (live && sub == 0)
            reinterpret_cast<ulonglong2 *>(keys_out)[(long long)prob * p.nq + qi] =
                make_ulonglong2(refine_widen(k1, glo, ghi, p.chunk, 0), refine_widen(k2, glo, ghi, p.chunk, 0));
    }
}

// cudaFuncSetAttribute once per (kernel, device); the flag is atomic so concurrent first calls from two host
// threads are benign (setting the attribute twice is harmless)
template <typename K>
int configure_smem(K kernel, std::atomic<bool> (&configured)[64], size_t smem)
{
    int dev = 0;
    SLM_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63].load(std::memory_order_acquire)) {
        SLM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63].store(true, std::memory_order_release);
    }
    return SLM_OK;
}

template <int MT>
int launch_tc(const TcParams &p, int n_prob, cudaStream_t stream)
{
    const size_t smem = (size_t)MT * kATileBytes + 2 * kBTileBytes + sizeof(TcBarriers) + 64;
    static std::atomic<bool> configured[64];
    SLM_TRY(configure_smem(knn2_tc_kernel<MT>, configured, smem));
    dim3 grid((unsigned)(p.n_groups * p.n_ranges), (unsigned)n_prob);
    knn2_tc_kernel<MT><<<grid, kThreads, smem, stream>>>(p);
    SLM_CUDA(cudaGetLastError());
    return SLM_OK;
}

template <int MT, bool CHAIN>
int launch_tc2_impl(const TcParams &p, int n_prob, cudaStream_t stream)
{
    const size_t smem = (size_t)MT * kATileBytes + kBStages2 * kBHalfBytes + sizeof(TcBarriers2) + 64;
    static std::atomic<bool> configured[64];
    SLM_TRY(configure_smem(knn2_tc2_kernel<MT, CHAIN>, configured, smem));
    const int n_gpairs = (p.n_groups + 1) / 2;
    dim3 grid((unsigned)(2 * n_gpairs * p.cpg), (unsigned)n_prob);
    knn2_tc2_kernel<MT, CHAIN><<<grid, kThreads2, smem, stream>>>(p);
    SLM_CUDA(cudaGetLastError());
    return SLM_OK;
}

template <int MT>
int launch_tc2(const TcParams &p, int n_prob, cudaStream_t stream)
{
    return p.chain_pairs ? launch_tc2_impl<MT, true>(p, n_prob, stream) : launch_tc2_impl<MT, false>(p, n_prob, stream);
}

// Clusters per unit for the resident-cluster kernels: minimise  waves x (start-up + tiles per cluster x cycles per tile)
// with waves = ceil(units * cpg / cluster slots).  Few units (long train sets) end up as one resident wave, many units
// as several waves of longer-lived clusters.
long long plan_cpg(const slm_ctx *ctx, long long units, long long n_tiles, double startup_clk, double tile_clk)
{
    const long long slots = ctx->sm_count / 2;
    long long cpg = 1;
    double best = 1e300;
    long long cpg_max = n_tiles < 4 * slots ? n_tiles : 4 * slots;
    if (cpg_max > ctx->max_cpg) cpg_max = ctx->max_cpg;
    for (long long c = 1; c <= cpg_max; ++c) {
        const long long waves = (units * c + slots - 1) / slots;
        const double cost = (double)waves * (startup_clk + (double)((n_tiles + c - 1) / c) * tile_clk);
        if (cost < best * 0.999) { best = cost; cpg = c; }
    }
    return cpg;
}

int tc_run(slm_ctx *ctx, TcParams p, int n_prob, long long base, uint64_t *keys_out, cudaStream_t stream,
           const slm_exchange *exchange = nullptr, const slm_chain *chain = nullptr, bool allow_fp4 = true,
           int *phase_out = nullptr)
{
    p.n_prob = n_prob;
    const int m_tiles = (p.nq + kTileM - 1) / kTileM;
    // CTA pairs (cta_group::2) need at least two query tiles to keep both halves of the M = 256 MMA busy
    const bool two_cta = m_tiles >= 2 && !ctx->force_1cta;
    // fp4 (kind::mxf4, twice the fp8 rate) whenever CTA pairs apply; batches of frame-sized problems only as chained
    // batches (their fp4 instantiation walks the pairs of a query frame back to back and tracks 20-row chunks)
    const bool chain_ok = chain && chain->n_units > 0 && chain->n_units < n_prob;
    const bool fp4 = two_cta && allow_fp4 && (p.desc == nullptr || chain_ok);
    ctx->last_variant = fp4 ? SLM_VARIANT_TENSOR4 : SLM_VARIANT_TENSOR;
    p.tile_n = fp4 ? tc4::kTileN : kTileN;
    const int n_tiles = (p.nt + p.tile_n - 1) / p.tile_n;
    const long long n_q = (long long)n_prob * p.nq;
    int work_units;                      // CTAs (1-CTA kernel) or CTA pairs (2-CTA kernels) per train range
    if (fp4) {
        int n_groups = 2 * ((m_tiles + 2 * kMaxMT4 - 1) / (2 * kMaxMT4));
        p.mt = (m_tiles + n_groups - 1) / n_groups;
        p.n_groups = (m_tiles + p.mt - 1) / p.mt;
        work_units = (p.n_groups + 1) / 2;
        // wide chunks keep the hot kernel's epilogue at 0.52 ALU ops per element; the refine pass re-scores two chunks per
        // query, so many-query problems take narrower ones
        p.chunk = p.desc != nullptr ? 20 : (ctx->tc4_chunk > 0 ? ctx->tc4_chunk : (n_q <= 32768 ? 120 : 40));
    } else if (two_cta) {
        // split the query tiles evenly over an even number of groups of at most kMaxMT2 tiles
        int n_groups = 2 * ((m_tiles + 2 * kMaxMT2 - 1) / (2 * kMaxMT2));
        p.mt = (m_tiles + n_groups - 1) / n_groups;
        if (ctx->tc_plan_mt && n_prob == 1) {
            // EXPERIMENTAL (SLM_TC_PLAN_MT=1, off by default): also choose the query tiles per CTA.
            // Fewer tiles per CTA = more, shorter-lived clusters with less to expand at start-up -- better for
            // mid-size problems (c2: 2000 x 20000) where the start-up dominates.  Start-up model: 6000 + 2000 per
            // query tile (14000 at 4 tiles, the measured figure).
            const long long slots = ctx->sm_count / 2;
            double best = 1e300;
            int best_mt = p.mt;
            for (int mt = 1; mt <= kMaxMT2; ++mt) {
                const long long units = ((m_tiles + mt - 1) / mt + 1) / 2;
                long long cmax = n_tiles < 4 * slots ? n_tiles : 4 * slots;
                if (cmax > ctx->max_cpg) cmax = ctx->max_cpg;
                for (long long c = 1; c <= cmax; ++c) {
                    const long long waves = (units * c + slots - 1) / slots;
                    const double cost = (double)waves * (6000.0 + 2000.0 * mt + (double)((n_tiles + c - 1) / c) * 1024.0 * mt);
                    if (cost < best * 0.999) { best = cost; best_mt = mt; }
                }
            }
            p.mt = best_mt;
        }
        p.n_groups = (m_tiles + p.mt - 1) / p.mt;
        work_units = (p.n_groups + 1) / 2;
        p.chunk = kChunk;
    } else {
        p.mt = m_tiles >= kMaxMT ? kMaxMT : m_tiles;
        p.n_groups = (m_tiles + p.mt - 1) / p.mt;
        work_units = p.n_groups;
        p.chunk = kChunk;
    }
    if (two_cta) {
        // CTA pairs.  `units` = (group pair, problem) combinations; each gets `cpg` clusters that walk the train
        // ranges with stride cpg.  Few units (long train sets): one resident wave, the query tiles are expanded
        // once per cluster and short ranges keep the static stride balanced.  Many units: several waves of clusters.
        const long long units = (long long)work_units * n_prob;
        const double startup = fp4 ? 8000.0 + 1500.0 * p.mt : (ctx->tc_plan_mt ? 6000.0 + 2000.0 * p.mt : 14000.0);
        const double tile_clk = (fp4 ? 2.0 * tc4::kTileN : 1024.0) * p.mt;
        const long long cpg = plan_cpg(ctx, units, n_tiles, startup, tile_clk);
        long long range_tiles = n_tiles / (cpg * 128);      // short ranges: the static stride stays balanced
        if (range_tiles < 1) range_tiles = 1;
        if (range_tiles > 8) range_tiles = 8;
        const long long n_ranges = (n_tiles + range_tiles - 1) / range_tiles;
        // chunk counters of one candidate epoch must fit the key's 15 bits
        long long epoch_tiles = (1 << kChunkBits) / (p.tile_n / p.chunk);
        if (epoch_tiles > ctx->epoch_tiles) epoch_tiles = ctx->epoch_tiles;
        long long rpe = epoch_tiles / range_tiles;
        if (rpe < 1) rpe = 1;
        const long long ranges_per_unit = (n_ranges + cpg - 1) / cpg;
        p.cpg = (int)cpg;
        p.range_tiles = (int)range_tiles;
        p.n_ranges = (int)n_ranges;
        p.rpe = (int)rpe;
        p.n_epochs = (int)((ranges_per_unit + rpe - 1) / rpe);
        if (units * cpg > 0x3FFFFFFFll) return slm_fail(SLM_ERR_UNSUPPORTED, "problem too large for one launch");
    } else {
        // single CTAs (fewer than two query tiles): one contiguous range per CTA, ~8 CTAs per SM when possible
        const long long slots = ctx->sm_count;
        long long n_ranges = (8 * slots + (long long)work_units * n_prob - 1) / ((long long)work_units * n_prob);
        if (n_ranges < 1) n_ranges = 1;
        long long range_tiles = (n_tiles + n_ranges - 1) / n_ranges;
        if (range_tiles < 2) range_tiles = 2;
        if (range_tiles > kMaxEpochTiles) range_tiles = kMaxEpochTiles;
        if (range_tiles > n_tiles) range_tiles = n_tiles;
        n_ranges = (n_tiles + range_tiles - 1) / range_tiles;
        if ((long long)p.n_groups * n_ranges > 0x3FFFFFFFll)
            return slm_fail(SLM_ERR_UNSUPPORTED, "problem too large for one launch");
        p.range_tiles = (int)range_tiles;
        p.n_ranges = (int)n_ranges;
        p.cpg = (int)n_ranges;
        p.rpe = 1;
        p.n_epochs = 1;
    }
    if (n_prob > 65535) return slm_fail(SLM_ERR_UNSUPPORTED, "at most 65535 problems per launch");

    // Chained batch: short train frames (one cluster per query-group pair walks a whole frame in one epoch) whose
    // pairs share query frames -- grid.y runs over the units instead of the pairs.
    int grid_y = n_prob;
    if (fp4 && p.desc != nullptr) {
        // batched fp4 = chained fp4; a plan that needs several clusters or epochs per pair goes to the fp8 kernels
        if (p.cpg != 1 || p.n_epochs != 1) return tc_run(ctx, p, n_prob, base, keys_out, stream, exchange, chain, false, phase_out);
        p.chain_pairs = chain->pairs_sorted;
        p.chain_prob = chain->prob;
        p.chain_units = chain->units;
        grid_y = chain->n_units;
    } else if (!fp4 && chain_ok && two_cta && p.cpg == 1 && p.n_epochs == 1 && n_tiles <= kMaxEpochTiles / 2) {
        p.chunk = 16;      // the chained instantiation tracks 16-column candidate chunks
        p.chain_pairs = chain->pairs_sorted;
        p.chain_prob = chain->prob;
        p.chain_units = chain->units;
        grid_y = chain->n_units;
    }
    const size_t cand_bytes = (size_t)n_prob * (size_t)p.cpg * p.n_epochs * 2 * (size_t)p.nq * sizeof(float2);
    SLM_TRY(slm_buf_reserve(ctx, &ctx->scratch, cand_bytes));
    p.cand = reinterpret_cast<float2 *>(ctx->scratch.p);

    if (phase_out) *phase_out = 0;
    ctx->last_kernel = fp4 ? "knn2_tc4_kernel" : two_cta ? "knn2_tc2_kernel" : "knn2_tc_kernel";
    SLM_TRY(slm_prof_begin(ctx, stream));
    if (fp4) {
        SLM_TRY(slm_tc4_launch(ctx, p, grid_y, stream));
    } else if (two_cta) {
        switch (p.mt) {
        case 1: SLM_TRY(launch_tc2<1>(p, grid_y, stream)); break;
        case 2: SLM_TRY(launch_tc2<2>(p, grid_y, stream)); break;
        case 3: SLM_TRY(launch_tc2<3>(p, grid_y, stream)); break;
        default: SLM_TRY(launch_tc2<4>(p, grid_y, stream)); break;
        }
    } else {
        switch (p.mt) {
        case 1: SLM_TRY(launch_tc<1>(p, n_prob, stream)); break;
        case 2: SLM_TRY(launch_tc<2>(p, n_prob, stream)); break;
        default: SLM_TRY(launch_tc<3>(p, n_prob, stream)); break;
        }
    }
    SLM_TRY(slm_prof_end(ctx, stream));
    slm_exchange ex{};
    if (exchange) ex = *exchange;
    const size_t frame_smem = (size_t)p.nt * 32;
    // the refine kernels start under programmatic dependent launch: their prologue (query loads, train-frame staging)
    // overlaps the drain of the search kernel, which releases its dependents at start-up
    const bool pdl = !ctx->no_pdl;
    if (p.desc != nullptr && !exchange && p.cpg * p.n_epochs == 1 && n_prob >= ctx->sm_count && frame_smem <= 200 * 1024 &&
        !ctx->no_frame_refine) {
        // batch of frame-sized problems: one CTA per pair, train frame staged in shared memory
        static std::atomic<bool> configured[64];
        SLM_TRY(configure_smem(tc_refine_frame_kernel, configured, 200 * 1024));
        SLM_CUDA(slm_launch(tc_refine_frame_kernel, dim3((unsigned)n_prob), dim3(kRefineFrameThreads), frame_smem, stream, pdl, p,
                            reinterpret_cast<unsigned long long *>(keys_out)));
    } else if (exchange && p.desc == nullptr && ctx->exchange_two_phase_min > 0 && n_q >= ctx->exchange_two_phase_min &&
               (exchange->world >= ctx->exchange_two_phase_world || exchange->world == 1)) {
        // sharded, many queries: agree on the global best two chunks first, only their owners refine them
        const bool few = p.cpg * p.n_epochs * 4 <= 32;
        const int iters = n_q >= 65536 ? kRefineMaxIters : 1;
        if (p.cpg * p.n_epochs == 1)
            SLM_CUDA(slm_launch(tc_chunk_keys_flat_kernel, dim3((unsigned)((n_q + 255) / 256)), dim3(256), 0, stream, pdl, p, base, ex));
        else if (few)
            SLM_CUDA(slm_launch(tc_chunk_keys_kernel<8>, dim3((unsigned)((n_q + 32 * iters - 1) / (32 * iters))), dim3(256), 0, stream,
                                pdl, p, base, ex, iters));
        else
            SLM_CUDA(slm_launch(tc_chunk_keys_kernel<32>, dim3((unsigned)((n_q + 8 * iters - 1) / (8 * iters))), dim3(256), 0, stream,
                                pdl, p, base, ex, iters));
        long long blocks = (n_q + kOwnedTile - 1) / kOwnedTile, cap = 4 * ctx->exchange_max_blocks;
        if (blocks > cap) blocks = cap;
        SLM_CUDA(slm_launch(tc_refine_owned_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, pdl, p, base, ex));
        ctx->launches += 1;
        if (phase_out) *phase_out = 1;
    } else if (p.cpg * p.n_epochs * 4 <= 32) {   // few candidates per query (many queries, short train sets): 8 lanes each
        // sharded path with many queries: 8 groups of 32 queries per block (one fence + one counter update per 256 queries)
        const int iters = (exchange && n_q >= 65536) ? kRefineMaxIters : 1;
        SLM_CUDA(slm_launch(tc_refine_kernel<8>, dim3((unsigned)((n_q + 32 * iters - 1) / (32 * iters))), dim3(256), 0, stream, pdl, p,
                            base, reinterpret_cast<unsigned long long *>(keys_out), ex, iters));
    } else {                          // many ranges (long train sets): a full warp per query
        SLM_CUDA(slm_launch(tc_refine_kernel<32>, dim3((unsigned)((n_q + 7) / 8)), dim3(256), 0, stream, pdl, p, base,
                            reinterpret_cast<unsigned long long *>(keys_out), ex, 1));
    }
    ctx->launches += 2;
    return SLM_OK;
}

}  // namespace

int slm_tc_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                     uint64_t *keys_out, cudaStream_t stream, bool fp4)
{
    TcParams p{};
    p.q = q; p.t = t; p.desc = nullptr; p.pairs = nullptr; p.frame_words = 0;
    p.nq = (int)nq; p.nt = (int)nt;
    return tc_run(ctx, p, 1, base, keys_out, stream, nullptr, nullptr, fp4);
}

int slm_tc_knn2_exchange(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                         const slm_exchange &ex, cudaStream_t stream, bool fp4, int *phase_out)
{
    TcParams p{};
    p.q = q; p.t = t; p.desc = nullptr; p.pairs = nullptr; p.frame_words = 0;
    p.nq = (int)nq; p.nt = (int)nt;
    return tc_run(ctx, p, 1, base, nullptr, stream, &ex, nullptr, fp4, phase_out);
}

int slm_tc_knn2_keys_batched(slm_ctx *ctx, const uint32_t *desc, int64_t n_per_frame, const int32_t *pairs_dev,
                             int64_t n_pairs, uint64_t *keys_out, cudaStream_t stream, const slm_chain *chain, bool fp4)
{
    TcParams p{};
    p.q = nullptr; p.t = nullptr; p.desc = desc; p.pairs = pairs_dev;
    p.frame_words = n_per_frame * 8;
    p.nq = (int)n_per_frame; p.nt = (int)n_per_frame;
    return tc_run(ctx, p, (int)n_pairs, 0, keys_out, stream, nullptr, chain, fp4);
}
