// knn2_tc4.cu -- variant T4: exact Hamming 2-NN on the block-scaled FP4 tensor path (tcgen05.mma kind::mxf4).
//
// Replaces the O(nq*nt) distance evaluation inside matcher.knnMatch(des1, des2, k=2)
// (reference call sites tracking.py:22, keypoint.py:44, Point3D.py:40; exhaustive semantics).
//
// Same contraction as variant T (knn2_tc.cu) -- every descriptor bit becomes +1 / -1, dot = 256 - 2 * Hamming exactly --
// but the operands are e2m1 (4 bits per value) and the MMA is kind::mxf4.block_scale with every UE8M0 scale = 1.0:
// twice the fp8 rate (measured 16 381 MAC/clk/SM, profiles/r2_fp4_probe.txt) from HALF the shared-memory bytes.
//
// One cluster of 2 CTAs (cta_group::2, M = 256) per SM pair, 384 threads per CTA, 1 CTA / SM:
//   warps 0-7   epilogue: TMEM -> registers (tcgen05.ld 32x32b x64/x32/x16/x8), branch-free candidate-chunk tracking
//   warps 8-10  expanders: raw packed train rows (shared memory, landed by TMA bulk copies) -> +-1 e2m1 operand rows
//               in the K-major no-swizzle UMMA layout; lane 0 of warp 8 is also the TMA producer
//               (cp.async.bulk global -> shared + mbarrier complete_tx, 4-stage ring of raw half tiles)
//   warp  11    leader CTA: one elected thread issues tcgen05.mma.cta_group::2.kind::mxf4 (N = 240, K = 64 x 4)
//               into two 240-column TMEM accumulators and tcgen05.commit's onto mbarriers
// TMEM: [0,240) [240,480) accumulators, [480,512) scale factors (all bytes 0x7F).  A job = 256 queries x 240 train rows;
// each CTA keeps up to 8 query tiles (128 rows x 128 B) resident, so BASELINE config 5's 2000 queries are ONE cluster
// group: every train tile is expanded exactly once on the whole GPU and serves 8 jobs.
//
// Epilogue: thread = one query row = one TMEM lane; warps 0-3 own columns [0,120) of the accumulator, warps 4-7
// [120,240).  Per row only the best two candidate CHUNKS (CH = 120 or 40 train rows) are tracked, packed in one fp32
// (tc_params.cuh); the exact top-2 rows lie inside the best two chunks, which tc_refine_kernel re-scores with XOR+POPC.
#include <atomic>
#include <cfloat>
#include <type_traits>

#include "slm_internal.cuh"
#include "exchange.cuh"
#include "tc_common.cuh"
#include "tc_ld.cuh"
#include "tc_params.cuh"

namespace {

using namespace tcp;

constexpr int kEpiWarps = 8;
constexpr int kExpWarps = 3;
constexpr int kEpiThreads = kEpiWarps * 32;          // 256
constexpr int kExpThreads = kExpWarps * 32;          // 96
constexpr int kThreads4 = kEpiThreads + kExpThreads + 32;   // 384: 12 warps keep 168 registers per thread
constexpr int kBStages = 3;                          // expanded half tiles (15 KB each)
constexpr int kRawStages = 4;                        // raw packed half tiles (3840 B each)
constexpr uint32_t kRawBytes = tc4::kHalfN * 32;
constexpr int kHalf = tc4::kHalfN;                   // 120
#ifndef SLM_TC4_FENCE_BEFORE_HANDBACK
#define SLM_TC4_FENCE_BEFORE_HANDBACK 1              // A/B switch (compile time): no measurable cost, kept
#endif

struct Bars4 {
    uint64_t a_full;
    uint64_t raw_full[kRawStages];
    uint64_t b_full[kBStages], b_empty[kBStages];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// ---- TMA bulk copy (non-tensor form): global -> shared, completion counted in bytes on an mbarrier ----
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_sync_expanders() { asm volatile("bar.sync 1, %0;" ::"n"(kExpThreads) : "memory"); }

// The tile sequence every role of a cluster walks: for every link of the unit (one link unless the batch is chained: the
// train frames of the pairs that share this cluster's query frame), ranges unit, unit + cpg, ... of up to range_tiles tiles.
struct TileIter {
    int link, r, bt, n_bt;
    int unit, n_links;
    bool done;
};
__device__ __forceinline__ void iter_start(TileIter &it, const TcParams &p, int unit, int total_tiles, int n_links = 1)
{
    it.link = 0;
    it.unit = unit;
    it.n_links = n_links;
    it.r = unit;
    it.bt = 0;
    it.done = unit >= p.n_ranges || n_links <= 0;
    it.n_bt = it.done ? 0 : min(p.range_tiles, total_tiles - unit * p.range_tiles);
}
__device__ __forceinline__ void iter_next(TileIter &it, const TcParams &p, int total_tiles)
{
    if (++it.bt == it.n_bt) {
        it.bt = 0;
        it.r += p.cpg;
        if (it.r >= p.n_ranges) {
            if (++it.link >= it.n_links) { it.done = true; return; }
            it.r = it.unit;
        }
        it.n_bt = min(p.range_tiles, total_tiles - it.r * p.range_tiles);
    }
}

// mbarrier wait of the hot loop: no spin counter, no printf path (the bounded tc::mbar_wait costs ~10 extra instructions and
// a call frame per job).  Every barrier of this kernel is also waited for by a bounded wait somewhere (expanders, MMA
// issuer), so a protocol bug still ends in their trap rather than in a hang.
__device__ __forceinline__ void mbar_wait_lean(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra WAIT_%=;\n\t}" ::"r"(tc::smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void pin_timing(float &x) { asm volatile("" : "+f"(x) :: "memory"); }

template <int C>
__device__ __forceinline__ float max_cols(const uint32_t *v)
{
    // maximum of C columns: four independent chains of 3-input max (FMNMX3), ~C / 2 instructions
    float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
    for (int j = 4; j + 7 < C; j += 8) {
        m0 = fmaxf(fmaxf(m0, __uint_as_float(v[j + 0])), __uint_as_float(v[j + 1]));
        m1 = fmaxf(fmaxf(m1, __uint_as_float(v[j + 2])), __uint_as_float(v[j + 3]));
        m2 = fmaxf(fmaxf(m2, __uint_as_float(v[j + 4])), __uint_as_float(v[j + 5]));
        m3 = fmaxf(fmaxf(m3, __uint_as_float(v[j + 6])), __uint_as_float(v[j + 7]));
    }
    // C = 8 k: four columns left
    m0 = fmaxf(fmaxf(m0, __uint_as_float(v[C - 4])), __uint_as_float(v[C - 3]));
    m1 = fmaxf(fmaxf(m1, __uint_as_float(v[C - 2])), __uint_as_float(v[C - 1]));
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// TIMING (diagnostic build of the same kernel, SLM_TC4_TIMING=1): two epilogue warps and the MMA warp of cluster 0 sum
// clock64 intervals per phase and write them to p.timing -- where a job's 597 clk go.
// CHAIN (batches of frame-sized problems, config 3): grid.y indexes UNITS = runs of pairs that share the query frame; the
// cluster expands that frame's query tiles once and walks the train frames of the run back to back, one candidate flush
// per pair.  CH = 20 there: the refine pass re-scores two chunks per query and is the second-largest cost of such a batch.
// maximum of 20 columns in 10 instructions (four chains of two FMNMX3, then FMNMX3 + FMNMX)
__device__ __forceinline__ float max20(const uint32_t *v)
{
    auto f = [&](int i) { return __uint_as_float(v[i]); };
    float m0 = fmaxf(fmaxf(f(0), f(4)), f(5)), m1 = fmaxf(fmaxf(f(1), f(6)), f(7));
    float m2 = fmaxf(fmaxf(f(2), f(8)), f(9)), m3 = fmaxf(fmaxf(f(3), f(10)), f(11));
    m0 = fmaxf(fmaxf(m0, f(12)), f(13));
    m1 = fmaxf(fmaxf(m1, f(14)), f(15));
    m2 = fmaxf(fmaxf(m2, f(16)), f(17));
    m3 = fmaxf(fmaxf(m3, f(18)), f(19));
    return fmaxf(fmaxf(fmaxf(m0, m1), m2), m3);
}

template <int CH, bool CHAIN, bool TIMING>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads4, 1) knn2_tc4_kernel(TcParams p)
{
    static_assert(CH == 120 || CH == 40 || CH == 20, "candidate chunk width");
    constexpr int kCPT = tc4::kTileN / CH;             // chunks per tile: 2 or 6
    constexpr int kCPS = kCPT / 2;                     // chunks per epilogue set: 1 or 3
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                                        // p.mt tiles of 16 KB
    uint8_t *sB = smem + (size_t)p.mt * tc4::kATileBytes;                      // kBStages half tiles of 15 KB
    uint8_t *sRaw = sB + kBStages * tc4::kBHalfBytes;                          // kRawStages raw half tiles
    Bars4 *bars = reinterpret_cast<Bars4 *>(sRaw + kRawStages * kRawBytes);

    slm_pdl_launch_dependents();       // the refine kernel may start launching; it waits for this grid's results
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);   // warp-uniform for the compiler: TMEM addresses live in uniform registers
    const uint32_t rank = tc::cluster_ctarank();          // 0 = leader
    const int item = blockIdx.x >> 1;                     // cluster index
    const int gpair = item / p.cpg;                       // pair of query groups served by this cluster
    const int unit = item % p.cpg;                        // this cluster walks ranges unit, unit + cpg, ...
    const int group = gpair * 2 + (int)rank;              // may be == n_groups (idle half of an odd pair)
    const uint32_t *q = p.q;
    int link_first = 0, n_links = 1;              // chained batch: sorted pairs [link_first, link_first + n_links)
    if constexpr (CHAIN) {
        const int2 cu = reinterpret_cast<const int2 *>(p.chain_units)[blockIdx.y];
        link_first = cu.x;
        n_links = cu.y;
        q = p.desc + (long long)p.chain_pairs[2 * link_first] * p.frame_words;
    }
    // train rows / output slot of link l (the single problem when not chained)
    auto link_train = [&](int l) -> const uint32_t * {
        if constexpr (CHAIN) return p.desc + (long long)p.chain_pairs[2 * (link_first + l) + 1] * p.frame_words;
        else return p.t;
    };
    auto link_prob = [&](int l) -> int {
        if constexpr (CHAIN) return p.chain_prob[link_first + l];
        else return (int)blockIdx.y;
    };

    const int MT = p.mt;
    const int q_first = group * (MT * kTileM);
    const int mt_mine = max(0, min(MT, (p.nq - q_first + kTileM - 1) / kTileM));
    const int mt_pair = min(MT, (p.nq - gpair * 2 * (MT * kTileM) + kTileM - 1) / kTileM);   // leader's count
    const int total_tiles = (p.nt + tc4::kTileN - 1) / tc4::kTileN;

    if (tid == 0) {
        tc::mbar_init(&bars->a_full, 2 * kThreads4);
        for (int s = 0; s < kRawStages; ++s) tc::mbar_init(&bars->raw_full[s], 1);
        for (int s = 0; s < kBStages; ++s) {
            tc::mbar_init(&bars->b_full[s], 2 * kExpThreads);
            tc::mbar_init(&bars->b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&bars->acc_full[s], 1);
            tc::mbar_init(&bars->acc_empty[s], 2 * kEpiThreads);
        }
        tc::fence_barrier_init();
    }
    const int mma_warp = kEpiWarps + kExpWarps;
    if (warp == mma_warp) tc::tmem_alloc_2cta(&bars->tmem_base, 512);
    tc::tc_fence_before();
    tc::cluster_sync();          // barrier inits + TMEM allocation visible to both CTAs
    tc::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    // The TMA producer (lane 0 of the first expander warp) requests its first raw half tiles NOW, so that the HBM latency of
    // the first train tiles hides behind the expansion of the query tiles (matters on the sharded path, where a launch
    // lasts ~0.2 ms).
    const uint32_t sRaw_addr0 = tc::smem_u32(sRaw);
    auto valid_rows = [&](int tile) { return max(0, min(kHalf, p.nt - (tile * tc4::kTileN + (int)rank * kHalf))); };
    auto produce = [&](const TileIter &pi_, int slot) {
        const int tile = pi_.r * p.range_tiles + pi_.bt;
        const int v = valid_rows(tile);
        if (v > 0) {
            mbar_arrive_expect_tx(&bars->raw_full[slot], (uint32_t)v * 32);
            bulk_copy_g2s(sRaw_addr0 + (uint32_t)slot * kRawBytes,
                          link_train(pi_.link) + ((long long)tile * tc4::kTileN + (long long)rank * kHalf) * 8, (uint32_t)v * 32,
                          &bars->raw_full[slot]);
        } else {
            tc::mbar_arrive(&bars->raw_full[slot]);
        }
    };
    TileIter pi;                                             // producer position: kRawStages tiles ahead of the expanders
    iter_start(pi, p, unit, total_tiles, n_links);
    if (tid == kEpiThreads) {
        for (int s = 0; s < kRawStages && !pi.done; ++s) {
            produce(pi, s);
            iter_next(pi, p, total_tiles);
        }
    }

    // scale factors: every byte of columns [480, 512) = UE8M0 1.0, in both CTAs (4 warps = 128 lanes)
    if (warp < 4) tc4::tmem_fill32(tmem + ((uint32_t)(warp * 32) << 16) + tc4::kSfCol, tc4::kScaleOnes);

    // Query tiles: expanded once per cluster by ALL threads (loads issued first, then the expansion)
    {
        const uint32_t sA_addr = tc::smem_u32(sA);
        const int n_rows = mt_mine * kTileM;
        constexpr int kPer = (kMaxMT4 * kTileM + kThreads4 - 1) / kThreads4;   // 3 rows per thread at most
        uint4 d0[kPer], d1[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int r = tid + i * kThreads4;
            if (r < n_rows) {
                const uint4 *src = reinterpret_cast<const uint4 *>(q + (long long)min(q_first + r, p.nq - 1) * 8);
                d0[i] = __ldg(src);
                d1[i] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int r = tid + i * kThreads4;
            if (r < n_rows) tc4::expand_row_to_smem(sA_addr + (uint32_t)(r / kTileM) * tc4::kATileBytes, r % kTileM, d0[i], d1[i]);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        tc::mbar_arrive_cluster(&bars->a_full, 0);
    }

    if (warp < kEpiWarps) {
        // ===================== epilogue (own TMEM: own 128 query rows x 240 train columns) =====================
        const int set = warp >> 2, quad = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16) + set * kHalf;
        const int n_slots = p.cpg * p.n_epochs;
        auto cand_of = [&](int link) {
            return p.cand + ((long long)link_prob(link) * p.nq) * n_slots * 2 + (long long)(unit * p.n_epochs) * 2 + set;
        };
        float2 *cand = cand_of(0);
        float b1[kMaxMT4], b2[kMaxMT4];
#pragma unroll
        for (int m = 0; m < kMaxMT4; ++m) { b1[m] = -FLT_MAX; b2[m] = -FLT_MAX; }
        auto flush = [&](int epoch) {
#pragma unroll
            for (int m = 0; m < kMaxMT4; ++m) {
                const int qi = q_first + m * kTileM + quad * 32 + lane;
                if (m < mt_mine && qi < p.nq) cand[(long long)qi * n_slots * 2 + epoch * 2] = make_float2(b1[m], b2[m]);
                b1[m] = -FLT_MAX;
                b2[m] = -FLT_MAX;
            }
        };
        auto track = [&](float &t1, float &t2, float key) {
            t2 = fmaxf(t2, fminf(t1, key));
            t1 = fmaxf(t1, key);
        };
        // Jobs per train tile, padded to an even count: the accumulator (job & 1) and the register buffers of the software
        // pipeline then depend on the compile-time m only.  (The padding job of an odd count repeats the last query tile;
        // nobody tracks its result.)
        const int mt_jobs = (mt_pair + 1) & ~1;
        int job = 0, epoch = 0, lt = 0, j = 0, last_r = -1;     // lt = tiles seen in this epoch, j = ranges walked
        long long tm[5] = {0, 0, 0, 0, 0};                      // TIMING: wait full | loads 0+1 | reduce 0 + load 2 | reduce 1+2, track | jobs
        long long tq = 0;
        if constexpr (TIMING) tq = clock64();
        uint32_t X[40], Y[40], Z[40];        // three 40-column register buffers
        auto valid_cols_of = [&](const TileIter &ti) { return min(tc4::kTileN, p.nt - (ti.r * p.range_tiles + ti.bt) * tc4::kTileN); };
        int cur_link = 0;
        TileIter it;
        for (iter_start(it, p, unit, total_tiles, n_links); !it.done; iter_next(it, p, total_tiles), ++lt) {
            if (CHAIN && it.link != cur_link) {
                // next pair of the chain: flush this pair's candidates, start over for the next output slot
                for (; epoch < p.n_epochs; ++epoch) flush(epoch);
                cur_link = it.link;
                cand = cand_of(cur_link);
                epoch = 0; lt = 0; j = 0; last_r = -1;
            }
            if (it.r != last_r) {
                if (last_r >= 0 && ++j % p.rpe == 0) { flush(epoch); ++epoch; lt = 0; }
                last_r = it.r;
            }
            const int valid_cols = valid_cols_of(it);
            const int col0 = set * kHalf;
            const bool tile_active = col0 < valid_cols;
            // chunk k of this set covers columns [set * 120 + k * CH, + CH); counter of the first one
            const float bias0 = (float)(kChunkMask - (lt * kCPT + set * kCPS));
            auto one_job = [&](auto m_const) {
                constexpr int m = decltype(m_const)::value;
                if (m < mt_jobs) {
                    constexpr int ab = m & 1;
                    const uint32_t ph = (uint32_t)(job >> 1) & 1u;
                    const uint32_t acc = lane_addr + ab * tc4::kTileN;
                    const bool active = tile_active && m < mt_mine;
                    if (active) {
                        // All 120 columns of this warp are requested at once (three 40-column pieces, 120 registers); they land
                        // within ~40 clk and the accumulator is handed back BEFORE any reduction: with only two accumulators
                        // the chain MMA -> commit -> TMEM read -> hand-back -> next MMA bounds the kernel, and the diagnostic
                        // build (SLM_TC4_TIMING, profiles/r2_tc4_timing_*.txt) showed the reduction of piece 0 (190 clk with
                        // both warps of a scheduler on the ALU pipe) sitting inside that chain.
                        mbar_wait_lean(&bars->acc_full[ab], ph);
                        tc::tc_fence_after();
                        if constexpr (TIMING) { const long long c = clock64(); tm[0] += c - tq; tq = c; }
                        tc::tmem_ldx32(acc, X);
                        tc::tmem_ldx8(acc + 32, X + 32);
                        tc::tmem_ldx32(acc + 40, Y);
                        tc::tmem_ldx8(acc + 72, Y + 32);
                        tc::tmem_ldx32(acc + 80, Z);
                        tc::tmem_ldx8(acc + 112, Z + 32);
                        tc::tmem_wait_ld();
                        tc::tmem_pin32(X);
                        tc::tmem_pin8(X + 32);
                        tc::tmem_pin32(Y);
                        tc::tmem_pin8(Y + 32);
                        tc::tmem_pin32(Z);
                        tc::tmem_pin8(Z + 32);
                        if constexpr (TIMING) { const long long c = clock64(); tm[1] += c - tq; tq = c; }
                        // One unconditional arrival per THREAD.  Measured alternatives, all 7-10 % slower per tile although they
                        // send fewer / cheaper messages (any branch around the arrival changes ptxas's schedule of the
                        // hand-back path): one arrival per warp behind __syncwarp; local arrivals for the leader CTA's own
                        // threads.  Dropping the tcgen05.fence in front of it changes nothing (profiles/r2_tc4_epilogue_steps.txt).
                        if (SLM_TC4_FENCE_BEFORE_HANDBACK) tc::tc_fence_before();
                        tc::mbar_arrive_cluster_relaxed(&bars->acc_empty[ab], 0);
                        if constexpr (TIMING) { const long long c = clock64(); tm[2] += c - tq; tq = c; }
                        const float mx0 = fmaxf(max_cols<32>(X), max_cols<8>(X + 32));
                        const float mx1 = fmaxf(max_cols<32>(Y), max_cols<8>(Y + 32));
                        const float mx2 = fmaxf(max_cols<32>(Z), max_cols<8>(Z + 32));
                        if constexpr (CH == 120) {
                            track(b1[m], b2[m], fmaf(fmaxf(fmaxf(mx0, mx1), mx2), kKeyScale, bias0));
                        } else if constexpr (CH == 20) {
                            // six 20-column chunks per warp and job: the halves of the three pieces
                            track(b1[m], b2[m], fmaf(max20(X), kKeyScale, bias0));
                            if (col0 + 20 < valid_cols) track(b1[m], b2[m], fmaf(max20(X + 20), kKeyScale, bias0 - 1.0f));
                            if (col0 + 40 < valid_cols) track(b1[m], b2[m], fmaf(max20(Y), kKeyScale, bias0 - 2.0f));
                            if (col0 + 60 < valid_cols) track(b1[m], b2[m], fmaf(max20(Y + 20), kKeyScale, bias0 - 3.0f));
                            if (col0 + 80 < valid_cols) track(b1[m], b2[m], fmaf(max20(Z), kKeyScale, bias0 - 4.0f));
                            if (col0 + 100 < valid_cols) track(b1[m], b2[m], fmaf(max20(Z + 20), kKeyScale, bias0 - 5.0f));
                        } else {
                            track(b1[m], b2[m], fmaf(mx0, kKeyScale, bias0));
                            if (col0 + 40 < valid_cols) track(b1[m], b2[m], fmaf(mx1, kKeyScale, bias0 - 1.0f));
                            if (col0 + 80 < valid_cols) track(b1[m], b2[m], fmaf(mx2, kKeyScale, bias0 - 2.0f));
                        }
                        if constexpr (TIMING) {
                            pin_timing(b1[m]);
                            const long long c = clock64();
                            tm[3] += c - tq;
                            tq = c;
                            tm[4] += 1;
                        }
                    } else {
                        // nothing to read for this warp (idle half of a pair, padding job, columns past the train set)
                        mbar_wait_lean(&bars->acc_full[ab], ph);
                        tc::tc_fence_after();
                        tc::tc_fence_before();
                        tc::mbar_arrive_cluster_relaxed(&bars->acc_empty[ab], 0);
                    }
                    ++job;
                }
            };
            one_job(std::integral_constant<int, 0>{});
            one_job(std::integral_constant<int, 1>{});
            one_job(std::integral_constant<int, 2>{});
            one_job(std::integral_constant<int, 3>{});
            one_job(std::integral_constant<int, 4>{});
            one_job(std::integral_constant<int, 5>{});
            one_job(std::integral_constant<int, 6>{});
            one_job(std::integral_constant<int, 7>{});
            static_assert(kMaxMT4 == 8, "one_job is instantiated for m = 0..7");
        }
        for (; epoch < p.n_epochs; ++epoch) flush(epoch);     // remaining epochs are written as "none"
        if constexpr (TIMING) {
            if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4))
                for (int k = 0; k < 5; ++k) p.timing[(warp >> 2) * 8 + k] = tm[k];
        }
    } else if (warp < mma_warp) {
        // ===================== expanders + TMA producer: this CTA's half of every train tile =====================
        const int et = tid - kEpiThreads;   // 0..95
        const uint32_t sB_addr = tc::smem_u32(sB), sRaw_addr = tc::smem_u32(sRaw);
        const bool two_rows = et < kHalf - kExpThreads;          // threads 0..23 also expand row 96 + et
        int sb = 0, phb = 0, sr = 0, phr = 0;
        TileIter it;
        for (iter_start(it, p, unit, total_tiles, n_links); !it.done; iter_next(it, p, total_tiles)) {
            const int tile = it.r * p.range_tiles + it.bt;
            const int v = valid_rows(tile);
            tc::mbar_wait(&bars->raw_full[sr], phr, 50 + sr);
            // rows past the end of the train set repeat the last valid row of this half: their dot products equal a real
            // row's, so a chunk maximum is never inflated (columns of a half without any valid row are skipped by the epilogue)
            const uint4 *raw = reinterpret_cast<const uint4 *>(sRaw + (size_t)sr * kRawBytes);
            uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0, e0 = c0, e1 = c0;
            if (v > 0) {
                const int r0 = min(et, v - 1), r1 = min(et + kExpThreads, v - 1);
                c0 = raw[2 * r0];
                c1 = raw[2 * r0 + 1];
                if (two_rows) { e0 = raw[2 * r1]; e1 = raw[2 * r1 + 1]; }
            }
            tc::mbar_wait_backoff(&bars->b_empty[sb], phb ^ 1, 20 + sb, 100);
            tc4::expand_row_to_smem(sB_addr + (uint32_t)sb * tc4::kBHalfBytes, et, c0, c1);
            if (two_rows) tc4::expand_row_to_smem(sB_addr + (uint32_t)sb * tc4::kBHalfBytes, et + kExpThreads, e0, e1);
            tc::fence_proxy_async();
            tc::mbar_arrive_cluster(&bars->b_full[sb], 0);
            // every expander has read raw slot sr: refill it with the tile kRawStages ahead
            bar_sync_expanders();
            if (et == 0 && !pi.done) {
                produce(pi, sr);
                iter_next(pi, p, total_tiles);
            }
            if (++sb == kBStages) { sb = 0; phb ^= 1; }
            if (++sr == kRawStages) { sr = 0; phr ^= 1; }
        }
    } else {
        // ===================== MMA issuer: leader CTA; the warp stays converged, one elected lane issues ==========
        if (rank == 0) {
            const uint32_t idesc = tc4::idesc_mxf4(2 * kTileM, tc4::kTileN);
            const uint32_t a_lo0 = tc4::smem_desc_lo(tc::smem_u32(sA)), b_lo0 = tc4::smem_desc_lo(tc::smem_u32(sB));
            const uint32_t sf = tmem + tc4::kSfCol;
            tc::mbar_wait_cluster(&bars->a_full, 0, 30);
            tc::tc_fence_after();
            const int mt_jobs = (mt_pair + 1) & ~1;
            int job = 0, s = 0, ph = 0;
            long long tm[4] = {0, 0, 0, 0};      // TIMING: wait b_full | wait acc_empty | issue | jobs
            long long tq = 0;
            if constexpr (TIMING) tq = clock64();
            TileIter it;
            for (iter_start(it, p, unit, total_tiles, n_links); !it.done; iter_next(it, p, total_tiles)) {
                tc::mbar_wait_cluster(&bars->b_full[s], ph, 31 + s);
                tc::tc_fence_after();
                if constexpr (TIMING) { const long long c = clock64(); tm[0] += c - tq; tq = c; }
                const uint32_t b_lo = b_lo0 + s * (tc4::kBHalfBytes >> 4);
                for (int m = 0; m < mt_jobs; ++m) {          // even: the padding job repeats the last query tile
                    const int ab = job & 1;
                    tc::mbar_wait_cluster(&bars->acc_empty[ab], ((job >> 1) & 1) ^ 1, 40 + ab);
                    tc::tc_fence_after();
                    if constexpr (TIMING) { const long long c = clock64(); tm[1] += c - tq; tq = c; }
                    if (tc::elect_one()) {
                        tc4::umma_job(tmem + ab * tc4::kTileN, a_lo0 + min(m, mt_pair - 1) * (tc4::kATileBytes >> 4), b_lo, idesc, sf);
                        tc::umma_commit_2cta(&bars->acc_full[ab], 3);
                    }
                    __syncwarp();
                    if constexpr (TIMING) { const long long c = clock64(); tm[2] += c - tq; tq = c; tm[3] += 1; }
                    ++job;
                }
                if (tc::elect_one()) tc::umma_commit_2cta(&bars->b_empty[s], 3);
                __syncwarp();
                if (++s == kBStages) { s = 0; ph ^= 1; }
            }
            if constexpr (TIMING) {
                if (blockIdx.x == 0 && lane == 0)
                    for (int k = 0; k < 4; ++k) p.timing[16 + k] = tm[k];
            }
        }
    }

    tc::tc_fence_before();
    tc::cluster_sync();          // nobody exits while the peer may still signal its barriers / read its smem
    if (warp == mma_warp) tc::tmem_dealloc_2cta(tmem, 512);
}

template <int CH, bool CHAIN, bool TIMING>
int launch_tc4(const TcParams &p, int grid_y, cudaStream_t stream)
{
    const size_t smem_max = (size_t)kMaxMT4 * tc4::kATileBytes + kBStages * tc4::kBHalfBytes + kRawStages * kRawBytes + sizeof(Bars4) + 64;
    static std::atomic<bool> configured[64];
    int dev = 0;
    SLM_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63].load(std::memory_order_acquire)) {
        SLM_CUDA(cudaFuncSetAttribute(knn2_tc4_kernel<CH, CHAIN, TIMING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        configured[dev & 63].store(true, std::memory_order_release);
    }
    const size_t smem = (size_t)p.mt * tc4::kATileBytes + kBStages * tc4::kBHalfBytes + kRawStages * kRawBytes + sizeof(Bars4) + 64;
    const int n_gpairs = (p.n_groups + 1) / 2;
    dim3 grid((unsigned)(2 * n_gpairs * p.cpg), (unsigned)grid_y);
    knn2_tc4_kernel<CH, CHAIN, TIMING><<<grid, kThreads4, smem, stream>>>(p);
    SLM_CUDA(cudaGetLastError());
    return SLM_OK;
}

}  // namespace

int slm_tc4_launch(slm_ctx *ctx, const tcp::TcParams &p_in, int grid_y, cudaStream_t stream)
{
    tcp::TcParams p = p_in;
    if (p.mt < 1 || p.mt > kMaxMT4) return slm_fail(SLM_ERR_INVALID, "tc4: %d query tiles per CTA", p.mt);
    if (ctx->tc4_timing && p.chunk == 120 && !p.chain_pairs) {
        // diagnostic: per-phase cycle sums of cluster 0 (synchronous; SLM_TC4_TIMING=1)
        long long *dev = nullptr, h[24] = {};
        SLM_CUDA(cudaMalloc(&dev, sizeof(h)));
        SLM_CUDA(cudaMemset(dev, 0, sizeof(h)));
        p.timing = dev;
        SLM_TRY((launch_tc4<120, false, true>(p, grid_y, stream)));
        SLM_CUDA(cudaStreamSynchronize(stream));
        SLM_CUDA(cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost));
        SLM_CUDA(cudaFree(dev));
        for (int w = 0; w < 2; ++w) {
            const double n = h[w * 8 + 4] > 0 ? (double)h[w * 8 + 4] : 1.0;
            fprintf(stderr, "tc4 timing, epilogue warp %d (set %d): %lld jobs; per job: wait acc_full %.0f | loads 0+1 %.0f | reduce 0 + load 2 + hand-back %.0f | "
                    "reduce 1+2 + track %.0f clk (sum %.0f)\n", w * 4, w, h[w * 8 + 4], h[w * 8] / n, h[w * 8 + 1] / n, h[w * 8 + 2] / n, h[w * 8 + 3] / n,
                    (h[w * 8] + h[w * 8 + 1] + h[w * 8 + 2] + h[w * 8 + 3]) / n);
        }
        const double n = h[19] > 0 ? (double)h[19] : 1.0;
        fprintf(stderr, "tc4 timing, MMA warp: %lld jobs; per job: wait b_full %.0f | wait acc_empty %.0f | issue %.0f clk (sum %.0f)\n", h[19],
                h[16] / n, h[17] / n, h[18] / n, (h[16] + h[17] + h[18]) / n);
        return SLM_OK;
    }
    if (p.chain_pairs) {
        if (p.chunk != 20) return slm_fail(SLM_ERR_INVALID, "tc4: chained batches track 20-row chunks (got %d)", p.chunk);
        return launch_tc4<20, true, false>(p, grid_y, stream);
    }
    switch (p.chunk) {
    case 120: return launch_tc4<120, false, false>(p, grid_y, stream);
    case 40: return launch_tc4<40, false, false>(p, grid_y, stream);
    default: return slm_fail(SLM_ERR_INVALID, "tc4: unsupported candidate chunk width %d", p.chunk);
    }
}
