// knn2_frame.cu -- frame-to-frame shapes (about 1000 x 1000 descriptors) in ONE launch.
//
// This is the shape the reference itself runs: tracking.get_matches (tracking.py:12-34), keypoint.py:35-66 and
// Point3D.py:33-54 match the ~1000-2000 ORB descriptors of one frame against those of another and then walk the
// rows with the 0.7 ratio test.  A million comparisons are ~2 us of integer-pipe work, so the cost of the general
// path (distance kernel + refine/merge kernel + finalize kernel, 3 launches, ~30-40 us per synchronous call) is all
// fixed overhead.  Here the whole call -- distances, per-query best two, cross-split merge, ratio test and
// (optionally) the cross-check -- is one kernel:
//   * work item = (group of 32*KQ queries, slice of the train rows): ~one CTA per SM even for a 1000-query
//     problem (1000 x 1000 -> 16 groups x 9 slices = 144 CTAs);
//   * lane = query: every lane keeps KQ query descriptors in registers; the warps of a CTA stride over the CTA's
//     train rows, which sit in shared memory and are read as broadcast LDS.128 (LOP3 XOR + POPC, packed 32-bit key
//     (distance << 23 | train row), top-2 = min/max/min -- unsigned key order is OpenCV's (distance, trainIdx)
//     order, so the lowest train index wins ties);
//   * the warps merge through shared memory; the slices of a group merge through an L2-resident partial list and
//     an atomic ticket per group: the CTA that takes the group's last ticket merges, applies the integer ratio
//     test and writes idx / dist / accept -- no second kernel;
//   * cross-check: a second set of work items runs the same search with the roles swapped; both directions leave
//     packed keys in global memory and the CTA that finishes the last GROUP (second ticket) applies ratio +
//     mutual-best to all queries.
// An earlier version merged the slices of a group through distributed shared memory of an 8-CTA cluster.  It was
// dropped after measurement: with 128 CTAs in 8-clusters the hardware co-locates two CTAs on some SMs while ~26 SMs
// idle (profiles/r1_cluster_place.txt), and since every CTA is bound by its SM's POPC pipe those SMs doubled the
// kernel time (profiles/r1_calib_frame_cluster_version.txt).  Plain CTAs are placed one per SM up to 148.
// Bound: the quarter-rate POPC pipe (8 POPC32 per comparison), like variant P.
#include "slm_internal.cuh"

namespace {

constexpr int kFrameTileRows = 512;        // train rows per shared-memory tile (16 KB)
constexpr int kFrameIdxBits = 23;
constexpr unsigned kFrameNone = 0xFFFFFFFFu;
constexpr unsigned kFrameIdxMask = (1u << kFrameIdxBits) - 1u;
constexpr int kFrameMaxSplits = 64;

struct FrameDir {
    const uint32_t *q;      // "query" side of this direction
    const uint32_t *t;      // "train" side
    int nq, nt;
    int groups;             // ceil(nq / (32 KQ))
    int splits;             // train slices per group
    int rows_per_split;     // ceil(nt / splits)
};

struct FrameParams {
    FrameDir dir[2];
    int items0;                       // work items of direction 0 (= groups * splits); the rest belong to direction 1
    int cross;                        // 1: two directions + last-group finalize
    long long base;                   // global index of train row 0 (direction 0)
    int ratio_num, ratio_den;
    uint2 *part;                      // [item][32 KQ] per-slice (best, second) local keys
    unsigned *tickets;                // one per group of either direction; zero between launches
    unsigned long long *keys_out;     // uint64[nq][2] (optional unless cross)
    unsigned long long *rev_keys;     // uint64[nt][2] (cross only)
    int *idx_out, *dist_out;
    unsigned char *accept_out;
    unsigned *done_counter;           // zero between launches (cross only)
    // reverse-check mode (fwd_keys != nullptr): "query" i of direction 0 is train row best(i) of a finished forward
    // search, the "train" side is the query set; the result says whether query i is the best match of its own best
    // match, and the final stage writes idx / dist / accept from the forward keys
    const unsigned long long *fwd_keys;
};

// Ticket with release / acquire semantics at GPU scope.  Pattern (PTX memory model, cumulativity through bar.sync):
// every thread stores its results; __syncthreads(); ONE thread takes the ticket with atom.acq_rel -- the release half
// publishes the whole CTA's stores, the acquire half makes every earlier ticket holder's stores visible -- then
// __syncthreads() again before the other threads read.  Replaces two MEMBAR.GPU per CTA (ncu: 'membar' was the
// second-largest stall reason of the 1000 x 1000 call).
__device__ __forceinline__ unsigned take_ticket(unsigned *counter)
{
    unsigned old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    return old;
}

__device__ __forceinline__ void top2_insert(unsigned &k1, unsigned &k2, unsigned key)
{
    const unsigned m = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, m);
}

__device__ __forceinline__ unsigned long long widen_key(unsigned k, long long base)
{
    return k == kFrameNone ? kKeyNone
                           : ((unsigned long long)(k >> kFrameIdxBits) << 32) | (unsigned long long)(base + (k & kFrameIdxMask));
}

__device__ __forceinline__ void write_result(unsigned long long k1, unsigned long long k2, long long i, int ratio_num,
                                             int ratio_den, bool mutual, int *idx_out, int *dist_out,
                                             unsigned char *accept_out)
{
    const bool has1 = k1 != kKeyNone, has2 = k2 != kKeyNone;
    const int i1 = has1 ? (int)(k1 & 0xFFFFFFFFull) : -1, d1 = has1 ? (int)(k1 >> 32) : -1;
    const int i2 = has2 ? (int)(k2 & 0xFFFFFFFFull) : -1, d2 = has2 ? (int)(k2 >> 32) : -1;
    if (idx_out) reinterpret_cast<int2 *>(idx_out)[i] = make_int2(i1, i2);
    if (dist_out) reinterpret_cast<int2 *>(dist_out)[i] = make_int2(d1, d2);
    if (accept_out) {
        const bool ok = ratio_num > 0 ? (has1 && has2 && (long long)ratio_den * d1 < (long long)ratio_num * d2) : has1;
        accept_out[i] = (ok && mutual) ? 1 : 0;
    }
}

template <int KQ, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) knn2_frame_kernel(const FrameParams p)
{
    constexpr int QC = 32 * KQ;
    constexpr int kThreads = WARPS * 32;
    static_assert(QC <= kThreads, "one merging thread per query of the CTA");
    __shared__ __align__(16) uint4 tile[kFrameTileRows * 2];
    __shared__ uint2 part[WARPS][QC];
    __shared__ int is_last;

    const int dirn = (int)blockIdx.x >= p.items0 ? 1 : 0;
    const FrameDir d = p.dir[dirn];
    const int item = (int)blockIdx.x - (dirn ? p.items0 : 0);
    const int group = item / d.splits, split = item - group * d.splits;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint32_t qr[KQ][8];
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
        const int qi = min(group * QC + k * 32 + lane, d.nq - 1);
        long long qrow = qi;
        if (p.fwd_keys) {
            const unsigned long long fk = p.fwd_keys[2 * (long long)qi];
            qrow = fk == kKeyNone ? 0 : (long long)(fk & 0xFFFFFFFFull) - p.base;
        }
        const uint4 *src = reinterpret_cast<const uint4 *>(d.q + qrow * 8);
        const uint4 lo = __ldg(src), hi = __ldg(src + 1);
        qr[k][0] = lo.x; qr[k][1] = lo.y; qr[k][2] = lo.z; qr[k][3] = lo.w;
        qr[k][4] = hi.x; qr[k][5] = hi.y; qr[k][6] = hi.z; qr[k][7] = hi.w;
    }
    unsigned b1[KQ], b2[KQ];
#pragma unroll
    for (int k = 0; k < KQ; ++k) { b1[k] = kFrameNone; b2[k] = kFrameNone; }

    const int row_begin = split * d.rows_per_split;
    const int n_rows = min(d.nt, row_begin + d.rows_per_split) - row_begin;   // may be <= 0 for trailing slices
    const uint4 *tsrc = reinterpret_cast<const uint4 *>(d.t) + (long long)row_begin * 2;
    for (int t0 = 0; t0 < n_rows; t0 += kFrameTileRows) {
        const int rows = min(kFrameTileRows, n_rows - t0);
        if (t0 > 0) __syncthreads();
#pragma unroll 4
        for (int i = tid; i < rows * 2; i += kThreads) tile[i] = __ldg(tsrc + (long long)t0 * 2 + i);
        __syncthreads();
        const unsigned j0 = (unsigned)(row_begin + t0);
#pragma unroll 4
        for (int r = warp; r < rows; r += WARPS) {
            const uint4 a = tile[2 * r], b = tile[2 * r + 1];
#pragma unroll
            for (int k = 0; k < KQ; ++k) {
                const unsigned dist = __popc(qr[k][0] ^ a.x) + __popc(qr[k][1] ^ a.y) + __popc(qr[k][2] ^ a.z) +
                                      __popc(qr[k][3] ^ a.w) + __popc(qr[k][4] ^ b.x) + __popc(qr[k][5] ^ b.y) +
                                      __popc(qr[k][6] ^ b.z) + __popc(qr[k][7] ^ b.w);
                top2_insert(b1[k], b2[k], (dist << kFrameIdxBits) + (j0 + (unsigned)r));
            }
        }
    }

    // warps -> one (best, second) pair per query of this CTA
#pragma unroll
    for (int k = 0; k < KQ; ++k) part[warp][k * 32 + lane] = make_uint2(b1[k], b2[k]);
    __syncthreads();
    unsigned k1 = kFrameNone, k2 = kFrameNone;
    if (tid < QC) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint2 v = part[w][tid];
            top2_insert(k1, k2, v.x);
            top2_insert(k1, k2, v.y);
        }
    }
    // slices -> final pair: every CTA publishes its list, the CTA with the group's last ticket merges them
    if (d.splits > 1) {
        uint2 *mine = p.part + ((long long)blockIdx.x) * QC;
        if (tid < QC) mine[tid] = make_uint2(k1, k2);
        __syncthreads();
        unsigned *ticket = p.tickets + (dirn ? p.dir[0].groups : 0) + group;
        if (tid == 0) {
            is_last = take_ticket(ticket) == (unsigned)d.splits - 1u;
            if (is_last) *ticket = 0u;        // ready for the next launch
        }
        __syncthreads();
        if (!is_last) return;
        if (tid < QC) {
            const uint2 *first = p.part + ((long long)blockIdx.x - split) * QC + tid;
            k1 = kFrameNone; k2 = kFrameNone;
            for (int s = 0; s < d.splits; ++s) {
                const uint2 v = __ldcg(first + (long long)s * QC);
                top2_insert(k1, k2, v.x);
                top2_insert(k1, k2, v.y);
            }
        }
    }
    if (tid < QC) {
        const int qi = group * QC + tid;
        if (qi < d.nq && p.fwd_keys) {
            const ulonglong2 fk = reinterpret_cast<const ulonglong2 *>(p.fwd_keys)[qi];
            const bool mutual = fk.x != kKeyNone && k1 != kFrameNone && (int)(k1 & kFrameIdxMask) == qi;
            write_result(fk.x, fk.y, qi, p.ratio_num, p.ratio_den, mutual, p.idx_out, p.dist_out, p.accept_out);
        } else if (qi < d.nq) {
            const long long base = dirn == 0 ? p.base : 0;
            const unsigned long long g1 = widen_key(k1, base), g2 = widen_key(k2, base);
            unsigned long long *keys = dirn == 0 ? p.keys_out : p.rev_keys;
            if (keys) reinterpret_cast<ulonglong2 *>(keys)[qi] = make_ulonglong2(g1, g2);
            if (!p.cross) write_result(g1, g2, qi, p.ratio_num, p.ratio_den, true, p.idx_out, p.dist_out, p.accept_out);
        }
    }
    if (!p.cross) return;

    // cross-check: the CTA that finishes the last group of either direction sees every key of both directions
    __syncthreads();
    if (tid == 0) {
        is_last = take_ticket(p.done_counter) == (unsigned)(p.dir[0].groups + p.dir[1].groups) - 1u;
        if (is_last) *p.done_counter = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    const int nq = p.dir[0].nq, nt = p.dir[0].nt;
    for (int i = tid; i < nq; i += kThreads) {
        const ulonglong2 k = __ldcg(reinterpret_cast<const ulonglong2 *>(p.keys_out) + i);
        bool mutual = false;
        if (k.x != kKeyNone) {
            const long long j = (long long)(k.x & 0xFFFFFFFFull) - p.base;
            if (j >= 0 && j < nt) mutual = (long long)(__ldcg(p.rev_keys + 2 * j) & 0xFFFFFFFFull) == i;
        }
        write_result(k.x, k.y, i, p.ratio_num, p.ratio_den, mutual, p.idx_out, p.dist_out, p.accept_out);
    }
}

struct FramePlan {
    int kq = 0;
    int splits[2] = {1, 1};
    long long est_clk = -1;
};

// Cost model.  A CTA is bound by its SM's POPC pipe: 8 POPC32 per comparison at ~15.8 lanes/clk/SM
// (profiles/r1_pipe_rates.txt) = 0.5 clk per comparison, plus ~1200 clk of fixed cost (query / tile loads, merges,
// ticket).  CTAs are placed one per SM up to the SM count, so the kernel takes ceil(CTAs / SMs) CTA-times.
FramePlan frame_plan(const slm_ctx *ctx, int64_t nq, int64_t nt, bool cross)
{
    FramePlan best;
    if (nq < 1 || nt < 1 || nq > (1 << kFrameIdxBits) || nt > (1 << kFrameIdxBits)) return best;
    static const int kSplitChoices[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18, 20, 24, 28, 32, 40, 48, 56, 64};
    for (int kq = 4; kq >= 1; kq >>= 1) {
        const long long qc = 32 * kq;
        for (int s0 : kSplitChoices) {
            if (s0 > 1 && (nt + s0 - 1) / s0 < 16) break;           // slices shorter than 16 rows are all overhead
            const long long rps0 = (nt + s0 - 1) / s0;
            long long ctas = ((nq + qc - 1) / qc) * s0, worst = rps0;
            int s1 = 1;
            if (cross) {
                // reverse direction: same slice length, so both kinds of CTA cost the same
                long long want = (nq + rps0 - 1) / rps0;
                s1 = (int)(want < 1 ? 1 : (want > kFrameMaxSplits ? kFrameMaxSplits : want));
                const long long rps1 = (nq + s1 - 1) / s1;
                ctas += ((nt + qc - 1) / qc) * s1;
                if (rps1 > worst) worst = rps1;
            }
            const long long tiles = (worst + kFrameTileRows - 1) / kFrameTileRows;
            const long long cta_clk = qc * worst / 2 + 1200 + 700 * (tiles - 1) + 20 * (s0 > s1 ? s0 : s1);
            const long long est = cta_clk * ((ctas + ctx->sm_count - 1) / ctx->sm_count);
            if (best.est_clk < 0 || est < best.est_clk) {
                best.kq = kq; best.splits[0] = s0; best.splits[1] = s1; best.est_clk = est;
            }
        }
    }
    return best;
}

// One-entry plan cache: consecutive calls usually repeat the shape (the search above costs a few microseconds).
const FramePlan &cached_plan(slm_ctx *ctx, int64_t nq, int64_t nt, bool cross)
{
    static thread_local struct { const slm_ctx *ctx; int64_t nq, nt; bool cross; int warps; FramePlan plan; } c = {};
    if (c.ctx != ctx || c.nq != nq || c.nt != nt || c.cross != cross || c.warps != ctx->frame_warps) {
        c.plan = frame_plan(ctx, nq, nt, cross);
        c.ctx = ctx; c.nq = nq; c.nt = nt; c.cross = cross; c.warps = ctx->frame_warps;
    }
    return c.plan;
}

template <int WARPS>
cudaError_t launch_frame(int kq, unsigned grid, const FrameParams &p, cudaStream_t stream)
{
    if (kq == 4) knn2_frame_kernel<4, WARPS><<<grid, WARPS * 32, 0, stream>>>(p);
    else if (kq == 2) knn2_frame_kernel<2, WARPS><<<grid, WARPS * 32, 0, stream>>>(p);
    else knn2_frame_kernel<1, WARPS><<<grid, WARPS * 32, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

// SLM_VARIANT_AUTO prefers this kernel while its estimated time stays below ctx->frame_max_clk: beyond that the
// tensor-pipe variant's ~20 us of fixed cost per search (two searches with cross-check) is repaid by its 15x higher
// rate.  Measured crossover (scripts/calib_frame.py -> profiles/r1_calib_frame.txt): ~26k clocks, ~52k with cross-check.
bool slm_frame_eligible(slm_ctx *ctx, int64_t nq, int64_t nt, bool cross)
{
    if (ctx->frame_max_clk <= 0) return false;
    const FramePlan &pl = cached_plan(ctx, nq, nt, cross);
    return pl.est_clk >= 0 && pl.est_clk <= (cross ? 2 : 1) * ctx->frame_max_clk;
}

static int frame_run(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                     int32_t ratio_num, int32_t ratio_den, bool cross, const uint64_t *fwd_keys, uint64_t *keys_out,
                     int32_t *idx_out, int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream)
{
    const FramePlan pl = cached_plan(ctx, nq, nt, cross);
    if (pl.est_clk < 0) return slm_fail(SLM_ERR_UNSUPPORTED, "shape %lld x %lld is outside the frame kernel's range",
                                        (long long)nq, (long long)nt);
    ctx->last_variant = SLM_VARIANT_POPC;
    ctx->last_kernel = "knn2_frame_kernel";
    FrameParams p{};
    p.cross = cross ? 1 : 0;
    p.base = base;
    p.ratio_num = ratio_num; p.ratio_den = ratio_den;
    p.keys_out = reinterpret_cast<unsigned long long *>(keys_out);
    p.idx_out = idx_out; p.dist_out = dist_out; p.accept_out = accept_out;
    p.fwd_keys = reinterpret_cast<const unsigned long long *>(fwd_keys);
    const int qc = 32 * pl.kq;
    const int s0 = pl.splits[0], s1 = pl.splits[1];
    p.dir[0] = FrameDir{q, t, (int)nq, (int)nt, (int)((nq + qc - 1) / qc), s0, (int)((nt + s0 - 1) / s0)};
    p.dir[1] = FrameDir{t, q, (int)nt, (int)nq, (int)((nt + qc - 1) / qc), s1, (int)((nq + s1 - 1) / s1)};
    p.items0 = p.dir[0].groups * s0;
    long long items = p.items0, groups = p.dir[0].groups;
    if (cross) {
        items += (long long)p.dir[1].groups * s1;
        groups += p.dir[1].groups;
        if (!p.keys_out) {
            SLM_TRY(slm_buf_reserve(ctx, &ctx->keys, (size_t)nq * 16));
            p.keys_out = reinterpret_cast<unsigned long long *>(ctx->keys.p);
        }
        SLM_TRY(slm_buf_reserve(ctx, &ctx->rev, (size_t)nt * 16));
        p.rev_keys = reinterpret_cast<unsigned long long *>(ctx->rev.p);
        p.done_counter = ctx->done_counter;
    }
    if (items > 0x7FFFFFFFll) return slm_fail(SLM_ERR_UNSUPPORTED, "problem too large for one launch");
    SLM_TRY(slm_buf_reserve(ctx, &ctx->scratch, (size_t)items * qc * sizeof(uint2)));
    p.part = reinterpret_cast<uint2 *>(ctx->scratch.p);
    // group tickets: a zero-initialised array that every launch leaves zeroed again
    if ((size_t)groups * sizeof(unsigned) > ctx->tickets.bytes) {
        SLM_TRY(slm_buf_reserve(ctx, &ctx->tickets, (size_t)groups * sizeof(unsigned)));
        SLM_CUDA(cudaMemsetAsync(ctx->tickets.p, 0, ctx->tickets.bytes, stream));
    }
    p.tickets = reinterpret_cast<unsigned *>(ctx->tickets.p);

    // (the reverse check after a tensor-pipe search is not that call's dominant kernel: no profile bracket)
    if (!fwd_keys) SLM_TRY(slm_prof_begin(ctx, stream));
    cudaError_t e;
    if (ctx->frame_warps == 16) e = launch_frame<16>(pl.kq, (unsigned)items, p, stream);
    else if (ctx->frame_warps == 4) e = launch_frame<4>(pl.kq, (unsigned)items, p, stream);
    else e = launch_frame<8>(pl.kq, (unsigned)items, p, stream);
    if (e != cudaSuccess) return slm_fail(SLM_ERR_CUDA, "frame kernel launch failed: %s", cudaGetErrorString(e));
    if (!fwd_keys) SLM_TRY(slm_prof_end(ctx, stream));
    ctx->launches += 1;
    return SLM_OK;
}

int slm_frame_knn2(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                   int32_t ratio_num, int32_t ratio_den, int32_t cross_check, uint64_t *keys_out, int32_t *idx_out,
                   int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream)
{
    return frame_run(ctx, q, nq, t, nt, base, ratio_num, ratio_den, cross_check != 0 && accept_out != nullptr, nullptr,
                     keys_out, idx_out, dist_out, accept_out, stream);
}

// Cross-check after a forward search that ran elsewhere (tensor pipe): only the <= nq train rows that ARE somebody's
// best match can be mutual, so the reverse search is best(i) x all queries -- nq x nq instead of nt x nq -- and
// for frame-sized nq it is this kernel, which also applies the ratio test and writes the final outputs.
int slm_frame_revcheck(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t base, int32_t ratio_num,
                       int32_t ratio_den, const uint64_t *fwd_keys, int32_t *idx_out, int32_t *dist_out,
                       uint8_t *accept_out, cudaStream_t stream)
{
    return frame_run(ctx, t, nq, q, nq, base, ratio_num, ratio_den, false, fwd_keys, nullptr, idx_out, dist_out,
                     accept_out, stream);
}
