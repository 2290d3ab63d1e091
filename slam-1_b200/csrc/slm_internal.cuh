// slm_internal.cuh -- shared declarations of libslammatch.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "slammatch.h"

// Packed result key used between kernels: (distance << 32) | global train index.
// Unsigned order of the key == OpenCV's (distance, trainIdx) order.
static constexpr unsigned long long kKeyNone = 0xFFFFFFFFFFFFFFFFull;

static constexpr int kMaxHostChunks = 256;   // x 1M rows per chunk

struct slm_buf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct slm_ctx {
    int device = 0;
    int sm_count = 148;
    int variant = SLM_VARIANT_AUTO;
    int last_variant = 0;
    const char *last_kernel = "";   // name of the distance kernel the last search launched (slm_last_kernel)
    int epoch_tiles = 4096;   // tiles per candidate epoch (SLM_TC_EPOCH_TILES shrinks it so tests reach the epoch logic)
    int max_cpg = 1 << 30;    // cap on clusters per query-group pair (SLM_TC_MAX_CPG, tests only)
    int force_1cta = 0;   // debugging / A-B: run the single-CTA tcgen05 kernel even when CTA pairs apply
    int tc_chain_max = 8;  // batched calls: at most this many pairs of one query frame per cluster (SLM_TC_CHAIN; <= 1 = off)
    int no_frame_refine = 0;   // SLM_TC_NO_FRAME_REFINE: batched calls keep the L2-fed refine kernel (A/B)
    int tc_plan_mt = 0;    // SLM_TC_PLAN_MT: experimental planner that also picks the query tiles per CTA (off by default)
    int tc_chain_min = 1;  // ... and at least this many when the run is long enough (SLM_TC_CHAIN_MIN; tests)
    int frame_warps = 8;               // warps per CTA of the frame kernel: 4, 8 or 16 (SLM_FRAME_WARPS)
    long long frame_max_clk = 26000;   // AUTO prefers the single-launch frame kernel up to this estimated cost, twice that with
                                       // cross-check (SLM_FRAME_MAX_CLK; 0 = never; calibration in profiles/r1_calib_frame.txt)
    int64_t launches = 0;
    // device buffers owned by the ctx, each grown on demand (never shrunk)
    slm_buf scratch;   // variant-internal partial results
    slm_buf keys;      // forward packed keys
    slm_buf rev;       // reverse-search packed keys (cross-check)
    slm_buf misc;      // pair lists, expanded operands, ...
    slm_buf io;        // device copies of host inputs / outputs (slm_knn2_host)
    slm_buf tickets;   // per-group atomic tickets of the frame kernel (zero between launches)
    slm_buf chi2_leaves;   // leaves of numpy's pairwise-sum tree for chi2_leaves_k words (wide chi-square scan, bow.cu)
    int chi2_leaves_k = 0, chi2_n_leaves = 0, chi2_n_inner = 0;
    int chi2_levels[26] = {};   // Chi2Levels of bow.cu (number of levels, first inner node of every level)
    // pinned staging for host results
    void *pin = nullptr;
    size_t pin_bytes = 0;
    // pinned ring the host path stages PAGEABLE train sets through (kStageSlots chunks), filled by host_threads threads
    void *stage_pin = nullptr;
    size_t stage_bytes = 0;
    cudaEvent_t stage_ev[4] = {};                 // H2D of the chunk in slot i has finished: the slot may be refilled
    int host_threads = 4;                         // SLM_HOST_STAGE_THREADS (0 = let the driver stage pageable memory)
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;           // H2D chunks of slm_knn2_host
    cudaEvent_t chunk_ev[kMaxHostChunks] = {};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // optional timing of the dominant kernel (slm_profile_enable)
    unsigned *done_counter = nullptr;   // 4-byte device counter (last-block-done of the exchange producers)
    // sharded exchange (exchange.cu)
    int *exchange_status = nullptr;     // mapped pinned int[4]: (code, rank, step, value seen) written by the merge kernel
    long long exchange_max_blocks = 296;   // grid cap of the wait + merge kernel (SLM_EXCHANGE_MAX_BLOCKS; loopback tests lower it)
    unsigned exchange_max_polls = 1u << 23;   // flag polls before a peer is reported lost (~4 s; SLM_EXCHANGE_MAX_POLLS)
    long long exchange_two_phase_min = 32768; // sharded tensor path: from this many queries on, the ranks first agree on every
                                              // query's GLOBAL best two candidate chunks and only their owners refine them
                                              // (SLM_EXCHANGE_TWO_PHASE_MIN; 0 = never)
    int exchange_two_phase_world = 4;         // ... and only from this many ranks on: with 2 ranks the second exchange round
                                              // costs more than the halved refine saves (SLM_EXCHANGE_TWO_PHASE_WORLD)
    int exchange_wide_keys = 0;         // SLM_EXCHANGE_WIDE_KEYS: always exchange 64-bit keys (A/B of the compact form)
    // stream bookkeeping: every device entry point runs between slm_enter() and slm_leave()
    cudaStream_t cur_stream = nullptr;  // stream of the call in progress (workspace growth is ordered on it)
    cudaStream_t last_stream = nullptr; // stream of the previous call
    cudaEvent_t last_ev = nullptr;      // recorded at the end of every call on its stream
    bool have_last = false;
    int tc_fp4 = 1;                     // AUTO uses the mxf4 tensor kernel where it applies (SLM_TC_FP4=0: fp8 only)
    int tc4_chunk = 0;                  // SLM_TC4_CHUNK: force the fp4 kernel's candidate chunk width (120 or 40; tests / A-B)
    int tc4_timing = 0;                 // SLM_TC4_TIMING: run the diagnostic build of the mxf4 kernel and print per-phase cycles
    int no_pdl = 0;                     // SLM_NO_PDL: launch the refine / merge kernels without programmatic dependent launch
    int profile = 0;
    static constexpr int kMaxProf = 4096;
    cudaEvent_t *prof_ev = nullptr;   // kMaxProf events, created lazily
    unsigned char *prof_tag = nullptr;
    int prof_n = 0;
    long long prof_dropped = 0;       // marks lost because the buffer was full (reported by slm_profile_read)
    int trace = 0;                    // SLM_TRACE=1: slm_profile_read prints every interval to stderr
};

// Event marks on the launching stream when profiling is on (no-ops otherwise).  The dominant kernel is
// bracketed by MAIN_BEGIN / MAIN_END; the API entry points add CALL_BEGIN / CALL_END for tracing.
enum slm_prof_tag { SLM_TAG_CALL_BEGIN = 0, SLM_TAG_MAIN_BEGIN = 1, SLM_TAG_MAIN_END = 2, SLM_TAG_CALL_END = 3 };
int slm_prof_mark(slm_ctx *ctx, cudaStream_t stream, int tag);
inline int slm_prof_begin(slm_ctx *ctx, cudaStream_t stream) { return slm_prof_mark(ctx, stream, SLM_TAG_MAIN_BEGIN); }
inline int slm_prof_end(slm_ctx *ctx, cudaStream_t stream) { return slm_prof_mark(ctx, stream, SLM_TAG_MAIN_END); }

// error plumbing (api.cu)
int slm_fail(int code, const char *fmt, ...);
#define SLM_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return slm_fail(SLM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                               \
    } while (0)
#define SLM_TRY(expr)                  \
    do {                               \
        int r__ = (expr);              \
        if (r__ != SLM_OK) return r__; \
    } while (0)

// Grow `buf` to at least `bytes`.  Inside a call (slm_enter .. slm_leave) the old block is released and the new one
// allocated in stream order on the call's stream (cudaFreeAsync / cudaMallocAsync): kernels of earlier calls that
// still use the old block finish first, and nothing synchronises the device.
int slm_buf_reserve(slm_ctx *ctx, slm_buf *buf, size_t bytes);

// Stream bookkeeping of the device entry points.  The workspace of a ctx (keys / scratch / tickets ...) is shared by
// all calls, so a call on another stream than the previous one first waits (on the device) for that call's work.
int slm_enter(slm_ctx *ctx, cudaStream_t stream);
int slm_leave(slm_ctx *ctx, cudaStream_t stream);

// Kernel launch through cudaLaunchKernelEx; pdl = programmatic dependent launch (the kernel may start while the
// previous kernel of the stream drains; it must order itself with griddepcontrol.wait or its own flags).
template <typename... KArgs, typename... Args>
inline cudaError_t slm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                              Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- variant P: LOP3(XOR)+POPC on the integer pipe (knn2_popc.cu) --------------------------------
// Single problem or a batch of equally-shaped problems (n_prob >= 1, pairs given by frame indices).
// Writes packed keys uint64[n_prob][nq][2].
int slm_popc_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt,
                       int64_t base, uint64_t *keys_out, cudaStream_t stream);
int slm_popc_knn2_keys_batched(slm_ctx *ctx, const uint32_t *desc, int64_t n_per_frame,
                               const int32_t *pairs_dev, int64_t n_pairs, uint64_t *keys_out,
                               cudaStream_t stream);

// ---- streaming kernel for nq <= 8 (HBM-bound shapes; knn2_stream.cu) -------------------------------------
int slm_stream_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt,
                         int64_t base, uint64_t *keys_out, cudaStream_t stream);

// ---- masked search: knnMatch(q, t, k=2, mask=M), M uint8[nq][nt] with row stride mask_stride (knn2_masked.cu) ------
int slm_masked_knn2(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                    const uint8_t *mask, int64_t mask_stride, int32_t ratio_num, int32_t ratio_den, int32_t *idx_out,
                    int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream);

// ---- frame-to-frame shapes in one launch: search + merge + ratio (+ cross-check) (knn2_frame.cu) ------------
// keys_out (uint64[nq][2]) and idx/dist/accept are each optional; cross-check needs accept_out.
bool slm_frame_eligible(slm_ctx *ctx, int64_t nq, int64_t nt, bool cross);
int slm_frame_knn2(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                   int32_t ratio_num, int32_t ratio_den, int32_t cross_check, uint64_t *keys_out, int32_t *idx_out,
                   int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream);

int slm_frame_revcheck(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t base, int32_t ratio_num,
                       int32_t ratio_den, const uint64_t *fwd_keys, int32_t *idx_out, int32_t *dist_out,
                       uint8_t *accept_out, cudaStream_t stream);

// ---- finalize / merge / compaction (finalize.cu) -------------------------------------------------
// keys uint64[n][2] -> idx/dist/accept.  rev_keys (optional) = reverse search keys uint64[nt][2] used
// for the cross-check; train_index_base is subtracted from the forward index to address it.
// rev_by_query != 0: rev_keys is uint64[n][2], entry i = reverse search of query i's best train row over all queries.
int slm_finalize(slm_ctx *ctx, const uint64_t *keys, int64_t n, int32_t ratio_num, int32_t ratio_den,
                 const uint64_t *rev_keys, int64_t nt, int64_t train_index_base, int32_t *idx_out,
                 int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream, int rev_by_query = 0);
// out[i] = train row best(i) of the forward keys (zeros when query i has no neighbour): uint32[n][8]
int slm_gather_best_rows(slm_ctx *ctx, const uint64_t *keys, int64_t n, int64_t train_index_base, const uint32_t *t,
                         uint32_t *out, cudaStream_t stream);
int slm_merge_keys(slm_ctx *ctx, const uint64_t *gathered, int32_t n_shards, int64_t nq,
                   uint64_t *keys_out, cudaStream_t stream);
int slm_merge_finalize(slm_ctx *ctx, const uint64_t *gathered, int32_t n_shards, int64_t nq, int32_t ratio_num,
                       int32_t ratio_den, int32_t *idx_out, int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream);
int slm_gather(slm_ctx *ctx, const void *src, int32_t row_bytes, const int32_t *matches, const int32_t *count,
               int64_t capacity, int32_t column, void *out, cudaStream_t stream);
int slm_filter_points3d_impl(slm_ctx *ctx, const double *pts3d, int64_t n, double max_distance, uint8_t *accept,
                             cudaStream_t stream);
int slm_compact(slm_ctx *ctx, const int32_t *idx, const int32_t *dist, const uint8_t *accept, int64_t nq,
                int32_t stop_at_short_row, int32_t *matches_out, int32_t *count_out, cudaStream_t stream);

// ---- NVLink exchange of per-rank keys (sharded path; protocol in exchange.cuh) --------------------------------
// Every rank's key buffer [2 phases][2 halves][world][cap][2] keys and flag array uint32[2][2][world] are peer-mapped on
// this GPU.  Phase 0 carries what the first kernel of a step produces (the exact keys, or -- two-phase form for many
// queries -- the candidate-chunk keys), phase 1 the exact keys of the two-phase form.
static constexpr int kSlmMaxWorld = 16;
struct slm_exchange {
    unsigned char *peer_keys[kSlmMaxWorld];
    unsigned *peer_flags[kSlmMaxWorld];
    int rank, world;
    unsigned step;
    long long cap;
    int key_bytes;            // 8: uint64 keys (distance << 32 | index); 4: compact uint32 keys (distance << 16 | index)
    unsigned *done_counter;   // device counter for the last-block-done pattern (zero between launches)
    unsigned max_polls;       // flag polls before a peer is reported lost
    int *status;              // mapped host int[4]: (code, rank, step, value seen), written on a timeout
};
int slm_exchange_setup(slm_ctx *ctx, slm_exchange *ex, const uint64_t *peer_keys_host, const uint64_t *peer_flags_host,
                       int32_t rank, int32_t world, uint32_t step, int64_t cap, int64_t nt_global);
// producer for keys that already sit in local memory (non-tensor variants)
int slm_exchange_store(slm_ctx *ctx, const slm_exchange &ex, const uint64_t *local_keys, int64_t nq, cudaStream_t stream);
// wait for every rank's keys of this step, merge, finalise
int slm_exchange_wait_merge(slm_ctx *ctx, const slm_exchange &ex, int phase, int64_t nq, int32_t ratio_num, int32_t ratio_den,
                            int32_t *idx_out, int32_t *dist_out, uint8_t *accept_out, cudaStream_t stream);
// SLM_ERR_TIMEOUT if an earlier merge kernel reported a lost peer (clears the report)
int slm_exchange_check(slm_ctx *ctx);

// ---- variant T: +-1 fp8 expansion + tcgen05.mma with TMEM accumulators (knn2_tc.cu) ---------------
// fp4 = true: kind::mxf4 kernel (knn2_tc4.cu) where CTA pairs apply, else the fp8 kernels
int slm_tc_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt,
                     int64_t base, uint64_t *keys_out, cudaStream_t stream, bool fp4);
// Search + NVLink exchange: the refine kernel stores each query's keys straight into every peer's buffer and its
// last block publishes the flags (the caller follows with slm_exchange_wait_merge).
// *phase_out = the exchange phase whose keys the caller's wait + merge must read (1 after the two-phase form)
int slm_tc_knn2_exchange(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt, int64_t base,
                         const slm_exchange &ex, cudaStream_t stream, bool fp4, int *phase_out);
// Chained batch (config 3): the caller's pairs sorted by query frame and cut into units of pairs that share it
// (all device arrays; see TcParams in knn2_tc.cu).  Optional: nullptr = every pair is its own launch item.
struct slm_chain {
    const int32_t *pairs_sorted;   // int32[n_pairs][2]
    const int32_t *prob;           // int32[n_pairs]: caller's pair index of every sorted pair
    const int32_t *units;          // int32[n_units][2] = (first sorted pair, number of pairs)
    int n_units;
};
int slm_tc_knn2_keys_batched(slm_ctx *ctx, const uint32_t *desc, int64_t n_per_frame, const int32_t *pairs_dev,
                             int64_t n_pairs, uint64_t *keys_out, cudaStream_t stream, const slm_chain *chain = nullptr,
                             bool fp4 = false);

// ---- variant B: b1 AND.POPC mma.sync tiles (knn2_bmma.cu; emulated by ptxas on sm_100a) -----------
int slm_bmma_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt,
                       int64_t base, uint64_t *keys_out, cudaStream_t stream);

// ---- shape policy (dispatch.cu) -------------------------------------------------------------------
int slm_auto_knn2_keys(slm_ctx *ctx, const uint32_t *q, int64_t nq, const uint32_t *t, int64_t nt,
                       int64_t base, uint64_t *keys_out, cudaStream_t stream);
int slm_batched_knn2_keys(slm_ctx *ctx, const uint32_t *desc, int64_t n_per_frame, const int32_t *pairs_dev,
                          int64_t n_pairs, uint64_t *keys_out, cudaStream_t stream, const slm_chain *chain = nullptr);

// chi-square scan: up to this many words a stored histogram is ONE leaf of numpy's pairwise sum (an 8-lane group per stored
// histogram); above, one block per stored histogram (chi2_scan_wide_kernel).  Largest supported vocabulary: kChi2MaxWords.
static constexpr int kChi2LeafWords = 128;
static constexpr int kChi2MaxWords = 1 << 19;      // leaf + inner-node sums of one stored histogram must fit shared memory
// ---- bag-of-words follow-on (bow.cu) ------------------------------------------------------------------
int slm_bow_hist_impl(slm_ctx *ctx, const int32_t *idx, int64_t n, int32_t idx_stride, int32_t n_words, int32_t *hist,
                      cudaStream_t stream);
// binary vocabulary training, update step (vocab.cu)
int slm_vocab_update_impl(slm_ctx *ctx, const uint32_t *desc, int64_t n, const int32_t *words, int32_t stride,
                          uint32_t *vocab, int32_t n_words, int32_t *counts_out, int32_t *changed_out, cudaStream_t stream);
int slm_chi2_scan_impl(slm_ctx *ctx, const int32_t *hq, const int32_t *db, int64_t n_db, int32_t k, double *dist,
                       int32_t *best_idx, double *best_val, cudaStream_t stream);
