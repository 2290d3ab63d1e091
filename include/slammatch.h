/*
 * slammatch.h -- C-ABI of libslammatch.so: exact brute-force Hamming 2-nearest-neighbour matching of
 * 256-bit ORB descriptors on NVIDIA B200 (sm_100a), with the Lowe ratio test and cross-check.
 *
 * This is the drop-in boundary for ONE path of the reference pipeline (DavidHan008/SLAM-1, a Python
 * program; file:line below are relative to its tree).  The reference has no FFI of its own: it calls
 * OpenCV's DescriptorMatcher through the cv2 Python binding.  Each entry point cites the reference
 * interface it replaces; INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 *   reference today                                        replaced by
 *   -----------------------------------------------------  ---------------------------------------------
 *   cv2.FlannBasedMatcher(indexParams, searchParams)       slm_create / slm_destroy  (one ctx per thread+GPU,
 *     tracking.py:14-17  keypoint.py:40-43  Point3D.py:35-39   cached by the Python shim: the reference
 *                                                              re-constructs the matcher on every call)
 *   matcher.knnMatch(des1, des2, k=2)                      slm_knn2_host (numpy/host memory in, host memory out)
 *     tracking.py:22  keypoint.py:44  Point3D.py:40        slm_knn2 / slm_knn2_filter (device-resident fast path)
 *   for m, n in matches: if m.distance < 0.7*n.distance    ratio_num/ratio_den arguments of slm_knn2_filter /
 *     tracking.py:24-30  keypoint.py:45-51  Point3D.py:41-49   slm_knn2_host (integer form: den*d1 < num*d2)
 *   [kp[m.queryIdx].pt for m in good] gathers              slm_compact_matches (device-side compaction of accepted rows)
 *     tracking.py:32-33  keypoint.py:53-57  Point3D.py:50-52
 *   (config 3: all keyframe pairs)                         slm_knn2_batched
 *   (configs 4/5: train set sharded over GPUs)             slm_knn2_keys on each rank + slm_merge_top2 after the
 *                                                          all-gather (tie order = (distance, global train index),
 *                                                          OpenCV's (distance, imgIdx, trainIdx) collection order)
 *
 * Data layout
 *   A descriptor is 32 bytes.  Device entry points take `const uint32_t*` pointing at uint32[n][8]
 *   (a plain reinterpretation of the reference's uint8[n][32] rows, orb.py:23-24; little-endian,
 *   popcount is byte-order agnostic).  Rows must be 16-byte aligned (any cudaMalloc'd or torch tensor is).
 *   Results: idx int32[nq][2], dist int32[nq][2]; column 0 = nearest, column 1 = second nearest under
 *   the order (distance, train index) ascending -- lowest train index wins ties, exactly as
 *   cv2.BFMatcher(NORM_HAMMING).knnMatch.  A missing neighbour (nt < 2) is idx = -1, dist = -1, which
 *   the Python shim turns into OpenCV's short rows.  accept uint8[nq] is the ratio / cross-check verdict.
 *   Packed keys (sharded path): uint64 = (distance << 32) | global_train_index, SLM_KEY_NONE if missing.
 *
 * Conventions
 *   Every function returns 0 (SLM_OK) or a negative slm_status; slm_last_error() returns a
 *   thread-local message for the last failure.  `stream` is a cudaStream_t passed as void*
 *   (NULL = legacy default stream).  Device entry points are asynchronous and stream-ordered;
 *   *_host entry points synchronise before returning.  A ctx is not re-entrant: one ctx per
 *   (host thread, device).  The calls of one ctx share its workspace; a call on a different stream
 *   than the previous one is ordered (on the device) behind everything queued on that stream, so
 *   results stay correct when the caller switches streams -- calls of one ctx never overlap.  The library owns only its workspace (grown on demand); callers own all
 *   input and output buffers; inputs are never written.  There is NO CPU fallback: without a
 *   CUDA device every compute entry point fails with SLM_ERR_CUDA.
 */
#ifndef SLAMMATCH_H
#define SLAMMATCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLM_VERSION 200           /* major*10000 + minor*100 + patch */
#define SLM_DESC_WORDS 8          /* uint32 words per descriptor */
#define SLM_DESC_BYTES 32
#define SLM_KEY_NONE 0xFFFFFFFFFFFFFFFFull

typedef struct slm_ctx slm_ctx;

typedef enum slm_status {
    SLM_OK = 0,
    SLM_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, ratio_den <= 0, ...) */
    SLM_ERR_CUDA = -2,        /* CUDA runtime error (no device, launch failure, ...) */
    SLM_ERR_NOMEM = -3,       /* workspace allocation failed */
    SLM_ERR_UNSUPPORTED = -4, /* size beyond the supported range (global index must fit int32) */
    SLM_ERR_TIMEOUT = -5      /* sharded exchange: a peer rank never delivered its keys (reported, never a hang or a trap) */
} slm_status;

/* Distance-kernel variants (BASELINE.json north_star part (2)). */
typedef enum slm_variant {
    SLM_VARIANT_AUTO = 0,    /* pick per shape */
    SLM_VARIANT_POPC = 1,    /* LOP3(XOR)+POPC on the integer pipe */
    SLM_VARIANT_TENSOR = 2,  /* +-1 fp8 expansion + tcgen05.mma (TMEM accumulators) */
    SLM_VARIANT_BMMA = 3,    /* b1 AND.POPC mma.sync tiles (emulated by ptxas on sm_100a; kept for the A/B) */
    SLM_VARIANT_TENSOR4 = 4  /* +-1 e2m1 expansion + tcgen05.mma kind::mxf4 (block scales = 1.0): twice the fp8 rate */
} slm_variant;

/* Thread-local text of the last error returned on this thread ("" if none). */
const char *slm_last_error(void);
int slm_version(void);

/* Replaces the matcher constructor (tracking.py:17).  `device` is a CUDA ordinal. */
int slm_create(int device, slm_ctx **ctx_out);
int slm_destroy(slm_ctx *ctx);
/* Select the distance kernel for subsequent calls on this ctx (default SLM_VARIANT_AUTO). */
int slm_set_variant(slm_ctx *ctx, int variant);
/* The variant the last search on this ctx actually ran (resolves SLM_VARIANT_AUTO); 0 before any search. */
int slm_last_variant(const slm_ctx *ctx);
/* Name of the distance kernel the last search on this ctx launched ("knn2_tc2_kernel", "knn2_frame_kernel", ...);
 * "" before any search.  The string is static. */
const char *slm_last_kernel(const slm_ctx *ctx);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches evidence). */
int64_t slm_launch_count(const slm_ctx *ctx);
/*
 * Optional timing of the dominant (distance) kernel: when enabled, every launch of it is bracketed by
 * CUDA events on the launching stream.  slm_profile_read synchronises those events, returns the summed
 * duration in milliseconds and the number of launches since the last read, and resets both.
 */
int slm_profile_enable(slm_ctx *ctx, int enable);
int slm_profile_read(slm_ctx *ctx, double *kernel_ms_out, int64_t *launches_out);

/*
 * knnMatch(q, t, k=2), device-resident (tracking.py:22).  Global train index = train_index_base + row.
 * nq, nt may be 0.  train_index_base + nt must be <= 2^31 - 1.
 */
int slm_knn2(slm_ctx *ctx, const uint32_t *q_dev, int64_t nq, const uint32_t *t_dev, int64_t nt,
             int64_t train_index_base, int32_t *idx_out_dev, int32_t *dist_out_dev, void *stream);

/* Same search, result as packed keys uint64[nq][2] (input of the cross-shard all-gather). */
int slm_knn2_keys(slm_ctx *ctx, const uint32_t *q_dev, int64_t nq, const uint32_t *t_dev, int64_t nt,
                  int64_t train_index_base, uint64_t *keys_out_dev, void *stream);

/*
 * knnMatch + the reference's ratio loop (tracking.py:24-30) + optional cross-check, one call.
 *   accept[i] = has 2 neighbours && ratio_den*d1 < ratio_num*d2      (ratio_num <= 0 disables the ratio test:
 *                                                                     then accept = has >= 1 neighbour)
 *             && (cross_check == 0 || argmin_i' D[i', idx[i][0]] == i) (lowest query index on ties)
 * (7,10) is the reference's 0.7; (3,4) is BASELINE config 1's 0.75.  idx/dist/accept may each be NULL.
 */
int slm_knn2_filter(slm_ctx *ctx, const uint32_t *q_dev, int64_t nq, const uint32_t *t_dev, int64_t nt,
                    int64_t train_index_base, int32_t ratio_num, int32_t ratio_den, int32_t cross_check,
                    int32_t *idx_out_dev, int32_t *dist_out_dev, uint8_t *accept_out_dev, void *stream);

/*
 * knnMatch(q, t, k=2, mask=M) -- the optional `mask` argument of OpenCV's DescriptorMatcher::knnMatch, which the matcher
 * object built at tracking.py:17 / keypoint.py:43 / Point3D.py:39 accepts (the reference itself never passes one).
 * mask_dev is uint8[nq][mask_row_stride] on the device, mask_row_stride >= nt; pair (i, j) takes part iff
 * mask_dev[i * mask_row_stride + j] != 0.  A query with fewer than two allowed rows gets idx/dist = -1 in the missing
 * columns (the binding turns them into a short DMatch row, as OpenCV does).  Ratio test as in slm_knn2_filter; there is
 * no cross-check (OpenCV asserts mask.empty() when crossCheck is set, batch_distance.cpp:303).
 */
int slm_knn2_masked(slm_ctx *ctx, const uint32_t *q_dev, int64_t nq, const uint32_t *t_dev, int64_t nt,
                    int64_t train_index_base, const uint8_t *mask_dev, int64_t mask_row_stride, int32_t ratio_num,
                    int32_t ratio_den, int32_t *idx_out_dev, int32_t *dist_out_dev, uint8_t *accept_out_dev, void *stream);

/*
 * Config 3 (local-mapping batch): desc_dev is uint32[n_frames][n_per_frame][8]; pairs_host is a HOST
 * array int32[n_pairs][2] of (query frame, train frame).  Outputs are [n_pairs][n_per_frame][2] /
 * [n_pairs][n_per_frame]; train indices are local to the train frame.
 */
int slm_knn2_batched(slm_ctx *ctx, const uint32_t *desc_dev, int64_t n_frames, int64_t n_per_frame,
                     const int32_t *pairs_host, int64_t n_pairs, int32_t ratio_num, int32_t ratio_den,
                     int32_t *idx_out_dev, int32_t *dist_out_dev, uint8_t *accept_out_dev, void *stream);

/*
 * Cross-shard merge after the all-gather: gathered_keys_dev is uint64[n_shards][nq][2]; the result is the
 * top-2 in unsigned key order == (distance, global train index).  ratio as in slm_knn2_filter
 * (accept_out_dev may be NULL).
 */
int slm_merge_top2(slm_ctx *ctx, const uint64_t *gathered_keys_dev, int32_t n_shards, int64_t nq,
                   int32_t ratio_num, int32_t ratio_den, int32_t *idx_out_dev, int32_t *dist_out_dev,
                   uint8_t *accept_out_dev, void *stream);

/*
 * NVLink exchange + merge (the sharded path's replacement for all-gather + slm_merge_top2 when the gather buffers are
 * peer-mapped, e.g. torch symmetric memory over NVLink / NVSwitch).
 *   peer_keys_host[r]  : device address, valid on THIS GPU, of rank r's key buffer: 4 * world * nq_capacity * 16 bytes
 *                        (layout [2 phases][2 halves][world][nq_capacity][2] keys; 64-bit keys, or 32-bit compact keys
 *                        (distance << 16 | index) in the front half of every slot when nt_global <= 65536 -- config 4's
 *                        vocabulary)
 *   peer_flags_host[r] : device address of rank r's flag array uint32[2 phases][2 halves][world] (zero-initialised once)
 *   nt_global          : number of train rows over ALL ranks (selects the key width; must be the same on every rank;
 *                        0 = unknown, 64-bit keys)
 * A producer kernel stores this rank's nq x 2 keys into slot [step & 1][rank] of every peer's buffer (16- / 8-byte
 * stores over NVLink) and publishes `step` into every peer's flag [step & 1][rank] with a system-scope release; a
 * second kernel (programmatic dependent launch: resident while the producer drains) polls this rank's own flags
 * until every rank shows `step`, then merges and finalises like slm_merge_top2.  `step` must be the same on all
 * ranks, start at 1 and increase by 1 per call; two buffer halves make the scheme safe without any other
 * synchronisation.  nq <= nq_capacity (any size).  A peer that never delivers is reported, not waited for forever:
 * the merge kernel gives up after a bounded number of polls, leaves the outputs untouched, and the NEXT exchange call
 * (or slm_exchange_status) on this ctx returns SLM_ERR_TIMEOUT naming the rank and step.
 */
int slm_exchange_merge(slm_ctx *ctx, const uint64_t *local_keys_dev, int64_t nq, int64_t nq_capacity, int64_t nt_global,
                       const uint64_t *peer_keys_host, const uint64_t *peer_flags_host, int32_t rank, int32_t world,
                       uint32_t step, int32_t ratio_num, int32_t ratio_den, int32_t *idx_out_dev,
                       int32_t *dist_out_dev, uint8_t *accept_out_dev, void *stream);

/*
 * The whole sharded step in one call: search this rank's train block, exchange over NVLink, merge, finalise.
 * Arguments as slm_knn2_keys + slm_exchange_merge.  With the tensor variants the refine kernel is the producer (every
 * query's exact keys go straight into the peers' buffers); other variants run the search followed by a store kernel.
 * From 32 768 queries on (config 4) the tensor variants take a two-phase form: the ranks first exchange every query's best
 * two candidate-chunk keys (phase 0), each rank picks the GLOBAL best two chunks and re-scores only those inside its own row
 * block, and the exact keys travel in phase 1 -- the exact re-scoring is then shared by the ranks instead of repeated on each.
 */
int slm_knn2_exchange(slm_ctx *ctx, const uint32_t *q_dev, int64_t nq, const uint32_t *t_dev, int64_t nt,
                      int64_t train_index_base, int64_t nq_capacity, int64_t nt_global, const uint64_t *peer_keys_host,
                      const uint64_t *peer_flags_host, int32_t rank, int32_t world, uint32_t step, int32_t ratio_num,
                      int32_t ratio_den, int32_t *idx_out_dev, int32_t *dist_out_dev, uint8_t *accept_out_dev,
                      void *stream);
/* SLM_OK, or SLM_ERR_TIMEOUT if an earlier exchange on this ctx lost a peer (clears the report).  Call after the
 * stream has been synchronised. */
int slm_exchange_status(slm_ctx *ctx);

/*
 * Point3D.find_2D_and_3D_correspondenses' extra condition (Point3D.py:45-46): accept_dev[i] &= |X|, |Y|, |Z| <
 * max_distance of query i's triangulated point, pts3d_dev = float64[nq][3].  Run between slm_knn2_filter and
 * slm_compact_matches; the gathers of Point3D.py:50-52 are slm_gather_rows.
 */
int slm_filter_points3d(slm_ctx *ctx, const double *pts3d_dev, int64_t nq, double max_distance, uint8_t *accept_dev,
                        void *stream);

/*
 * In-process ceilings of the pipes the distance kernels run on (bench.py's roofline denominators; synchronous).
 * slm_probe_tensor_peak: back-to-back tcgen05.mma on every SM from shared memory; kind 0 = kind::f8f6f4 (variant
 *   TENSOR), 1 = kind::mxf4.block_scale (variant TENSOR4).  `loops` jobs per CTA, best of `reps` launches.
 *   tflops_out = dense TFLOP/s (2 flop per MAC), mac_per_clk_per_sm_out from clock64 inside the kernel.
 * slm_probe_popc_peak: the inner loop of variant POPC (8 XOR + 8 POPC + top-2 update per comparison) on every SM.
 */
int slm_probe_tensor_peak(slm_ctx *ctx, int32_t kind, int32_t loops, int32_t reps, double *tflops_out,
                          double *mac_per_clk_per_sm_out);
int slm_probe_popc_peak(slm_ctx *ctx, int32_t reps, double *tcmp_per_s_out, double *popc_lanes_per_clk_per_sm_out);

/*
 * Device-side compaction of accepted rows (the gathers at tracking.py:32-33 start from this list):
 * writes (queryIdx, trainIdx, distance) int32 triples in ascending queryIdx order to matches_out_dev
 * (capacity nq triples) and the count to count_out_dev (int32[1]).
 * If stop_at_short_row != 0 the list is truncated at the first row with fewer than two neighbours,
 * as the reference's swallowed ValueError does (tracking.py:25-30).
 */
int slm_compact_matches(slm_ctx *ctx, const int32_t *idx_dev, const int32_t *dist_dev,
                        const uint8_t *accept_dev, int64_t nq, int32_t stop_at_short_row,
                        int32_t *matches_out_dev, int32_t *count_out_dev, void *stream);

/*
 * Device-side gather driven by the compacted match list (the list comprehensions at tracking.py:32-33,
 * keypoint.py:53-57, Point3D.py:50-52): out[m] = src[matches[m][column]] for m < *count_dev, rows of
 * row_bytes bytes (a multiple of 4: keypoint coordinates float32[.][2] -> 8, descriptors -> 32, 3-D points
 * float64[.][3] -> 24).  column 0 gathers by queryIdx, 1 by trainIdx.  out_dev must hold capacity rows.
 */
int slm_gather_rows(slm_ctx *ctx, const void *src_dev, int32_t row_bytes, const int32_t *matches_dev,
                    const int32_t *count_dev, int64_t capacity, int32_t column, void *out_dev, void *stream);

/*
 * Bag-of-words follow-on (bag_of_words.py:23-42).
 * slm_bow_hist: histogram of word ids, hist_out_dev int32[n_words] (zeroed first).  words_dev holds one word
 *   id every `stride` int32 (stride 2 reads column 0 of the idx[n][2] result of a word-assignment search);
 *   equals np.histogram(labels, bins=n_words, range=(0, n_words-1)) for labels in [0, n_words).
 * slm_chi2_scan: dist_out_dev[i] = sum_w 2*(h[w]-db[i][w])^2 / max(1, h[w]+db[i][w]) in float64 for the n_db stored
 *   histograms db_dev int32[n_db][n_words], bit-exact with numpy (same division, same pairwise summation
 *   order), plus best_idx_dev / best_val_dev = (np.argmin, np.min) -- first minimum wins.
 *   n_words <= 2^19.  Up to 128 words (the reference's 50) eight lanes scan one stored histogram; larger vocabularies
 *   (config 4's 65536 words) take one block per stored histogram, still in numpy's summation order.
 */
int slm_bow_hist(slm_ctx *ctx, const int32_t *words_dev, int64_t n, int32_t stride, int32_t n_words,
                 int32_t *hist_out_dev, void *stream);
int slm_chi2_scan(slm_ctx *ctx, const int32_t *hist_dev, const int32_t *db_dev, int64_t n_db, int32_t n_words,
                  double *dist_out_dev, int32_t *best_idx_dev, double *best_val_dev, void *stream);

/*
 * Binary vocabulary training, one update step (bag_of_words.py:14,20 `KMeans.fit`, re-specified in Hamming space as
 * k-majority).  The assignment step is the search itself (word = idx[:,0] of
 * slm_knn2 against the current vocabulary); this call then rewrites every word of vocab_dev uint32[n_words][8]
 * IN PLACE as the bitwise majority of the descriptors assigned to it: bit <- 1 if more than half of the members
 * have it set, unchanged on an exact tie; words without members keep their centroid.
 *   desc_dev  uint32[n][8]; words_dev holds one word id every `stride` int32 (ids outside [0, n_words) are skipped)
 *   counts_out_dev   int32[n_words] members per word (optional)
 *   changed_out_dev  int32[1] number of words whose centroid changed (optional; 0 = converged)
 */
int slm_vocab_update(slm_ctx *ctx, const uint32_t *desc_dev, int64_t n, const int32_t *words_dev, int32_t stride,
                     uint32_t *vocab_dev, int32_t n_words, int32_t *counts_out_dev, int32_t *changed_out_dev,
                     void *stream);

/*
 * Host-memory convenience: what the Python shim's knnMatch(des1, des2, k=2) calls.  q_host/t_host are
 * uint8[n][32] in host memory (pinned or pageable; pageable memory is staged through the ctx's pinned
 * buffers); outputs are host arrays.  Copies, kernels and the read-back run on the ctx's own stream and
 * the call returns after they complete.  Any output pointer may be NULL.
 */
int slm_knn2_host(slm_ctx *ctx, const uint8_t *q_host, int64_t nq, const uint8_t *t_host, int64_t nt,
                  int32_t ratio_num, int32_t ratio_den, int32_t cross_check,
                  int32_t *idx_out_host, int32_t *dist_out_host, uint8_t *accept_out_host);

#ifdef __cplusplus
}
#endif
#endif /* SLAMMATCH_H */
