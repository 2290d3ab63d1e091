// tc_common.cuh -- sm_100a device helpers for variant T (tcgen05.mma + TMEM), hand-written PTX.
//
// The distance matrix of a (128 query) x (256 train) tile is a dense binary contraction.  Every
// descriptor bit is expanded to an fp8 (e4m3) value +1.0 (bit 0) or -1.0 (bit 1); then
//      dot(x, y) = #agreeing bits - #differing bits = 256 - 2 * Hamming(a, b)
// so ONE tcgen05.mma chain (K = 256 = 8 instructions of K = 32) yields the Hamming distance as an
// exact small integer in an fp32 TMEM accumulator -- no popc(a)+popc(b) correction terms.
//
// Shared-memory operand layout: K-major, SWIZZLE_NONE ("interleaved") canonical layout, 8-bit elements:
//   core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes (row r at +16*r)
//   LBO (leading byte offset) = distance between the two 16-byte K-chunks of one MMA      = 128 B
//   SBO (stride  byte offset) = distance between consecutive 8-row groups along M / N     = 2048 B
// i.e. an 8-row group is 2 KB: 16 K-chunks x 128 B.  Byte (row r, k) lives at
//   (r / 8) * 2048 + (k / 16) * 128 + (r % 8) * 16 + (k % 16).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

namespace tc {

constexpr uint32_t kLBO = 128;
constexpr uint32_t kSBO = 2048;
constexpr uint32_t kRowGroupBytes = 2048;          // 8 rows x 256 expanded bytes
constexpr uint32_t kFp8PlusOne = 0x38383838u;      // e4m3 +1.0 in every byte; bit 7 set => -1.0

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (a hung box is a strike); trap instead.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int tag)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("slammatch: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag, blockIdx.x,
                   threadIdx.x, parity);
            __trap();
        }
    }
}
// Same, for waiters that are far ahead of their producer (expanders three stages ahead of the MMA): back off
// between polls so the spinning warp does not burn issue slots and power -- the kernel runs power-limited
// (GPC clock 1.75 GHz under load), so idle polling costs clock.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, int tag, unsigned ns)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (++spins > (1u << 24)) {
            printf("slammatch: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag, blockIdx.x,
                   threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM allocation --------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_in_smem, uint32_t ncols)   // whole warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)          // whole warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors -------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64-bit): start address >> 4 in [0,14), LBO >> 4 in [16,30),
// SBO >> 4 in [32,46), descriptor version 1 (Blackwell) in [46,48), layout type in [61,64) (0 = no swizzle).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo = kLBO, uint32_t sbo = kSBO)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor for kind::f8f6f4: D = F32 (bits [4,6) = 1), A = B = E4M3 (0), both K-major,
// N >> 3 in [17,23), M >> 4 in [24,29).
__host__ __device__ constexpr uint32_t idesc_e4m3_f32(uint32_t M, uint32_t N)
{
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one elected thread.
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives when every tcgen05 op issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers ---------------------------------------------------------------------------
// 32 lanes x 32 consecutive 32-bit columns: thread l of the warp gets lane (base_lane + l), columns
// [col, col + 32).  The wait is fused so the registers are valid when the statement retires.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// Same load without the wait (several loads in flight); pair with tmem_ld_wait() / pin_regs().
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32])
{
    // the "+r" operands pin every consumer of v behind the wait in program order
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

// Pins the consumers of v behind this point in program order without emitting an instruction (used after a
// single tcgen05.wait::ld that covers several outstanding loads).
__device__ __forceinline__ void pin_regs(uint32_t (&v)[32])
{
    asm volatile(""
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

// ---- bit -> fp8 (+-1) expansion -----------------------------------------------------------------------
// The contraction index K may be permuted freely as long as both operands use the same permutation, so
// the expansion picks the cheapest bit -> byte mapping: output word k of a 32-bit descriptor word w is
//      ((w << k) & 0x80808080) | 0x38383838          k = 0..7
// i.e. bits {7-k, 15-k, 23-k, 31-k} of w land on the sign bits of the four e4m3 bytes (+1.0 = 0x38,
// -1.0 = 0xB8).  One shift + one LOP3 per 4 output bytes.
__device__ __forceinline__ uint32_t expand_shifted(uint32_t x)
{
    return (x & 0x80808080u) | kFp8PlusOne;
}
// One 32-bit descriptor word -> 32 expanded bytes = two 16-byte K-chunks.
__device__ __forceinline__ void expand_word(uint32_t w, uint4 &lo, uint4 &hi)
{
    lo.x = expand_shifted(w);
    lo.y = expand_shifted(w << 1);
    lo.z = expand_shifted(w << 2);
    lo.w = expand_shifted(w << 3);
    hi.x = expand_shifted(w << 4);
    hi.y = expand_shifted(w << 5);
    hi.z = expand_shifted(w << 6);
    hi.w = expand_shifted(w << 7);
}
// Expand one 256-bit descriptor (8 words) into its row of a K-major no-swizzle operand tile.
// `tile` = shared-memory byte address of the tile, `row` = row inside the tile.
__device__ __forceinline__ void expand_row_to_smem(uint32_t tile, int row, const uint4 &d0, const uint4 &d1)
{
    const uint32_t base = tile + (uint32_t)(row >> 3) * kRowGroupBytes + (uint32_t)(row & 7) * 16;
    const uint32_t w[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint4 lo, hi;
        expand_word(w[i], lo, hi);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (2 * i) * kLBO), "r"(lo.x), "r"(lo.y),
                     "r"(lo.z), "r"(lo.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (2 * i + 1) * kLBO), "r"(hi.x), "r"(hi.y),
                     "r"(hi.z), "r"(hi.w) : "memory");
    }
}


// ---- 2-CTA (cta_group::2) helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.
// Default semantics (.release.cta), as CUTLASS's ClusterBarrier::arrive(cta_id): the data handed over here
// never travels through generic-proxy global memory -- operand tiles are published to the async proxy with
// fence.proxy.async, accumulators with tcgen05.fence -- so no cluster/gpu-scope membar is wanted
// (a .release.cluster arrive costs MEMBAR.ALL.GPU + ERRBAR per call: 18 % of all stall samples in the first
// ncu capture of this kernel).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t cta)
{
    asm volatile("{\n\t.reg .b32 ra;\n\t"
                 "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// Same, relaxed: for signals that order only tcgen05 / TMEM work (accumulator hand-back).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t *bar, uint32_t cta)
{
    asm volatile("{\n\t.reg .b32 ra;\n\t"
                 "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity, int tag)
{
    mbar_wait(bar, parity, tag);   // local barrier; remote arrivals need no cluster-scope acquire (see above)
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t *slot_in_smem, uint32_t ncols)   // one warp in EACH CTA
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256 split over the CTA pair (128 rows each); each CTA supplies
// its own A rows and its half of the B rows from the same shared-memory offsets.  Leader CTA, one thread.
__device__ __forceinline__ void umma_f8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Commit to the mbarrier at this offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_2cta(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}


// ---- lean MMA issue -----------------------------------------------------------------------------------
// The issuing thread is on the critical path: one 128x256x32 MMA retires every 128 cycles, so building
// 64-bit descriptors with generic ALU code per instruction (ncu: ~190 SASS instructions per 8-MMA job,
// the issuer warp busy 100 % of the time, tensor pipe 85 % active) costs throughput.  The descriptors of a
// job differ only in the low word (start address >> 4, advanced by 16 per K step); the high word is a
// constant: SBO >> 4 | version 1 << 14.
constexpr uint32_t kDescHi = (kSBO >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFF) | ((kLBO >> 4) << 16); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int CTAS>
__device__ __forceinline__ void umma_f8_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate)
{
    if (CTAS == 1)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                     "setp.ne.b32 p, %5, 0;\n\t"
                     "mov.b64 da, {%1, %4};\n\t"
                     "mov.b64 db, {%2, %4};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %3, p;\n\t}"
                     ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHi), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                     "setp.ne.b32 p, %5, 0;\n\t"
                     "mov.b64 da, {%1, %4};\n\t"
                     "mov.b64 db, {%2, %4};\n\t"
                     "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %3, p;\n\t}"
                     ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHi), "r"(accumulate) : "memory");
}
// One job = 8 MMAs over K = 256 (K step = 32 bytes = 2 K-chunks = +16 in descriptor units).
template <int CTAS>
__device__ __forceinline__ void umma_job(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_f8_lo<CTAS>(tmem_d, a_lo + k * 16, b_lo + k * 16, idesc, k > 0 ? 1u : 0u);
}

}  // namespace tc

// =====================================================================================================
// Variant T4: the same contraction on the block-scaled FP4 path (tcgen05.mma kind::mxf4, twice the fp8 rate).
// +-1 is exact in e2m1 (+1.0 = 0x2, -1.0 = 0xA), every UE8M0 block scale is 1.0 (0x7F), so
// dot(x, y) = 256 - 2 * Hamming(a, b) still holds exactly in the fp32 accumulator.
// Operand layout: K-major, no swizzle, 4-bit elements: core matrix = 8 rows x 16 bytes (32 values),
//   LBO = 128 B between the 8 K-chunks of a row group, SBO = 1024 B between 8-row groups; byte (row r, chunk c) at
//   (r / 8) * 1024 + c * 128 + (r % 8) * 16.  One MMA consumes K = 64 = 2 chunks (descriptor start + 256 B).
// TMEM: two 240-column accumulators + 32 scale-factor columns (all bytes 0x7F, so the scale-factor layout is moot)
// fill the 512 columns; a job is therefore 256 query rows (CTA pair) x 240 train rows.
// Validated stand-alone by csrc/microbench/fp4_probe2.cu (profiles/r2_fp4_probe.txt): exact, 16381 MAC/clk/SM.
// =====================================================================================================
namespace tc4 {

constexpr int kTileN = 240;                        // train rows per job (N of the MMA; 120 per CTA of the pair)
constexpr int kHalfN = kTileN / 2;
constexpr uint32_t kSfCol = 2 * kTileN;            // TMEM columns [480, 512)
constexpr uint32_t kScaleOnes = 0x7F7F7F7Fu;       // four UE8M0 1.0
constexpr uint32_t kE2m1PlusOne = 0x22222222u;     // eight e2m1 +1.0; bit 3 of a nibble set => -1.0
constexpr uint32_t kLBO = 128, kSBO = 1024;
constexpr uint32_t kRowBytes = 128;                // 256 e2m1 values
constexpr uint32_t kATileBytes = 128 * kRowBytes;  // 16 KB: 128 query rows
constexpr uint32_t kBHalfBytes = kHalfN * kRowBytes;   // 15 KB: this CTA's 120 rows of a train tile
constexpr uint32_t kDescHi = (kSBO >> 4) | (1u << 14);

// Instruction descriptor of the block-scaled kinds (cute/arch/mma_sm100_desc.hpp, InstrDescriptorBlockScaled):
// a_format [7,10) = b_format [10,13) = 1 (MXF4 E2M1), both K-major, N >> 3 in [17,23), scale format [23] = 1 (UE8M0),
// M >> 4 in [24,29), scale-factor ids 0, K = 64.
__host__ __device__ constexpr uint32_t idesc_mxf4(uint32_t M, uint32_t N)
{
    return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFF) | ((kLBO >> 4) << 16); }

// cta_group::2, leader CTA, one thread; `sf` = TMEM address of the scale-factor columns (same for A and B)
__device__ __forceinline__ void umma_mxf4_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate,
                                             uint32_t sf)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 da, {%1, %4};\n\t"
                 "mov.b64 db, {%2, %4};\n\t"
                 "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], da, db, %3, [%6], [%6], p;\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHi), "r"(accumulate), "r"(sf) : "memory");
}
// One job = 4 MMAs over K = 256.
__device__ __forceinline__ void umma_job(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t sf)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_mxf4_lo(tmem_d, a_lo + k * 16, b_lo + k * 16, idesc, k > 0 ? 1u : 0u, sf);
}

// 32 lanes x 32 columns <- one 32-bit value (whole warp, its own lane quarter)
__device__ __forceinline__ void tmem_fill32(uint32_t taddr, uint32_t v)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n\t"
        "tcgen05.wait::st.sync.aligned;"
        ::"r"(taddr), "r"(v) : "memory");
}

// One 32-bit descriptor word -> one 16-byte K-chunk: output word j takes bits {3-j, 7-j, ..., 31-j} of w onto the sign
// bits of its eight nibbles (the K permutation is shared by both operands): 1 shift + 1 LOP3 per 8 values.
__device__ __forceinline__ void expand_word_to_smem(uint32_t addr, uint32_t w)
{
    // (x & 0x88888888) | 0x22222222 as ONE LOP3 (LUT 0xEA = (a & b) | c) with both constants in registers; written as C the
    // two different immediates cost two LOP3 per word
    uint32_t o0, o1, o2, o3;
    const uint32_t m = 0x88888888u, c = kE2m1PlusOne;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(o0) : "r"(w), "r"(m), "r"(c));
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(o1) : "r"(w << 1), "r"(m), "r"(c));
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(o2) : "r"(w << 2), "r"(m), "r"(c));
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(o3) : "r"(w << 3), "r"(m), "r"(c));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
}
// byte offset of (row, K-chunk) inside an operand tile
__device__ __forceinline__ uint32_t tile_offset(int row, int chunk)
{
    return (uint32_t)(row >> 3) * kSBO + (uint32_t)chunk * kLBO + (uint32_t)(row & 7) * 16;
}
__device__ __forceinline__ void expand_row_to_smem(uint32_t tile, int row, const uint4 &d0, const uint4 &d1)
{
    const uint32_t base = tile + tile_offset(row, 0);
    const uint32_t w[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) expand_word_to_smem(base + i * kLBO, w[i]);
}

}  // namespace tc4
