// exchange.cuh -- device side of the NVLink key exchange of the sharded path.
//
// Every rank's buffer  keys[2][world][cap][2]  (uint64 keys, or uint32 compact keys when every global train index
// fits 16 bits -- BASELINE config 4's 65 536-word vocabulary) and its flag words  flags[2][world]  are peer-mapped on
// every GPU of the box.  A producer kernel (the tensor path's refine kernel, or exchange_store_kernel for the other
// variants) stores each query's top-2 keys straight into slot [step & 1][rank] of EVERY rank's buffer (peer stores over
// NVLink / NVSwitch), then slm_exchange_publish() -- system-scope fence, last-block-done, st.release.sys of `step`
// into flag [step & 1][rank] of every peer.  exchange_wait_merge_kernel (exchange.cu) acquires the flags of all
// ranks and merges.  Two buffer halves indexed by the step's parity make the scheme race-free without any other
// synchronisation: a rank can only publish step s + 1 after its own merge of step s, so nobody overwrites the half
// a peer may still be reading.
#pragma once
#include "slm_internal.cuh"

// compact key: (distance << 16) | global index, 0xFFFFFFFF = none; unsigned order is preserved
__device__ __forceinline__ unsigned slm_key_compact(unsigned long long k)
{
    return k == kKeyNone ? 0xFFFFFFFFu : ((unsigned)(k >> 32) << 16) | (unsigned)(k & 0xFFFFull);
}
__device__ __forceinline__ unsigned long long slm_key_widen(unsigned k)
{
    return k == 0xFFFFFFFFu ? kKeyNone : ((unsigned long long)(k >> 16) << 32) | (unsigned long long)(k & 0xFFFFu);
}

// Store query q's keys into slot [step & 1][rank][q] of ONE rank's buffer (callers spread (peer, query) pairs over threads
// so that consecutive threads write consecutive queries of the same peer).
// first key slot of (phase, this step's half, source rank) in any rank's buffer
__device__ __forceinline__ long long slm_exchange_slot(const slm_exchange &ex, int phase, int src_rank)
{
    return ((long long)(phase * 2 + (int)(ex.step & 1u)) * ex.world + src_rank) * ex.cap;
}
__device__ __forceinline__ void slm_exchange_store_to(const slm_exchange &ex, int peer, long long q, unsigned long long k1,
                                                      unsigned long long k2, int phase = 0)
{
    const long long slot = slm_exchange_slot(ex, phase, ex.rank) + q;
    if (ex.key_bytes == 4) reinterpret_cast<uint2 *>(ex.peer_keys[peer])[slot] = make_uint2(slm_key_compact(k1), slm_key_compact(k2));
    else reinterpret_cast<ulonglong2 *>(ex.peer_keys[peer])[slot] = make_ulonglong2(k1, k2);
}

// Candidate-chunk keys of the two-phase form: ((max dot + 257) << 32) | ~first global row of the chunk, 0 = none; a 64-bit
// max prefers the larger dot, then the chunk with the lower rows.  Compact form (whole train set <= 65 536 rows):
// (dot + 257) << 22 | (0x3FFFFF - first row).
__device__ __forceinline__ unsigned slm_chunk_compact(unsigned long long k)
{
    return k == 0 ? 0u : ((unsigned)(k >> 32) << 22) | (0x3FFFFFu - (0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)));
}
__device__ __forceinline__ unsigned long long slm_chunk_widen(unsigned k)
{
    return k == 0 ? 0ull : ((unsigned long long)(k >> 22) << 32) | (unsigned long long)(0xFFFFFFFFu - (0x3FFFFFu - (k & 0x3FFFFFu)));
}
__device__ __forceinline__ void slm_exchange_store_chunks_to(const slm_exchange &ex, int peer, long long q, unsigned long long c1,
                                                             unsigned long long c2)
{
    const long long slot = slm_exchange_slot(ex, 0, ex.rank) + q;
    if (ex.key_bytes == 4) reinterpret_cast<uint2 *>(ex.peer_keys[peer])[slot] = make_uint2(slm_chunk_compact(c1), slm_chunk_compact(c2));
    else reinterpret_cast<ulonglong2 *>(ex.peer_keys[peer])[slot] = make_ulonglong2(c1, c2);
}
// Poll this rank's own flags of `phase` until every rank shows this step (threads 0..world-1 of the block poll; bounded).
// Returns false -- after reporting (code, rank, step, value) once per grid -- if a peer never delivered.
__device__ __forceinline__ bool slm_exchange_wait_flags(const slm_exchange &ex, int phase)
{
    __shared__ int s_failed;
    if (threadIdx.x == 0) s_failed = 0;
    __syncthreads();
    if ((int)threadIdx.x < ex.world) {
        const unsigned *mine = ex.peer_flags[ex.rank] + (phase * 2 + (int)(ex.step & 1u)) * ex.world + threadIdx.x;
        unsigned v, polls = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int)(v - ex.step) >= 0) break;
            if (++polls > ex.max_polls) {
                if (atomicExch(&s_failed, 1) == 0 && blockIdx.x == 0) {
                    volatile int *status = ex.status;
                    status[1] = (int)threadIdx.x;
                    status[2] = (int)ex.step;
                    status[3] = (int)v;
                    __threadfence_system();
                    status[0] = 1;
                }
                break;
            }
            __nanosleep(polls < 64 ? 20 : 500);
        }
    }
    __syncthreads();
    return s_failed == 0;
}

// Called by ALL threads of EVERY block of the producer kernel after their stores: the block that finishes last
// publishes `step` into this rank's flag on every peer.
// `wrote` is kept for the callers' bookkeeping only.
__device__ __forceinline__ void slm_exchange_publish(const slm_exchange &ex, bool wrote = true, int phase = 0)
{
    __shared__ bool s_last;
    (void)wrote;
    // ONE system-scope fence per block, by the thread that counts the block in: the barrier orders every thread's peer
    // stores before it and the fence is cumulative (the pattern of a grid-wide barrier).  A fence per storing thread was a
    // third of the refine kernel on config 4 (a million MEMBAR.SYS).
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(ex.done_counter, 1u) == gridDim.x - 1;
        if (s_last) *ex.done_counter = 0;
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < ex.world) {
        __threadfence_system();
        unsigned *flag = ex.peer_flags[threadIdx.x] + (phase * 2 + (int)(ex.step & 1u)) * ex.world + ex.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(ex.step) : "memory");
    }
}

// Programmatic dependent launch (PDL): let the next kernel of the stream start launching while this one drains /
// wait (in the dependent kernel) until the previous kernel's memory is visible.
__device__ __forceinline__ void slm_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void slm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
