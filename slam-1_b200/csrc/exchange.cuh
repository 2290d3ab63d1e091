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
__device__ __forceinline__ void slm_exchange_store_to(const slm_exchange &ex, int peer, long long q, unsigned long long k1,
                                                      unsigned long long k2)
{
    const long long slot = ((long long)(ex.step & 1u) * ex.world + ex.rank) * ex.cap + q;
    if (ex.key_bytes == 4) reinterpret_cast<uint2 *>(ex.peer_keys[peer])[slot] = make_uint2(slm_key_compact(k1), slm_key_compact(k2));
    else reinterpret_cast<ulonglong2 *>(ex.peer_keys[peer])[slot] = make_ulonglong2(k1, k2);
}

// Called by ALL threads of EVERY block of the producer kernel after their stores: the block that finishes last
// publishes `step` into this rank's flag on every peer.
// `wrote` = this thread issued peer stores (only those threads need the system-scope fence).
__device__ __forceinline__ void slm_exchange_publish(const slm_exchange &ex, bool wrote = true)
{
    __shared__ bool s_last;
    if (wrote) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(ex.done_counter, 1u) == gridDim.x - 1;
        if (s_last) *ex.done_counter = 0;
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < ex.world) {
        __threadfence_system();
        unsigned *flag = ex.peer_flags[threadIdx.x] + (ex.step & 1u) * ex.world + ex.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(ex.step) : "memory");
    }
}

// Programmatic dependent launch (PDL): let the next kernel of the stream start launching while this one drains /
// wait (in the dependent kernel) until the previous kernel's memory is visible.
__device__ __forceinline__ void slm_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void slm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
