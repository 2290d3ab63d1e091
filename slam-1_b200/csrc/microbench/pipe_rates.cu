// pipe_rates.cu -- per-SM issue rates of the instructions on the matching hot path, measured with
// clock64() inside one CTA per SM (independent of DVFS).  Calibrates the roofline denominators in
// POPC32/clk/SM bounds variant P (8 POPC per comparison); the min/max and FFMA rates bound
// the tcgen05 epilogue.  Output: one line per op, "op lanes_per_clk_per_SM".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kIters = 2048;
constexpr int kChains = 8;

enum Op { POPC, LOP3, IADD3, IMNMX_U32, IMNMX3_U32, FMNMX, FMNMX3, HMNMX2, IMAD, FFMA, POPC_XOR, MIX_CMP };

template <int OP>
__global__ void __launch_bounds__(1024) rate_kernel(unsigned *out, long long *cycles, unsigned seed)
{
    unsigned v[kChains];
    float f[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) { v[c] = seed * (threadIdx.x + 1) + c * 0x9E3779B9u; f[c] = (float)(v[c] & 1023); }
    unsigned a = seed ^ 0x5bd1e995u, b = seed * 31u + 7u;
    float fa = (float)(a & 255), fb = (float)(b & 255);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if (OP == POPC) v[c] = __popc(v[c]) + a;                     // POPC + IADD (ALU is wider)
            if (OP == LOP3) v[c] = (v[c] & a) ^ b;
            if (OP == IADD3) v[c] = v[c] + a + b;
            if (OP == IMNMX_U32) v[c] = min(v[c] ^ 0u, a) + 0u, a += 0u, v[c] = max(v[c], b);
            if (OP == IMNMX3_U32) v[c] = __vimax3_u32(v[c], a, b) , a ^= 0u;
            if (OP == FMNMX) f[c] = fmaxf(fminf(f[c], fa), fb);
            if (OP == FMNMX3) f[c] = fmaxf(fmaxf(f[c], fa), fb) , fa += 0.0f;
            if (OP == HMNMX2) { asm volatile("max.f16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(a)); }
            if (OP == IMAD) v[c] = v[c] * a + b;
            if (OP == FFMA) f[c] = f[c] * fa + fb;
            if (OP == POPC_XOR) v[c] = __popc(v[c] ^ a) + b;
        }
        a += 1; b += 3; fa += 1.0f; fb += 0.5f;
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc ^= v[c] ^ __float_as_uint(f[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the real inner loop of variant P for one comparison: 8 XOR + 8 POPC + adds + packed top-2 update
__global__ void __launch_bounds__(1024) cmp_kernel(unsigned *out, long long *cycles, unsigned seed)
{
    unsigned q[8], b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
#pragma unroll
    for (int w = 0; w < 8; ++w) q[w] = seed * (threadIdx.x + 1) + w * 0x9E3779B9u;
    unsigned t0w = seed;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < kIters; ++it) {
        unsigned d = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) d += __popc(q[w] ^ (t0w + w * 0x85EBCA6Bu));
        unsigned key = (d << 23) + it;
        unsigned m = max(b1, key);
        b1 = min(b1, key);
        b2 = min(b2, m);
        t0w = t0w * 1664525u + 1013904223u;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = b1 ^ b2;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double g_med_cycles = 0;
template <typename K>
static double run(K kernel, int threads, int sms, unsigned *out, long long *cyc, double ops_per_thread_iter, float *ms_out)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kernel<<<sms, threads>>>(out, cyc, 12345u);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kernel<<<sms, threads>>>(out, cyc, 54321u);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(ms_out, e0, e1));
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    double med = (double)h[sms / 2];
    g_med_cycles = med;
    return ops_per_thread_iter * kIters * threads / med;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, sms);
    unsigned *out; long long *cyc;
    CK(cudaMalloc(&out, (size_t)sms * 1024 * 4));
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    float ms;
    const int T = 1024;
#define RUN(OP, name, ops) { double r = run(rate_kernel<OP>, T, sms, out, cyc, ops * kChains, &ms); \
        printf("%-12s %8.2f lanes/clk/SM   (%.3f ms, eff clk %.0f MHz)\n", name, r, ms, g_med_cycles / (ms * 1e-3) / 1e6); }
    RUN(POPC, "popc+iadd", 1.0);
    RUN(POPC_XOR, "xor+popc+add", 1.0);
    RUN(LOP3, "lop3", 1.0);
    RUN(IADD3, "iadd3", 1.0);
    RUN(IMNMX_U32, "imnmx_u32 x2", 2.0);
    RUN(IMNMX3_U32, "vimnmx3_u32", 1.0);
    RUN(FMNMX, "fmnmx x2", 2.0);
    RUN(FMNMX3, "fmnmx3", 1.0);
    RUN(HMNMX2, "hmnmx2", 1.0);
    RUN(IMAD, "imad", 1.0);
    RUN(FFMA, "ffma", 1.0);
    {
        double r = run(cmp_kernel, T, sms, out, cyc, 1.0, &ms);
        double clk_mhz = (double)kIters * T / r / (ms * 1e-3) / 1e6;  // cycles / time
        printf("%-12s %8.4f cmp/clk/SM  => %.3f Tcmp/s at %.0f MHz effective (%.3f ms)\n", "popc_cmp", r,
               r * sms * clk_mhz * 1e6 / 1e12, clk_mhz, ms);
    }
    return 0;
}
