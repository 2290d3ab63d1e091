// tc_probe.cu -- standalone validation + rate probe for the tcgen05 building blocks of variant T.
//   (1) correctness: one 128 x 256 x 256 job (+-1 fp8 expansion, 8 x tcgen05.mma K=32, TMEM read-back)
//       against popcount on the host, for the descriptor strides in tc_common.cuh (and alternatives
//       given on the command line: tc_probe <lbo> <sbo>);
//   (2) tensor-pipe rate: back-to-back MMA chains on every SM (clock64 and CUDA events);
//   (3) TMEM -> register read rate (tcgen05.ld 32x32b.x32) with 4 and 8 warps;
//   (4) expansion rate: bits -> fp8 rows in shared memory.
// Used to calibrate DESIGN.md's roofline denominators; not part of the product library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../tc_common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kM = 128, kN = 256;
constexpr uint32_t kABytes = kM * 256, kBBytes = kN * 256;

struct ProbeSmem {
    uint64_t bar;
    uint32_t tmem_base;
};

// mode 0: correctness (dump D).  mode 1: MMA rate (loops).  mode 2: TMEM ld rate.  mode 3: expansion rate.
__global__ void __launch_bounds__(256, 1)
probe_kernel(const uint4 *a_bits, const uint4 *b_bits, float *d_out, long long *cycles, int mode, int loops,
             uint32_t lbo, uint32_t sbo, int ld_warps, uint32_t kstep)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa = smem;
    uint8_t *sb = smem + kABytes;
    ProbeSmem *ps = reinterpret_cast<ProbeSmem *>(smem + kABytes + kBBytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        tc::mbar_init(&ps->bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(&ps->tmem_base, 512);
    // expansion: thread t -> B row t; threads 0..127 also A row t
    long long t_exp0 = clock64();
    int exp_loops = mode == 3 ? loops : 1;
    for (int l = 0; l < exp_loops; ++l) {
        {
            uint4 d0 = b_bits[2 * tid], d1 = b_bits[2 * tid + 1];
            d0.x ^= l;
            tc::expand_row_to_smem(tc::smem_u32(sb), tid, d0, d1);
        }
        if (tid < kM) {
            uint4 d0 = a_bits[2 * tid], d1 = a_bits[2 * tid + 1];
            d0.x ^= l;
            tc::expand_row_to_smem(tc::smem_u32(sa), tid, d0, d1);
        }
    }
    long long t_exp1 = clock64();
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = ps->tmem_base;
    const uint32_t idesc = tc::idesc_e4m3_f32(kM, kN);

    long long t0 = clock64();
    int mma_loops = mode == 1 ? loops : 1;
    if (tid == 0) {
        for (int l = 0; l < mma_loops; ++l) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint64_t ad = tc::smem_desc(tc::smem_u32(sa) + k * kstep, lbo, sbo);
                uint64_t bd = tc::smem_desc(tc::smem_u32(sb) + k * kstep, lbo, sbo);
                tc::umma_f8(tmem + (l & 1) * 256, ad, bd, idesc, k > 0 ? 1u : 0u);
            }
        }
        tc::umma_commit(&ps->bar);
    }
    tc::mbar_wait(&ps->bar, 0, 1);
    long long t1 = clock64();
    tc::tc_fence_after();

    if (mode == 0 && warp < 4) {
        // warp w reads TMEM lanes 32w..32w+31 (= D rows), 8 chunks of 32 columns
        for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * kN + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    long long t2 = clock64(), t3 = t2;
    if (mode == 2) {
        uint32_t acc = 0;
        __syncthreads();
        t2 = clock64();
        if (warp < ld_warps) {
            for (int l = 0; l < loops; ++l) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t v[32];
                    tc::tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((l & 1) * 256) + c * 32, v);
                    acc ^= v[0] ^ v[31];
                }
            }
        }
        t3 = clock64();
        if (acc == 0x12345) d_out[tid] = 1.0f;
    }
    if (tid == 0 && cycles) {
        cycles[blockIdx.x * 4 + 0] = t1 - t0;
        cycles[blockIdx.x * 4 + 1] = t3 - t2;
        cycles[blockIdx.x * 4 + 2] = t_exp1 - t_exp0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static int popc256(const uint32_t *a, const uint32_t *b)
{
    int d = 0;
    for (int w = 0; w < 8; ++w) d += __builtin_popcount(a[w] ^ b[w]);
    return d;
}

int main(int argc, char **argv)
{
    uint32_t lbo = argc > 2 ? (uint32_t)atoi(argv[1]) : tc::kLBO;
    uint32_t sbo = argc > 2 ? (uint32_t)atoi(argv[2]) : tc::kSBO;
    uint32_t kstep = argc > 3 ? (uint32_t)atoi(argv[3]) : 2 * tc::kLBO;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s sm_%d%d SMs %d  lbo=%u sbo=%u\n", prop.name, prop.major, prop.minor, sms, lbo, sbo);

    std::vector<uint32_t> ha(kM * 8), hb(kN * 8);
    srand(7);
    for (auto &x : ha) x = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    for (auto &x : hb) x = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    // make a few rows special: identical, complement
    memcpy(&hb[0], &ha[0], 32);
    for (int w = 0; w < 8; ++w) hb[8 + w] = ~ha[8 + w];
    uint4 *da, *db;
    float *dd;
    long long *dc;
    CK(cudaMalloc(&da, ha.size() * 4));
    CK(cudaMalloc(&db, hb.size() * 4));
    CK(cudaMalloc(&dd, kM * kN * 4));
    CK(cudaMalloc(&dc, sms * 4 * sizeof(long long)));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dd, 0xFF, kM * kN * 4));
    const size_t smem = kABytes + kBBytes + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // (1) correctness
    probe_kernel<<<1, 256, smem>>>(da, db, dd, dc, 0, 1, lbo, sbo, 4, kstep);
    CK(cudaDeviceSynchronize());
    std::vector<float> hd(kM * kN);
    CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < kM; ++i)
        for (int j = 0; j < kN; ++j) {
            float want = 256.0f - 2.0f * popc256(&ha[i * 8], &hb[j * 8]);
            if (hd[i * kN + j] != want) {
                if (bad < 8) printf("  mismatch D[%d][%d] = %g want %g\n", i, j, hd[i * kN + j], want);
                ++bad;
            }
        }
    printf("correctness: %d / %d mismatches  (D[0][0]=%g want 256, D[1][1]=%g want -256)\n", bad, kM * kN, hd[0],
           hd[kN + 1]);
    if (bad) {
        printf("PROBE_FAIL\n");
        return 2;
    }
    printf("PROBE_OK\n");

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<long long> hc(sms * 4);
    // (2) MMA rate on all SMs
    for (int loops : {64, 2048}) {
        probe_kernel<<<sms, 256, smem>>>(da, db, dd, dc, 1, loops, lbo, sbo, 4, kstep);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        probe_kernel<<<sms, 256, smem>>>(da, db, dd, dc, 1, loops, lbo, sbo, 4, kstep);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(hc.data(), dc, hc.size() * 8, cudaMemcpyDeviceToHost));
        std::vector<long long> c;
        for (int s = 0; s < sms; ++s) c.push_back(hc[s * 4]);
        std::sort(c.begin(), c.end());
        double macs = (double)loops * kM * kN * 256;
        printf("mma: loops %d  median %lld cyc/CTA => %.0f MAC/clk/SM, %.1f clk per 128x256x256 job; whole kernel %.3f ms => %.1f TFLOP/s (%.2f Tcmp/s)\n",
               loops, c[sms / 2], macs / c[sms / 2], (double)c[sms / 2] / loops, ms, 2.0 * macs * sms / (ms * 1e-3) / 1e12,
               (double)loops * kM * kN * sms / (ms * 1e-3) / 1e12);
    }
    // (3) TMEM read rate
    for (int w : {4, 8}) {
        const int loops = 512;
        probe_kernel<<<sms, 256, smem>>>(da, db, dd, dc, 2, loops, lbo, sbo, w, kstep);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hc.data(), dc, hc.size() * 8, cudaMemcpyDeviceToHost));
        std::vector<long long> c;
        for (int s = 0; s < sms; ++s) c.push_back(hc[s * 4 + 1]);
        std::sort(c.begin(), c.end());
        double bytes = (double)loops * 8 * 32 * 32 * 4 * w;
        printf("tmem_ld: %d warps  %.1f B/clk/SM  (%.1f fp32 elem/clk/SM)\n", w, bytes / c[sms / 2], bytes / 4 / c[sms / 2]);
    }
    // (4) expansion rate
    {
        const int loops = 256;
        probe_kernel<<<sms, 256, smem>>>(da, db, dd, dc, 3, loops, lbo, sbo, 4, kstep);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hc.data(), dc, hc.size() * 8, cudaMemcpyDeviceToHost));
        std::vector<long long> c;
        for (int s = 0; s < sms; ++s) c.push_back(hc[s * 4 + 2]);
        std::sort(c.begin(), c.end());
        printf("expand: %.1f clk per 384 rows (8 warps) => %.2f clk/row/SM\n", (double)c[sms / 2] / loops,
               (double)c[sms / 2] / loops / 384.0);
    }
    return 0;
}
