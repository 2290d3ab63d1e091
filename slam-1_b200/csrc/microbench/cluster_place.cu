// cluster_place.cu -- where does the hardware put the CTAs of a thread-block cluster?
//
// The single-launch frame kernel (csrc/knn2_frame.cu) runs one cluster of S CTAs per query group.  Its cost model
// needs to know (a) how many clusters of a given size are co-resident and (b) whether CTAs of one launch share SMs
// while other SMs idle.  For grid sizes around the SM count this probe records %smid of every CTA and prints the
// number of distinct SMs used and the largest number of CTAs that landed on one SM, for cluster sizes 1..8 and for
// a small and a "one CTA per SM" shared-memory footprint, together with cudaOccupancyMaxActiveClusters.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) place_kernel(int *smid_out, long long spin_clk)
{
    extern __shared__ int dyn[];
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) {
        smid_out[blockIdx.x] = (int)smid;
        dyn[0] = (int)smid;
    }
    // stay resident long enough that the whole grid must be placed at once
    const long long t0 = clock64();
    while (clock64() - t0 < spin_clk) { }
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d\n", prop.name, prop.multiProcessorCount);
    int *d_smid;
    CK(cudaMalloc(&d_smid, 4096 * sizeof(int)));
    CK(cudaFuncSetAttribute(place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int smems[] = {24 * 1024, 60 * 1024, 120 * 1024};
    const int grids[] = {64, 128, 144, 160, 256, 512};
    printf("%8s %8s %6s | %12s | %10s %14s %14s\n", "smem KB", "cluster", "grid", "max clusters", "SMs used", "max CTAs/SM", "ms (spin 50us)");
    for (int smem : smems) {
        for (int S = 1; S <= 8; S *= 2) {
            cudaLaunchConfig_t cfg{};
            cfg.blockDim = dim3(256, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cfg.gridDim = dim3(S, 1, 1);
            int max_clusters = -1;
            cudaOccupancyMaxActiveClusters(&max_clusters, place_kernel, &cfg);
            for (int g : grids) {
                cfg.gridDim = dim3(g, 1, 1);
                CK(cudaMemset(d_smid, 0xFF, 4096 * sizeof(int)));
                cudaEvent_t a, b;
                CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
                CK(cudaLaunchKernelEx(&cfg, place_kernel, d_smid, 1000LL));   // warm
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(a));
                CK(cudaLaunchKernelEx(&cfg, place_kernel, d_smid, 100000LL));  // ~50 us of spinning per CTA
                CK(cudaEventRecord(b));
                CK(cudaDeviceSynchronize());
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, a, b));
                std::vector<int> h(g);
                CK(cudaMemcpy(h.data(), d_smid, g * sizeof(int), cudaMemcpyDeviceToHost));
                std::vector<int> per(1024, 0);
                for (int v : h) if (v >= 0 && v < 1024) per[v]++;
                int used = 0, mx = 0;
                for (int v : per) { if (v) used++; mx = std::max(mx, v); }
                printf("%8d %8d %6d | %12d | %10d %14d %14.3f\n", smem / 1024, S, g, max_clusters, used, mx, ms);
            }
        }
    }
    return 0;
}
