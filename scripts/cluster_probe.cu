// cluster_probe.cu -- how many clusters of 2 / 4 / 8 CTAs with the tower kernel's footprint (192 threads, ~218 KiB
// dynamic shared memory, 1 CTA per SM) can be co-resident on this GPU?
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(192, 1) k_dummy(int* p) { extern __shared__ int s[]; if (p && threadIdx.x == 9999) p[0] = s[0]; }
int main() {
    const int smem = 218112;
    cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_dummy, &cfg);
        printf("cluster size %2d: max active clusters %d (%d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
