// How many thread-block clusters of size 1/2/4/8 (1 CTA per SM, ~200 KB dynamic smem, 384 threads) can be resident?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    printf("%s SMs=%d smem/block optin=%zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
