// How many thread-block clusters of a given size can be resident at once (640 threads, ~215 KB dynamic smem per CTA)?
// Decides whether a 4-CTA-cluster kernel can cover 33 row tiles in one wave.   nvcc -arch=sm_100a -o cluster_occ cluster_occ.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(640, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = 220 * 1024;
    cudaLaunchAttribute a; a.id = cudaLaunchAttributeClusterDimension; a.val.clusterDim.x = cs; a.val.clusterDim.y = 1; a.val.clusterDim.z = 1;
    cfg.attrs = &a; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
