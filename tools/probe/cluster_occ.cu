// How many clusters of size 1/2/4/8 with ~200 KB of dynamic shared memory can be resident at once (GPC granularity)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { if (p) p[0] = 1; }
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s SMs=%d\n", pr.name, pr.multiProcessorCount);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pr.multiProcessorCount / cs * cs); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters = %d (%d SMs) %s\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
