// How many clusters of 1/2/4/8 CTAs (one 227 KB CTA per SM, 320 threads) can be co-resident on this GPU
// (GPC shapes decide; diagnostics for the scan kernel's cluster size).  nvcc -arch=sm_100a cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) dummy(int* p) { extern __shared__ int s[]; if (p) s[threadIdx.x] = *p; }
int main() {
  const int smem = 227 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cl : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl * 64); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %3d (= %3d CTAs)  %s\n", cl, n, n * cl, cudaGetErrorString(e));
  }
  return 0;
}
