// How many thread-block clusters of a given size fit on this GPU at once (GPC floorsweeping decides).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/cluster_probe tools/cluster_probe.cu && tools/cluster_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s, %d SMs\n", pr.name, pr.multiProcessorCount);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int smem : {220 * 1024, 110 * 1024, 72 * 1024, 48 * 1024}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cs : {1, 2, 4, 8, 16}) {
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(cs * 64); lc.blockDim = dim3(256); lc.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &lc);
      printf("smem %3d KB cluster %2d: max active clusters %d (%d CTAs) %s\n", smem / 1024, cs, n, n * cs, e ? cudaGetErrorString(e) : "");
      cudaGetLastError();
    }
  }
  return 0;
}
