// Probe: how many thread-block clusters of 1/2/4/8/16 CTAs with the MLP kernel's footprint (512 threads, 225 KB of
// shared memory = one CTA per SM) can be resident at once.  B200: 148 / 74 / 33 / 15 / 7 clusters = 148 / 148 / 132 / 120 / 112 SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/occ tools/cluster_occupancy.cu && /tmp/occ
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) dummy(int* p) { extern __shared__ char s[]; if (p) p[0] = s[threadIdx.x]; }
int main() {
  int smem = 225 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster %d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
