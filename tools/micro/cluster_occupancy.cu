// How many thread-block clusters of a given size are co-resident on this part (cudaOccupancyMaxActiveClusters), for the
// shared-memory footprints the weight-stationary rollout pipeline and the split LayerNorm layers need.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 cluster_occupancy.cu -o cluster_occupancy && ./cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe_kernel(float* p) {
  extern __shared__ float sm[];
  if (p && threadIdx.x == 0) p[blockIdx.x] = sm[0];
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int smem_kb[] = {100, 150, 200, 212, 227};
  const int threads[] = {256, 608};
  for (int t : threads)
    for (int kb : smem_kb) {
      cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
      printf("threads %3d smem %3d KB:", t, kb);
      for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(sms / cs * cs > 0 ? sms / cs * cs : cs));
        cfg.blockDim = dim3((unsigned)t);
        cfg.dynamicSmemBytes = (size_t)kb * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
        if (e != cudaSuccess) { n = -1; cudaGetLastError(); }
        printf("  cluster %2d -> %3d clusters (%3d CTAs)", cs, n, n > 0 ? n * cs : 0);
      }
      printf("\n");
    }
  return 0;
}
