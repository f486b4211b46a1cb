// How many CTAs of a 1-CTA-per-SM kernel (227 KB dynamic shared memory, like score_qs_kernel) can be co-resident
// for cluster sizes 1/2/4/8 on this device: cudaOccupancyMaxActiveClusters.  Evidence for DESIGN.md section 8.1
// (a 4-CTA cluster with TMA multicast strands SMs: GPCs hold 16/18/20 SMs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/cluster_probe tools/cluster_probe.cu && tools/cluster_probe
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(224, 1) probe_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = smem[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int smem = 227 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_active_ctas_by_cluster_size\": {", prop.name, prop.multiProcessorCount);
  const int sizes[4] = {1, 2, 4, 8};
  for (int i = 0; i < 4; ++i) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sizes[i] * 148);
    cfg.blockDim = dim3(224);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = sizes[i]; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("%s\"%d\": %d", i ? ", " : "", sizes[i], e == cudaSuccess ? n * sizes[i] : -1);
  }
  printf("}}\n");
  return 0;
}
