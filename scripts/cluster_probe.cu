// cluster_probe.cu -- can 24 clusters of 6 CTAs (one 220 KB CTA per SM, 384 threads) be co-resident on a B200, and does a
// cooperative launch with a cluster dimension work?  (Feasibility probe for a DSMEM tile meeting in k_advi_persist.)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(384, 1) k_probe(unsigned *out, long long *cyc) {
  extern __shared__ unsigned char smem[];
  cg::cluster_group cl = cg::this_cluster();
  cg::grid_group grid = cg::this_grid();
  unsigned *mine = (unsigned *)smem;
  if (threadIdx.x == 0) mine[0] = blockIdx.x;
  cl.sync();
  long long t0 = clock64();
  unsigned s = 0;
  for (unsigned r = 0; r < cl.num_blocks(); ++r) s += *cl.map_shared_rank(mine, r);
  cl.sync();
  long long t1 = clock64();
  grid.sync();
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = s; if (blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; } }
}

int main() {
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {2, 4, 6, 8}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(144 / cs * cs); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    int nclusters = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_probe, &cfg);
    printf("cluster size %d: max active clusters %d (%s) -> %d CTAs\n", cs, nclusters, cudaGetErrorString(e), nclusters * cs);
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 4 * 160); cudaMalloc(&cyc, 16); cudaMemset(out, 0, 4 * 160);
    e = cudaLaunchKernelEx(&cfg, k_probe, out, cyc);
    cudaError_t e2 = cudaDeviceSynchronize();
    unsigned h[160]; long long hc[2] = {0, 0};
    cudaMemcpy(h, out, 4 * 160, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    printf("  cooperative + cluster launch of %d CTAs: %s / %s; block 0 sum %u; cluster round (sync + %d DSMEM reads + sync) %lld cycles, grid.sync %lld cycles\n",
           cfg.gridDim.x, cudaGetErrorString(e), cudaGetErrorString(e2), h[0], cs, hc[0], hc[1]);
    cudaGetLastError();
    cudaFree(out); cudaFree(cyc);
  }
  return 0;
}
