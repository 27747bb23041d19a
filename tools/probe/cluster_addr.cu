// probe: how do shared::cta addresses look inside a 4-CTA cluster, and what does mapa return?
#include <cstdio>
#include <cstdint>
__global__ void __cluster_dims__(4, 1, 1) probe() {
  __shared__ uint64_t bar;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint32_t local = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    uint32_t m[4];
    for (uint32_t t = 0; t < 4; ++t) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(m[t]) : "r"(local), "r"(t));
    printf("block %d rank %u local 0x%08x mapa -> 0x%08x 0x%08x 0x%08x 0x%08x\n", blockIdx.x, rank, local, m[0], m[1], m[2], m[3]);
  }
}
int main() { probe<<<8, 32>>>(); cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError())); int n=0; cudaLaunchConfig_t cfg{}; cfg.gridDim=dim3(148*4); cfg.blockDim=dim3(32); cfg.dynamicSmemBytes=200*1024; cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024); cudaLaunchAttribute at[1]; at[0].id=cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x=4; at[0].val.clusterDim.y=1; at[0].val.clusterDim.z=1; cfg.attrs=at; cfg.numAttrs=1; cudaError_t e=cudaOccupancyMaxActiveClusters(&n, probe, &cfg); printf("max active clusters of 4 with 200KB smem: %d (%s)\n", n, cudaGetErrorString(e)); return 0; }
