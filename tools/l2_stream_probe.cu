// L2 -> shared-memory weight-streaming probe (development tool): every CTA (or cluster, with multicast) streams the
// same `bytes` buffer `passes` times through a 4-stage ring of 16 KB bulk copies.  Reports bytes/cycle per SM.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "tc_common.cuh"
using namespace gldm::tc;
namespace cg = cooperative_groups;

#ifndef CHUNK_B
#define CHUNK_B 16384
#endif
#ifndef NSTAGES
#define NSTAGES 4
#endif
constexpr int CHUNK = CHUNK_B, STAGES = NSTAGES;

__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void remote_arrive(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

template <int CSZ>
__global__ void __launch_bounds__(64) stream_kernel(const uint8_t* __restrict__ w, int chunks_per_pass, int passes,
                                                    long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x;
  uint32_t rank = 0;
  if (CSZ > 1) rank = cg::this_cluster().block_rank();
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CSZ); }
    fence_barrier_init();
  }
  if (CSZ > 1) cg::this_cluster().sync(); else __syncthreads();
  const int total = chunks_per_pass * passes;
  const long long t0 = clock64();
  if (tid == 0) {
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
      mbar_arrive_expect_tx(&full[s], CHUNK);
      const uint8_t* src = w + (size_t)(it % chunks_per_pass) * CHUNK;
      if (CSZ == 1) bulk_g2s(smem + s * CHUNK, src, CHUNK, &full[s]);
      else {
        constexpr int SL = CHUNK / CSZ;
        bulk_g2s_mc(smem + s * CHUNK + rank * SL, src + rank * SL, SL, &full[s], (uint16_t)((1u << CSZ) - 1));
      }
    }
  } else if (tid == 32) {
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      if (CSZ == 1) mbar_arrive(&empty[s]);
      else for (uint32_t r = 0; r < CSZ; ++r) remote_arrive(&empty[s], r);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (CSZ > 1) cg::this_cluster().sync();
}

template <int CSZ>
void run(const uint8_t* w, int bytes, int passes, int grid) {
  long long* d;
  cudaMalloc(&d, 8 * grid);
  cudaFuncSetAttribute(stream_kernel<CSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK + 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = STAGES * CHUNK + 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {CSZ, 1, 1};
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(a);
    cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel<CSZ>, w, bytes / CHUNK, passes, d);
    cudaEventRecord(b);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("csz %d: %s %s\n", CSZ, cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
  }
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h[256]; cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per_cta = (double)bytes * passes;
  printf("cluster %d grid %3d: %.3f ms, max %lld cyc; per-SM delivered %.1f B/cyc, chip delivered %.0f B/cyc (%.2f TB/s), L2 reads/cyc (if no dedup) %.0f\n",
         CSZ, grid, ms, mx, per_cta / mx, per_cta * grid / mx, per_cta * grid / (ms * 1e-3) / 1e12, per_cta * grid / CSZ / mx);
  cudaFree(d);
}

int main() {
  const int bytes = 2 * 1024 * 1024, passes = 100;
  uint8_t* w; cudaMalloc(&w, bytes); cudaMemset(w, 1, bytes);
  printf("chunk %d x %d stages\n", CHUNK, STAGES);
  run<1>(w, bytes, passes, 1);
  run<1>(w, bytes, passes, 80);
  run<1>(w, bytes, passes, 148);
  return 0;
}
