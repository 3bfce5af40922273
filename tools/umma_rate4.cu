// tcgen05 rate probe 4 (development tool): cost of the warp-collective issue forms used by the sampler driver.
#include <cuda_runtime.h>
#include <stdio.h>
#include "tc_common.cuh"
using namespace gldm::tc;

__global__ void __launch_bounds__(256) rate_kernel(long long* out, int reps, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done, ebar[8], fbar[8];
  __shared__ uint32_t slot;
  __shared__ uint4 tab[8];
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (192 * 1024) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    mbar_init(&done, 1);
    for (int i = 0; i < 8; ++i) { mbar_init(&ebar[i], 1); mbar_init(&fbar[i], 1); }
    fence_barrier_init();
    const uint32_t a_hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
    for (int i = 0; i < 8; ++i)
      tab[i] = make_uint4(0x10000u | ((smem_u32(smem) + i * 16384) >> 4), 0x10000u | ((smem_u32(smem) + 131072 + (i % 3) * 2048) >> 4), a_hi, i);
  }
  if (wid == 0) tmem_alloc<512>(&slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t idesc = idesc_bf16(128, 64);
  const uint64_t ad0 = smem_desc(smem_u32(smem), 1024, SW_128), bd0 = smem_desc(smem_u32(smem) + 131072, 1024, SW_128);
  long long t0 = 0, t1 = 0;
  if (variant == 0) {
    if (tid == 0) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem, ad0 + ks * 2, bd0 + ks * 2, idesc, 1u);
      }
      t1 = clock64();
      umma_commit(&done);
      mbar_wait(&done, 0);
      out[0] = t1 - t0; out[1] = clock64() - t0;
    }
  } else if (__shfl_sync(0xffffffffu, wid, 0) == 7) {
    if (variant >= 4 && lane == 0) for (int i = 0; i < 8; ++i) mbar_arrive(&fbar[i]);
    __syncwarp();
    t0 = clock64();
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
      uint64_t ad = ad0, bd = bd0;
      if (variant >= 2) {
        const uint4 op = tab[r & 7];
        ad = ((uint64_t)op.z << 32) | op.x;
        bd = ((uint64_t)b_hi << 32) | op.y;
      }
      if (variant >= 4) { mbar_wait(&fbar[r & 7], 0); tc_fence_after(); }
      umma_bf16_block_elect<4>(tmem, ad, bd, idesc, 1u);
      if (variant >= 3) umma_commit_elect(&ebar[r & 7]);
    }
    t1 = clock64();
    umma_commit_elect(&done);
    mbar_wait(&done, 0);
    if (lane == 0) { out[0] = t1 - t0; out[1] = clock64() - t0; }
  } else if (variant >= 5) {
    if (variant == 5) mbar_wait(&done, 0);                       // all lanes park on the mbarrier
    else if (variant == 6) { if (lane == 0) mbar_wait(&done, 0); __syncwarp(); }   // one lane per warp
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem);
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  const char* names[] = {"thread 0, plain umma, invariant descriptors", "warp 7, block_elect<4>, invariant descriptors",
                         "warp 7, block_elect<4>, descriptors from smem table", "  + commit_elect per block",
                         "  + wait(full) + fence per block", "  + other 7 warps parked in mbar_wait (all lanes)",
                         "  + other 7 warps parked in mbar_wait (lane 0 only)"};
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int v = 0; v < 7; ++v) {
    long long* d; cudaMalloc(&d, 16);
    const int reps = 2048;
    rate_kernel<<<1, 256, 200 * 1024>>>(d, reps, v);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[2] = {0, 0};
    cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
    printf("%-60s issue %.1f  complete %.1f cyc/UMMA %s\n", names[v], c[0] / (reps * 4.0), c[1] / (reps * 4.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
  }
  return 0;
}
