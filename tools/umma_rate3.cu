// tcgen05 rate probe 3 (development tool): UMMA throughput vs operand placement in shared memory / CTA size.
#include <cuda_runtime.h>
#include <stdio.h>
#include "tc_common.cuh"
using namespace gldm::tc;

__global__ void rate_kernel(long long* out, int reps, uint32_t a_off, uint32_t b_off, int n, int zero_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, wid = tid >> 5;
  for (int i = tid; i < zero_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&done, 1); fence_barrier_init(); }
  if (wid == 0) tmem_alloc<512>(&slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, n);
    const uint64_t ad = smem_desc(smem_u32(smem) + a_off, 1024, SW_128);
    const uint64_t bd = smem_desc(smem_u32(smem) + b_off, 1024, SW_128);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem, ad + ks * 2, bd + ks * 2, idesc, 1u);
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem);
}

void run(int threads, int smem_kb, uint32_t a_off, uint32_t b_off, int n, int zero_kb) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  const int reps = 2048;
  rate_kernel<<<1, threads, smem_kb * 1024>>>(d, reps, a_off, b_off, n, zero_kb * 1024);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("threads %3d smem %3d KB (zeroed %3d KB) A@%6u B@%6u N=%3d : %.1f cyc/UMMA %s\n", threads, smem_kb, zero_kb, a_off, b_off, n,
         c / (reps * 4.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  run(128, 66, 0, 16384, 64, 64);
  run(256, 66, 0, 16384, 64, 64);
  run(128, 200, 0, 16384, 64, 64);
  run(128, 200, 0, 16384, 64, 200);
  run(128, 200, 0, 131072, 64, 200);
  run(128, 200, 0, 65536, 64, 200);
  run(128, 200, 0, 32768, 64, 200);
  run(128, 200, 0, 24576, 64, 200);
  run(128, 200, 0, 20480, 64, 200);
  run(128, 200, 131072, 0, 64, 200);
  run(128, 200, 51200, 0, 64, 200);
  run(128, 200, 51200 + 16384, 2048, 64, 200);
  run(128, 200, 51200 + 7 * 16384, 12288 + 4096, 64, 200);
  run(128, 200, 0, 16384, 32, 200);
  run(128, 200, 0, 131072, 32, 200);
  run(128, 200, 0, 131072, 128, 200);
  return 0;
}
