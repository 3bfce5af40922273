// tcgen05 issue-rate probe (development tool): tight single-thread UMMA loop, descriptors in registers.
#include <cuda_runtime.h>
#include <stdio.h>
#include "tc_common.cuh"
using namespace gldm::tc;

template <int N, int NACC, int M>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc<512>(&slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(M, N);
    const uint64_t ad = smem_desc(smem_u32(smem), 1024, SW_128);            // A: 128 rows x 64 bf16
    const uint64_t bd = smem_desc(smem_u32(smem) + 16384, 1024, SW_128);    // B: N rows x 64 bf16
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int j = 0; j < NACC; ++j) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem + j * N, ad + ks * 2, bd + ks * 2, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tmem);
}

template <int N, int NACC, int M = 128>
void run(int interleave_k) {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel<N, NACC, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  const int reps = 2048;
  rate_kernel<N, NACC, M><<<1, 128, 66 * 1024>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double n = (double)reps * NACC * 4;
  printf("M=%3d N=%3d nacc=%d : %.1f cyc/UMMA  (floor M*N/... = %.1f)  %s\n", M, N, NACC, c / n, 128.0 * N / 256.0,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<16, 1>(0); run<32, 1>(0); run<64, 1>(0); run<128, 1>(0); run<256, 1>(0);
  run<16, 2>(0); run<32, 2>(0); run<64, 2>(0); run<128, 2>(0);
  run<16, 4>(0); run<32, 4>(0); run<64, 4>(0); run<128, 4>(0);
  run<32, 1, 64>(0); run<64, 1, 64>(0); run<32, 4, 64>(0);
  return 0;
}
