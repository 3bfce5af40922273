// tcgen05 issue-rate probe 2 (development tool): which ingredient of the sampler's driver loop slows UMMA issue?
#include <cuda_runtime.h>
#include <stdio.h>
#include "tc_common.cuh"
using namespace gldm::tc;

enum { F_COMMIT4 = 1, F_WAITFULL = 2, F_ELECT = 4, F_OTHERS_MBAR = 8, F_BULK = 16, F_VARADDR = 32, F_FENCE = 64, F_WARP = 128 };

__global__ void __launch_bounds__(256) rate_kernel(long long* out, int reps, int flags, const uint8_t* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, done, ebar[8], fbar[8], gbar[8];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (160 * 1024) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    mbar_init(&bar, 1); mbar_init(&done, 1);
    for (int i = 0; i < 8; ++i) { mbar_init(&ebar[i], 1); mbar_init(&fbar[i], 1); mbar_init(&gbar[i], 1); }
    fence_barrier_init();
  }
  if (wid == 0) tmem_alloc<512>(&slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const bool driver_warp = (flags & F_WARP) ? (__shfl_sync(0xffffffffu, wid, 0) == 7) : false;
  const bool driver_thread = (flags & F_WARP) ? false : (tid == 0);
  if (driver_warp || driver_thread) {
    const uint32_t idesc = idesc_bf16(128, 64);
    const uint32_t ring = smem_u32(smem), bbase = smem_u32(smem) + 131072;
    // pre-complete the "full" barriers once so that waits on parity 0 return immediately
    if (flags & F_WAITFULL) { if (!(flags & F_WARP) || lane == 0) for (int i = 0; i < 8; ++i) mbar_arrive(&fbar[i]); if (flags & F_WARP) __syncwarp(); }
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t s = (flags & F_VARADDR) ? (r & 7) : 0;
      if (flags & F_WAITFULL) { mbar_wait(&fbar[s], 0); }
      if (flags & F_FENCE) tc_fence_after();
      const uint64_t ad = smem_desc(ring + s * 16384, 1024, SW_128);
      const uint64_t bd = smem_desc(bbase + ((flags & F_VARADDR) ? (r % 3) * 2048 : 0), 1024, SW_128);
      const bool go = (flags & F_ELECT) ? elect_one_sync() : true;
      if (go) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem, ad + ks * 2, bd + ks * 2, idesc, 1u);
        if (flags & F_COMMIT4) umma_commit(&ebar[s]);
        if (flags & F_BULK) {
          mbar_arrive_expect_tx(&gbar[r & 7], 8192);   // never waited on; completes by itself
          bulk_g2s(smem + 65536 + (r & 7) * 8192, gsrc + (r & 63) * 16384, 8192, &gbar[r & 7]);
        }
      }
    }
    const long long t1 = clock64();
    if (!(flags & F_WARP) || elect_one_sync()) umma_commit(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    if (lane == 0) mbar_arrive(&ebar[7]);   // dummy
  } else if (flags & F_OTHERS_MBAR) {
    mbar_wait(&done, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem);
}

void run(int flags, const uint8_t* g, const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 1024;
  rate_kernel<<<1, 256, 200 * 1024>>>(d, reps, flags, g);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[2] = {0, 0};
  cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
  printf("%-58s issue %.1f cyc/UMMA, complete %.1f cyc/UMMA %s\n", name, c[0] / (reps * 4.0), c[1] / (reps * 4.0),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  uint8_t* g; cudaMalloc(&g, 2 << 20); cudaMemset(g, 0, 2 << 20);
  run(0, g, "thread0, tight");
  run(F_COMMIT4, g, "thread0 + commit every 4");
  run(F_COMMIT4 | F_WAITFULL, g, "thread0 + commit4 + wait(full)");
  run(F_COMMIT4 | F_WAITFULL | F_FENCE, g, "thread0 + commit4 + wait + fence::after");
  run(F_COMMIT4 | F_WAITFULL | F_FENCE | F_VARADDR, g, "thread0 + ... + varying stage/tap address");
  run(F_COMMIT4 | F_WAITFULL | F_FENCE | F_VARADDR | F_OTHERS_MBAR, g, "thread0 + ... + others parked on mbarrier");
  run(F_WARP | F_ELECT, g, "warp7 elect, tight");
  run(F_WARP | F_ELECT | F_COMMIT4, g, "warp7 elect + commit4");
  run(F_WARP | F_ELECT | F_COMMIT4 | F_WAITFULL | F_FENCE, g, "warp7 elect + commit4 + wait + fence");
  run(F_WARP | F_ELECT | F_COMMIT4 | F_WAITFULL | F_FENCE | F_VARADDR, g, "warp7 elect + ... + varaddr");
  run(F_WARP | F_ELECT | F_COMMIT4 | F_WAITFULL | F_FENCE | F_VARADDR | F_OTHERS_MBAR, g, "warp7 elect + ... + others on mbarrier");
  run(F_WARP | F_ELECT | F_COMMIT4 | F_WAITFULL | F_FENCE | F_VARADDR | F_OTHERS_MBAR | F_BULK, g, "warp7 elect + ... + bulk copies");
  run(F_COMMIT4 | F_WAITFULL | F_FENCE | F_VARADDR | F_OTHERS_MBAR | F_BULK, g, "thread0 + ... + bulk copies");
  return 0;
}
