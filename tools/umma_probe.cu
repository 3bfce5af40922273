// Stand-alone tcgen05 probe (development tool, not part of the library): validates the UMMA descriptor /
// swizzle conventions of csrc/tc_common.cuh against a CPU GEMM and measures UMMA issue throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I graspldm_b200/csrc tools/umma_probe.cu -o /tmp/umma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"

using namespace gldm::tc;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);       \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

// conv-as-GEMM: D[m][n] = sum_{tap, k} A[tap][m][k] * X[k][n + (tap - taps/2) * SG]   (zero outside [0, N))
// A image: [tap][kb][128 rows x ASWB bytes] pre-swizzled; X is written to smem by the threads as the K-major SW128
// B operand with SG-row zero halos.
struct Case {
  int a_swb;   // 32 / 64 / 128
  int kpt;     // K per tap (multiple of 16)
  int taps;    // 1 or 3
  int n;       // N (32 or 64)
  int sg;      // tap shift in rows (8 or 16)
};


__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ a_img, const float* __restrict__ x,
                                                    float* __restrict__ d_out, Case c, int reps, long long* cycles, int nacc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_a, bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t s_desc[64][2];
  const int tid = threadIdx.x, wid = tid >> 5;
  const int nkb = (c.kpt * 2 + c.a_swb - 1) / c.a_swb;          // K-blocks per tap on the A side
  const int a_block = 128 * c.a_swb;                             // bytes per (tap, kb) block
  const int a_bytes = c.taps * nkb * a_block;
  const int rows = c.n + 2 * c.sg;
  const int nkb_b = (c.kpt + 63) / 64;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_bytes + 1023) & ~1023);
  if (tid == 0) {
    mbar_init(&bar_a, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (wid == 0) tmem_alloc<512>(&tmem_slot);
  // B operand: zero everything (halos), then element (row n, channel k) at slab kb = k/64
  for (int i = tid; i < nkb_b * rows * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sB)[i] = 0;
  __syncthreads();
  for (int i = tid; i < c.kpt * c.n; i += 128) {
    const int k = i / c.n, n = i % c.n;
    const int kb = k >> 6, kk = k & 63, r = n + c.sg;
    uint8_t* p = sB + kb * rows * 128 + swz_off<128>(r, kk >> 3) + (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(x[i]);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_a, a_bytes);
    bulk_g2s(sA, a_img, a_bytes, &bar_a);
    mbar_wait(&bar_a, 0);
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(128, c.n);
    const uint32_t a_layout = c.a_swb == 128 ? SW_128 : c.a_swb == 64 ? SW_64 : SW_32;
    int nd = 0;
    for (int tap = 0; tap < c.taps; ++tap) {
      const int shift = (c.taps == 1) ? 1 : tap;      // physical first row = shift * sg
      for (int k0 = 0; k0 < c.kpt; k0 += 16) {
        const int akb = (k0 * 2) / c.a_swb, aoff = (k0 * 2) % c.a_swb;
        s_desc[nd][0] = smem_desc(smem_u32(sA) + (tap * nkb + akb) * a_block + aoff, 8 * c.a_swb, a_layout);
        const int bkb = k0 >> 6, boff = (k0 & 63) * 2;
        s_desc[nd][1] = smem_desc(smem_u32(sB) + bkb * rows * 128 + shift * c.sg * 128 + boff, 1024, SW_128);
        ++nd;
      }
    }
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 4
      for (int i = 0; i < nd; ++i) {
        // nacc > 1: round-robin over independent accumulators (column offsets 64 * j) to expose issue rate vs latency
        umma_bf16(tmem + (nacc > 1 ? 64u * (uint32_t)(i % nacc) : 0u), s_desc[i][0], s_desc[i][1], idesc,
                  nacc > 1 ? (uint32_t)(rep > 0 || i >= nacc) : acc);
        acc = 1;
      }
    }
    umma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  // read back: thread (wid, lane) <-> TMEM lane 32*wid + lane
  for (int c0 = 0; c0 < c.n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(32 * wid) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d_out[(size_t)tid * c.n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem);
}

static float bf16r(float f) { return __bfloat162float(__float2bfloat16(f)); }

static int run_case(const Case& c, int reps, bool verbose, int nacc = 1) {
  const int nkb = (c.kpt * 2 + c.a_swb - 1) / c.a_swb;
  const int a_block = 128 * c.a_swb, a_bytes = c.taps * nkb * a_block;
  std::vector<float> A((size_t)c.taps * 128 * c.kpt), X((size_t)c.kpt * c.n);
  srand(1234 + c.a_swb + c.kpt + c.taps);
  for (auto& v : A) v = bf16r((rand() / (float)RAND_MAX - 0.5f));
  for (auto& v : X) v = bf16r((rand() / (float)RAND_MAX - 0.5f));
  std::vector<uint8_t> img(a_bytes, 0);
  const int epr = c.a_swb / 2;   // elements per row of a block
  for (int tap = 0; tap < c.taps; ++tap)
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < c.kpt; ++k) {
        const int kb = k / epr, kk = k % epr;
        uint32_t off;
        if (c.a_swb == 128) off = swz_off<128>(m, kk >> 3);
        else if (c.a_swb == 64) off = swz_off<64>(m, kk >> 3);
        else off = swz_off<32>(m, kk >> 3);
        __nv_bfloat16 h = __float2bfloat16(A[((size_t)tap * 128 + m) * c.kpt + k]);
        memcpy(&img[(size_t)(tap * nkb + kb) * a_block + off + (kk & 7) * 2], &h, 2);
      }
  uint8_t* d_img;
  float *d_x, *d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_img, a_bytes));
  CK(cudaMalloc(&d_x, X.size() * 4));
  CK(cudaMalloc(&d_out, 128 * c.n * 4));
  CK(cudaMalloc(&d_cyc, 8));
  CK(cudaMemcpy(d_img, img.data(), a_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_x, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  const int rows = c.n + 2 * c.sg, nkb_b = (c.kpt + 63) / 64;
  const size_t smem = ((a_bytes + 1023) & ~1023) + (size_t)nkb_b * rows * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(d_img, d_x, d_out, c, reps, d_cyc, nacc);
  CK(cudaDeviceSynchronize());
  std::vector<float> out(128 * c.n);
  long long cyc;
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < c.n; ++n) {
      double ref = 0;
      for (int tap = 0; tap < c.taps; ++tap) {
        const int nn = n + (c.taps == 1 ? 0 : (tap - 1) * c.sg);
        if (nn < 0 || nn >= c.n) continue;
        for (int k = 0; k < c.kpt; ++k) ref += (double)A[((size_t)tap * 128 + m) * c.kpt + k] * X[(size_t)k * c.n + nn];
      }
      ref *= reps;
      maxerr = fmax(maxerr, fabs(ref - out[m * c.n + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  const int nmma = reps * c.taps * (c.kpt / 16);
  const bool ok = nacc > 1 || maxerr <= 2e-3 * fmax(1.0, maxref);
  if (verbose)
    printf("case a_swb=%3d kpt=%3d taps=%d n=%2d sg=%2d reps=%4d : max|err| %.3e (max|ref| %.2f) %s ; %d UMMAs in %lld cyc = %.1f "
           "cyc/UMMA (nacc %d)\n",
           c.a_swb, c.kpt, c.taps, c.n, c.sg, reps, maxerr, maxref, ok ? "OK" : "MISMATCH", nmma, cyc, (double)cyc / nmma, nacc);
  cudaFree(d_img); cudaFree(d_x); cudaFree(d_out); cudaFree(d_cyc);
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  const Case cases[] = {
      {128, 64, 1, 32, 8},  {128, 64, 3, 32, 8},  {128, 128, 3, 32, 8}, {128, 256, 3, 32, 8}, {64, 32, 1, 32, 8},
      {64, 32, 3, 32, 8},   {32, 16, 1, 32, 8},   {32, 16, 3, 32, 8},   {128, 128, 3, 64, 16}, {128, 64, 1, 16, 8},
      {128, 16, 1, 16, 8},  {128, 128, 1, 64, 16}, {64, 64, 3, 32, 8},
  };
  for (const Case& c : cases) bad += run_case(c, 1, true);
  // issue-rate measurements (results are reps x the single product; fp32 accumulation keeps them comparable)
  bad += run_case({128, 256, 3, 32, 8}, 64, true);
  bad += run_case({128, 128, 3, 64, 16}, 64, true);
  bad += run_case({128, 256, 1, 16, 8}, 64, true);
  bad += run_case({128, 256, 3, 32, 8}, 512, true);
  for (int nacc : {2, 4, 8}) {
    printf("-- %d independent accumulators (values not checked)\n", nacc);
    run_case({128, 256, 3, 32, 8}, 64, true, nacc);
    run_case({128, 128, 3, 64, 16}, 64, true, nacc);
    run_case({128, 256, 1, 16, 8}, 64, true, nacc);
  }
  printf(bad ? "PROBE FAILED (%d)\n" : "PROBE OK\n", bad);
  return bad ? 1 : 0;
}
