// Tensor-core (tcgen05 / TMEM) GEMM for the point-wise layers of the PVCNN encoder:
//   SharedMLP 96->768, 768->1536 (Conv1d k1 + folded BatchNorm + ReLU, R/../pvcnn/modules/shared_mlp.py:18-28) and
//   conv_downscale 1536->768 (R/models/modules/pc_encoders.py:60-67) - 61 % of the encoder's FLOPs.
//
//   Y[m, n] = act(scale[n] * sum_k X[m, k] * W[n, k] + shift[n]),   m = cloud * N_points + point
//
// Operands live in HBM as ready-made UMMA "images": [row tile of 128][K block of 64][128 rows x 128 bytes,
// SWIZZLE_128B] bf16, so a pipeline stage is two 16 KB 1-D bulk copies (no tensor maps) and the epilogue of one layer
// writes the next layer's A image directly.  One CTA computes a 128 x 128 output tile: warp 0 streams the operand
// blocks through a 3-stage mbarrier ring, warp 1 issues the UMMAs (M = 128, N = 128, K = 16, fp32 accumulator in
// TMEM), warps 2-5 run the epilogue (TMEM -> registers -> scale/shift/ReLU -> bf16 -> image).  Two CTAs fit per SM
// (96 KB of shared memory, 128 TMEM columns each), so one CTA's epilogue overlaps the other's main loop.
// NT = 2 (layers whose width is a multiple of 256): a CTA computes 128 x 256 - one N = 256 UMMA per K step on two adjacent
// weight blocks, two 48 KB stages - so every A block fetched from L2 serves twice the output: the 128 x 128 form moved
// 2.4 GB from L2 to shared memory for the 768 -> 1536 layer of 64 clouds (13.5 TB/s, the L2's limit, not the tensor pipe's).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gldm {
using namespace tc;

namespace gtc {
constexpr int BLOCK = 16384;                 // one operand block: 128 rows x 64 bf16
constexpr int NTHREADS = 192;
constexpr int RING = 96 * 1024;              // NT = 1: three 32 KB stages; NT = 2: two 48 KB stages
constexpr int SMEM = RING + 1024 + 256 + 2048;                  // + scale / shift of the CTA's (up to 256) columns
constexpr int SMEM_PROJ = SMEM + 4096;                          // + the [256][4] projection weights of the tile
}  // namespace gtc

struct GemmTcParams {
  const uint8_t* a_img;     // [m_tiles][k_blocks][BLOCK]
  const uint8_t* b_img;     // [n_tiles][k_blocks][BLOCK]
  uint8_t* out_img;         // [m_tiles][n_tiles * 2][BLOCK]  (A image of the next layer, K = n_tiles * 128)
  const float* scale;       // [n] or NULL (1)
  const float* shift;       // [n] or NULL (0)
  int k_blocks, n_tiles, relu;
  // fused projection (PROJ): instead of writing the activation image, the epilogue multiplies the tile's fp32
  // activations by proj_w [128 columns][4] and writes the partial sums proj_out [n_tiles][rows][4]
  const float* proj_w;      // [n_tiles * 128][4] fp32 (output channel fastest, unused channels zero)
  float* proj_out;          // [n_tiles][rows][4] fp32
  long long rows;
};

template <bool PROJ, int NT>
__global__ void __launch_bounds__(gtc::NTHREADS, 2) gemm_tc_kernel(const __grid_constant__ GemmTcParams p) {
  using namespace gtc;
  constexpr int STAGES = NT == 1 ? 3 : 2, STAGE = (1 + NT) * BLOCK, NCOL = 128 * NT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // pointer arithmetic (no integer round trip) keeps the shared address space visible to the compiler: LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nt = blockIdx.x, mt = blockIdx.y;
  const int kb_n = p.k_blocks;
  // per-column scale / shift of this CTA's tile, staged once: a global load per element in the epilogue sits behind the
  // bulk-copy traffic (the same finding as in the Conv3d epilogue)
  float* s_sc = reinterpret_cast<float*>(smem + RING + 256);
  float* s_sh = s_sc + 256;
  for (int i = tid; i < NCOL; i += NTHREADS) {
    s_sc[i] = p.scale ? __ldg(p.scale + nt * NCOL + i) : 1.f;
    s_sh[i] = p.shift ? __ldg(p.shift + nt * NCOL + i) : 0.f;
  }
  float4* s_pw = reinterpret_cast<float4*>(s_sh + 256);
  if (PROJ)
    for (int i = tid; i < NCOL; i += NTHREADS) s_pw[i] = __ldg(reinterpret_cast<const float4*>(p.proj_w) + nt * NCOL + i);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (wid == 1) tmem_alloc<NCOL>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
    // ---- producer
    const uint8_t* a = p.a_img + (size_t)mt * kb_n * BLOCK;
    const uint8_t* b = p.b_img + (size_t)(nt * NT) * kb_n * BLOCK;      // NT adjacent 128-row weight tiles
#pragma unroll 1
    for (int kb = 0; kb < kb_n; ++kb) {
      const int s = kb % STAGES, round = kb / STAGES;
      if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[s], STAGE);
        bulk_g2s(smem + s * STAGE, a + (size_t)kb * BLOCK, BLOCK, &full[s]);
#pragma unroll
        for (int t = 0; t < NT; ++t)      // rows 128 t .. of the N = 128 NT operand: the same 1024-byte 8-row groups, contiguous
          bulk_g2s(smem + s * STAGE + (1 + t) * BLOCK, b + ((size_t)t * kb_n + kb) * BLOCK, BLOCK, &full[s]);
      }
      __syncwarp();
    }
  } else if (wid == 1) {
    // ---- UMMA issuer
    const uint32_t idesc = idesc_bf16(128, NCOL);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
    const uint32_t base = smem_u32(smem);
#pragma unroll 1
    for (int kb = 0; kb < kb_n; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full[s], (kb / STAGES) & 1);
      tc_fence_after();
      const uint64_t ad = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGE) >> 4));
      const uint64_t bd = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGE + BLOCK) >> 4));
      umma_bf16_block_elect<4>(tmem, ad, bd, idesc, kb != 0);
      umma_commit_elect(&empty[s]);
    }
    umma_commit_elect(acc_full);
  } else {
    // ---- epilogue: warp w reads TMEM lanes 32 * (w % 4); thread <-> output row
    const int q = wid & 3, r = q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
    if (PROJ) {
      // y = act(scale * acc + shift) stays fp32 in registers; four running dot products per row, columns in order
      float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int c0 = 0; c0 < NCOL; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float y = fmaf(__uint_as_float(v[j]), s_sc[c0 + j], s_sh[c0 + j]);
          if (p.relu) y = fmaxf(y, 0.f);
          const float4 w = s_pw[c0 + j];
          dot.x = fmaf(y, w.x, dot.x);
          dot.y = fmaf(y, w.y, dot.y);
          dot.z = fmaf(y, w.z, dot.z);
          dot.w = fmaf(y, w.w, dot.w);
        }
      }
      reinterpret_cast<float4*>(p.proj_out)[(size_t)nt * p.rows + (size_t)mt * 128 + r] = dot;
    } else {
      uint8_t* out_tile = p.out_img + ((size_t)mt * (p.n_tiles * 2) + (size_t)nt * 2 * NT) * BLOCK;
#pragma unroll 1
      for (int c0 = 0; c0 < NCOL; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t pk[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int n = c0 + j8 * 8 + 2 * h;
            float y0 = __uint_as_float(v[j8 * 8 + 2 * h]), y1 = __uint_as_float(v[j8 * 8 + 2 * h + 1]);
            const float2 sc = *reinterpret_cast<const float2*>(s_sc + n), sh = *reinterpret_cast<const float2*>(s_sh + n);
            y0 = fmaf(y0, sc.x, sh.x);
            y1 = fmaf(y1, sc.y, sh.y);
            if (p.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
            pk[h] = pack_bf16(y0, y1);
          }
          const int c = c0 + j8 * 8;                         // column inside the CTA's tile
          uint8_t* dst = out_tile + (size_t)(c >> 6) * BLOCK + swz_off<128>(r, (c & 63) >> 3);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (wid == 1) tmem_dealloc<NCOL>(tmem);
}

// Persistent form for large row counts: one CTA per SM walks 128 x 256 output tiles (the N tiles of a row tile are
// consecutive, so its A blocks stay in L2), four 48 KB stages, TWO 256-column TMEM accumulators - the eight epilogue warps
// (two per TMEM lane quarter, 128 columns each) finish tile i while the tensor pipe works on tile i + 1.  The one-tile
// CTAs above (two per SM) left the tensor pipe idle whenever both were outside their main loop (TMEM allocation, first
// loads, epilogue): 52 % active on the 768 -> 1536 layer.
namespace gtc {
constexpr int P_STAGES = 4, P_STAGE = 3 * BLOCK, P_THREADS = 320;
constexpr int P_PARAMS = 2 * (2 * 256 * 4 + 256 * 16);          // two buffers of scale | shift | projection weights
constexpr int P_SMEM = P_STAGES * P_STAGE + P_PARAMS + 256 + 1024;
}  // namespace gtc

template <bool PROJ>
__global__ void __launch_bounds__(gtc::P_THREADS, 1) gemm_tc_persistent_kernel(const __grid_constant__ GemmTcParams p, int n_work) {
  using namespace gtc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_par = smem + P_STAGES * P_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_par + P_PARAMS);
  uint64_t* full = bars;             // [4]
  uint64_t* empty = bars + 4;        // [4]
  uint64_t* acc_full = bars + 8;     // [2]
  uint64_t* acc_empty = bars + 10;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int kb_n = p.k_blocks, n2 = p.n_tiles >> 1;       // 256-column tiles per row tile

  if (tid == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    fence_barrier_init();
  }
  if (wid == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
    // ---- producer
    int it = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int mt = w / n2, nt = w - mt * n2;
      const uint8_t* a = p.a_img + (size_t)mt * kb_n * BLOCK;
      const uint8_t* b = p.b_img + (size_t)(nt * 2) * kb_n * BLOCK;
#pragma unroll 1
      for (int kb = 0; kb < kb_n; ++kb, ++it) {
        const int s = it % P_STAGES, round = it / P_STAGES;
        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[s], P_STAGE);
          bulk_g2s(smem + s * P_STAGE, a + (size_t)kb * BLOCK, BLOCK, &full[s]);
          bulk_g2s(smem + s * P_STAGE + BLOCK, b + (size_t)kb * BLOCK, BLOCK, &full[s]);
          bulk_g2s(smem + s * P_STAGE + 2 * BLOCK, b + ((size_t)kb_n + kb) * BLOCK, BLOCK, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (wid == 1) {
    // ---- UMMA issuer
    const uint32_t idesc = idesc_bf16(128, 256);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
    const uint32_t base = smem_u32(smem);
    int it = 0, wl = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++wl) {
      const int buf = wl & 1;
      if (wl >= 2) { mbar_wait(&acc_empty[buf], ((wl >> 1) - 1) & 1); tc_fence_after(); }
#pragma unroll 1
      for (int kb = 0; kb < kb_n; ++kb, ++it) {
        const int s = it % P_STAGES;
        mbar_wait(&full[s], (it / P_STAGES) & 1);
        tc_fence_after();
        const uint64_t ad = ((uint64_t)hi << 32) | (0x10000u | ((base + s * P_STAGE) >> 4));
        const uint64_t bd = ((uint64_t)hi << 32) | (0x10000u | ((base + s * P_STAGE + BLOCK) >> 4));
        umma_bf16_block_elect<4>(tmem + buf * 256, ad, bd, idesc, kb != 0);
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(&acc_full[buf]);
    }
  } else {
    // ---- epilogue: warps 2..9; warp w reads TMEM lanes 32 * (w % 4), columns 128 * half; thread <-> output row
    const int ew = wid - 2, q = wid & 3, half = ew >> 2, r = q * 32 + lane, et = tid - 64;
    int wl = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++wl) {
      const int buf = wl & 1;
      const int mt = w / n2, nt = w - mt * n2;
      // this tile's per-column constants (the buffer was last read two tiles ago; the barrier below orders those reads)
      float* s_sc = reinterpret_cast<float*>(s_par + buf * (P_PARAMS / 2));
      float* s_sh = s_sc + 256;
      float4* s_pw = reinterpret_cast<float4*>(s_sh + 256);
      s_sc[et] = p.scale ? __ldg(p.scale + nt * 256 + et) : 1.f;
      s_sh[et] = p.shift ? __ldg(p.shift + nt * 256 + et) : 0.f;
      if (PROJ) s_pw[et] = __ldg(reinterpret_cast<const float4*>(p.proj_w) + nt * 256 + et);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&acc_full[buf], (wl >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem + buf * 256 + half * 128 + ((uint32_t)(q * 32) << 16);
      const float* sc = s_sc + half * 128;
      const float* sh = s_sh + half * 128;
      if (PROJ) {
        const float4* pw = s_pw + half * 128;
        float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = fmaf(__uint_as_float(v[j]), sc[c0 + j], sh[c0 + j]);
            if (p.relu) y = fmaxf(y, 0.f);
            const float4 wv = pw[c0 + j];
            dot.x = fmaf(y, wv.x, dot.x);
            dot.y = fmaf(y, wv.y, dot.y);
            dot.z = fmaf(y, wv.z, dot.z);
            dot.w = fmaf(y, wv.w, dot.w);
          }
        }
        // partial-sum planes of 128 columns, as the one-tile kernel with NT = 1 writes them
        reinterpret_cast<float4*>(p.proj_out)[(size_t)(nt * 2 + half) * p.rows + (size_t)mt * 128 + r] = dot;
      } else {
        uint8_t* out_tile = p.out_img + ((size_t)mt * (p.n_tiles * 2) + (size_t)nt * 4 + half * 2) * BLOCK;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            uint32_t pk[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int n = c0 + j8 * 8 + 2 * h;
              float y0 = __uint_as_float(v[j8 * 8 + 2 * h]), y1 = __uint_as_float(v[j8 * 8 + 2 * h + 1]);
              const float2 s2 = *reinterpret_cast<const float2*>(sc + n), h2 = *reinterpret_cast<const float2*>(sh + n);
              y0 = fmaf(y0, s2.x, h2.x);
              y1 = fmaf(y1, s2.y, h2.y);
              if (p.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
              pk[h] = pack_bf16(y0, y1);
            }
            const int c = c0 + j8 * 8;                         // column inside this warp's 128
            uint8_t* dst = out_tile + (size_t)(c >> 6) * BLOCK + swz_off<128>(r, (c & 63) >> 3);
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 1) tmem_dealloc<512>(tmem);
}

// fp32 channel-major activations [b, c, n] -> A image (rows m = b * n + point, K = c padded to a multiple of 64)
__global__ void __launch_bounds__(256) to_image_kernel(const float* __restrict__ x, uint8_t* __restrict__ img, int c, int n,
                                                       long long rows, int k_blocks) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;                 // 8 channels
  if (m >= rows) return;
  const long long b = m / n;
  const int pt = (int)(m - b * n);
  const float* xb = x + ((size_t)b * c) * n + pt;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = chunk * 8 + j;
    v[j] = ch < c ? __ldg(xb + (size_t)ch * n) : 0.f;
  }
  const long long mt = m >> 7;
  const int r = (int)(m & 127), kb = chunk >> 3;
  uint8_t* dst = img + ((size_t)mt * k_blocks + kb) * gtc::BLOCK + swz_off<128>(r, chunk & 7);
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                              pack_bf16(v[6], v[7]));
}

// few-output point-wise conv reading an A image: y[b, o, pt] = sum_k W[o, k] * X[m, k] + bias[o]   (out_layer.0)
template <int CO>
__global__ void __launch_bounds__(128) image_small_co_kernel(const uint8_t* __restrict__ img, const float* __restrict__ w,
                                                             const float* __restrict__ bias, int k, int k_blocks, int n,
                                                             long long rows, float* __restrict__ y) {
  extern __shared__ float s_w[];   // [CO][k]
  for (int i = threadIdx.x; i < CO * k; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  const long long mt = m >> 7;
  const int r = (int)(m & 127);
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = 0.f;
  for (int kb = 0; kb < k_blocks; ++kb) {
    const uint8_t* row = img + ((size_t)mt * k_blocks + kb) * gtc::BLOCK + (size_t)r * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {                      // logical chunk j sits at physical chunk j ^ (r & 7)
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + ((j ^ (r & 7)) << 4)));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uu[h]));
        const int kk = kb * 64 + j * 8 + 2 * h;
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[o] = fmaf(s_w[o * k + kk + 1], f.y, fmaf(s_w[o * k + kk], f.x, acc[o]));
      }
    }
  }
  const long long b = m / n;
  const int pt = (int)(m - b * n);
#pragma unroll
  for (int o = 0; o < CO; ++o) y[((size_t)b * CO + o) * n + pt] = acc[o] + (bias ? bias[o] : 0.f);
}

// sum of the per-tile partial projections (fixed order: deterministic) + bias, to the reference's [b, co, n] layout
__global__ void __launch_bounds__(256) proj_sum_kernel(const float4* __restrict__ part, const float* __restrict__ bias,
                                                       int n_tiles, long long rows, int co, int n, float* __restrict__ y) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  float4 a = part[m];
  for (int t = 1; t < n_tiles; ++t) {
    const float4 v = part[(size_t)t * rows + m];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  const long long b = m / n;
  const int pt = (int)(m - b * n);
  const float o[4] = {a.x, a.y, a.z, a.w};
  for (int c = 0; c < co; ++c) y[((size_t)b * co + c) * n + pt] = o[c] + (bias ? bias[c] : 0.f);
}

// weight matrix fp32 [n_out][k] -> B image [n_tiles][k_blocks][BLOCK]
__global__ void __launch_bounds__(256) weight_image_kernel(const float* __restrict__ w, uint8_t* __restrict__ img, int n_out,
                                                           int k, int k_blocks) {
  const int nrow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;   // one warp per row
  const int n_tiles = gridDim.x * (blockDim.x >> 5) / 128;
  if (nrow >= n_tiles * 128) return;
  for (int chunk = lane; chunk < k_blocks * 8; chunk += 32) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = chunk * 8 + j;
      v[j] = (nrow < n_out && kk < k) ? w[(size_t)nrow * k + kk] : 0.f;
    }
    uint8_t* dst = img + ((size_t)(nrow >> 7) * k_blocks + (chunk >> 3)) * gtc::BLOCK + swz_off<128>(nrow & 127, chunk & 7);
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                                pack_bf16(v[6], v[7]));
  }
}

}  // namespace gldm

using namespace gldm;

extern "C" long long gldm_gemm_tc_image_bytes(long long rows, int k) {
  if (rows <= 0 || k <= 0) return -1;
  return ((rows + 127) / 128) * (long long)((k + 63) / 64) * gtc::BLOCK;
}

extern "C" int gldm_gemm_tc_pack_weight(const float* w, int n_out, int k, void* img, void* stream) {
  GLDM_REQUIRE(w && img, "gemm_tc_pack_weight: null pointer");
  GLDM_REQUIRE(n_out > 0 && k > 0, "gemm_tc_pack_weight: bad sizes");
  const int n_tiles = (n_out + 127) / 128, kb = (k + 63) / 64;
  weight_image_kernel<<<n_tiles * 128 / 8, 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<uint8_t*>(img), n_out, k, kb);
  return check_launch("weight_image_kernel");
}

extern "C" int gldm_gemm_tc_to_image(const float* x, int b, int c, int n, void* img, void* stream) {
  GLDM_REQUIRE(x && img, "gemm_tc_to_image: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && n > 0, "gemm_tc_to_image: bad sizes");
  const long long rows = (long long)b * n;
  GLDM_REQUIRE(rows % 128 == 0, "gemm_tc_to_image: b * n = %lld must be a multiple of 128", rows);
  if (rows == 0) return GLDM_OK;
  const int kb = (c + 63) / 64;
  dim3 grid((unsigned)((rows + 255) / 256), kb * 8);
  to_image_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<uint8_t*>(img), c, n, rows, kb);
  return check_launch("to_image_kernel");
}

// N tiles per CTA: 2 when the layer width allows it (GLDM_GEMM_NT=1 keeps the 128 x 128 form)
static int gemm_nt(int n_out) {
  const char* ev = getenv("GLDM_GEMM_NT");                 // read per call (tests switch it)
  return ((ev ? atoi(ev) : 2) == 2 && n_out % 256 == 0) ? 2 : 1;
}

// persistent kernel: widths that are multiples of 256 and at least two 128 x 256 tiles per SM (GLDM_GEMM_PERSISTENT: 0 =
// never, 2 = whenever the width allows, for tests; read per call)
static bool gemm_persistent(int n_out, long long rows) {
  const char* ev = getenv("GLDM_GEMM_PERSISTENT");
  const int mode = ev ? atoi(ev) : 1;
  if (mode == 0 || n_out % 256 != 0 || rows / 128 * (n_out / 256) > 0x7fffffffLL) return false;
  return mode == 2 || rows / 128 * (n_out / 256) >= 2 * kNumSMs;
}

extern "C" int gldm_gemm_tc_run(const void* a_img, const void* w_img, const float* scale, const float* shift,
                                long long rows, int k, int n_out, int relu, void* out_img, void* stream) {
  GLDM_REQUIRE(a_img && w_img && out_img, "gemm_tc_run: null pointer");
  GLDM_REQUIRE(rows >= 0 && rows % 128 == 0, "gemm_tc_run: rows = %lld must be a multiple of 128", rows);
  GLDM_REQUIRE(k > 0 && n_out > 0 && n_out % 128 == 0, "gemm_tc_run: n_out = %d must be a multiple of 128", n_out);
  if (rows == 0) return GLDM_OK;
  GemmTcParams p = {};
  p.a_img = reinterpret_cast<const uint8_t*>(a_img);
  p.b_img = reinterpret_cast<const uint8_t*>(w_img);
  p.out_img = reinterpret_cast<uint8_t*>(out_img);
  p.scale = scale; p.shift = shift;
  p.k_blocks = (k + 63) / 64; p.n_tiles = n_out / 128; p.relu = relu;
  if (gemm_persistent(n_out, rows)) {
    static SmemOptIn attrp;
    if (int rc = opt_in_smem(attrp, gemm_tc_persistent_kernel<false>, gtc::P_SMEM, "gemm_tc_persistent_kernel")) return rc;
    const int n_work = (int)(rows / 128) * (n_out / 256);
    gemm_tc_persistent_kernel<false><<<min(n_work, kNumSMs), gtc::P_THREADS, gtc::P_SMEM, (cudaStream_t)stream>>>(p, n_work);
    return check_launch("gemm_tc_persistent_kernel");
  }
  if (gemm_nt(n_out) == 2) {
    static SmemOptIn attr2;
    if (int rc = opt_in_smem(attr2, gemm_tc_kernel<false, 2>, gtc::SMEM, "gemm_tc_kernel")) return rc;
    gemm_tc_kernel<false, 2><<<dim3(p.n_tiles / 2, (unsigned)(rows / 128)), gtc::NTHREADS, gtc::SMEM, (cudaStream_t)stream>>>(p);
  } else {
    static SmemOptIn attr;
    if (int rc = opt_in_smem(attr, gemm_tc_kernel<false, 1>, gtc::SMEM, "gemm_tc_kernel")) return rc;
    gemm_tc_kernel<false, 1><<<dim3(p.n_tiles, (unsigned)(rows / 128)), gtc::NTHREADS, gtc::SMEM, (cudaStream_t)stream>>>(p);
  }
  return check_launch("gemm_tc_kernel");
}

extern "C" int gldm_gemm_tc_run_proj(const void* a_img, const void* w_img, const float* scale, const float* shift,
                                     long long rows, int k, int n_out, int relu, const float* proj_w, const float* proj_bias,
                                     int co, int n, void* partials, float* y, void* stream) {
  GLDM_REQUIRE(a_img && w_img && proj_w && partials && y, "gemm_tc_run_proj: null pointer");
  GLDM_REQUIRE(rows >= 0 && rows % 128 == 0, "gemm_tc_run_proj: rows = %lld must be a multiple of 128", rows);
  GLDM_REQUIRE(k > 0 && n_out > 0 && n_out % 128 == 0, "gemm_tc_run_proj: n_out = %d must be a multiple of 128", n_out);
  GLDM_REQUIRE(co >= 1 && co <= 4 && n > 0 && rows % n == 0, "gemm_tc_run_proj: bad projection sizes co=%d n=%d", co, n);
  GLDM_REQUIRE((reinterpret_cast<uintptr_t>(proj_w) & 15) == 0 && (reinterpret_cast<uintptr_t>(partials) & 15) == 0,
               "gemm_tc_run_proj: proj_w / partials must be 16-byte aligned");
  if (rows == 0) return GLDM_OK;
  GemmTcParams p = {};
  p.a_img = reinterpret_cast<const uint8_t*>(a_img);
  p.b_img = reinterpret_cast<const uint8_t*>(w_img);
  p.scale = scale; p.shift = shift;
  p.k_blocks = (k + 63) / 64; p.n_tiles = n_out / 128; p.relu = relu;
  p.proj_w = proj_w; p.proj_out = reinterpret_cast<float*>(partials); p.rows = rows;
  cudaStream_t s = (cudaStream_t)stream;
  int nt = gemm_nt(n_out);
  if (gemm_persistent(n_out, rows)) {
    static SmemOptIn attrp;
    if (int rc = opt_in_smem(attrp, gemm_tc_persistent_kernel<true>, gtc::P_SMEM, "gemm_tc_persistent_kernel<proj>")) return rc;
    const int n_work = (int)(rows / 128) * (n_out / 256);
    gemm_tc_persistent_kernel<true><<<min(n_work, kNumSMs), gtc::P_THREADS, gtc::P_SMEM, s>>>(p, n_work);
    nt = 1;                              // (partial-sum planes of 128 columns)
  } else if (nt == 2) {
    static SmemOptIn attr2;
    if (int rc = opt_in_smem(attr2, gemm_tc_kernel<true, 2>, gtc::SMEM_PROJ, "gemm_tc_kernel<proj>")) return rc;
    gemm_tc_kernel<true, 2><<<dim3(p.n_tiles / 2, (unsigned)(rows / 128)), gtc::NTHREADS, gtc::SMEM_PROJ, s>>>(p);
  } else {
    static SmemOptIn attr;
    if (int rc = opt_in_smem(attr, gemm_tc_kernel<true, 1>, gtc::SMEM_PROJ, "gemm_tc_kernel<proj>")) return rc;
    gemm_tc_kernel<true, 1><<<dim3(p.n_tiles, (unsigned)(rows / 128)), gtc::NTHREADS, gtc::SMEM_PROJ, s>>>(p);
  }
  if (int rc = check_launch("gemm_tc_kernel<proj>")) return rc;
  proj_sum_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(partials), proj_bias, p.n_tiles / nt,
                                                                  rows, co, n, y);
  return check_launch("proj_sum_kernel");
}

extern "C" int gldm_gemm_tc_image_small_co(const void* img, const float* w, const float* bias, long long rows, int k,
                                           int co, int n, float* y, void* stream) {
  GLDM_REQUIRE(img && w && y, "gemm_tc_image_small_co: null pointer");
  GLDM_REQUIRE(rows >= 0 && k > 0 && k % 64 == 0 && co >= 1 && co <= 4 && n > 0 && rows % n == 0,
               "gemm_tc_image_small_co: bad sizes");
  if (rows == 0) return GLDM_OK;
  const int kb = k / 64;
  const size_t smem = sizeof(float) * co * k;
  dim3 grid((unsigned)((rows + 127) / 128));
  const uint8_t* im = reinterpret_cast<const uint8_t*>(img);
  cudaStream_t s = (cudaStream_t)stream;
  switch (co) {
    case 1: image_small_co_kernel<1><<<grid, 128, smem, s>>>(im, w, bias, k, kb, n, rows, y); break;
    case 2: image_small_co_kernel<2><<<grid, 128, smem, s>>>(im, w, bias, k, kb, n, rows, y); break;
    case 3: image_small_co_kernel<3><<<grid, 128, smem, s>>>(im, w, bias, k, kb, n, rows, y); break;
    default: image_small_co_kernel<4><<<grid, 128, smem, s>>>(im, w, bias, k, kb, n, rows, y); break;
  }
  return check_launch("image_small_co_kernel");
}
