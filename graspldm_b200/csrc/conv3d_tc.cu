// Tensor-core (tcgen05 / TMEM) Conv3d k3 p1 of the PVConv voxel branch (R/../pvcnn/modules/pvconv.py:48-67),
// bf16 operands / fp32 accumulation, as an implicit GEMM over a zero-padded channels-last voxel grid.
//
// The input grid is X[b * P + p][Cpad] bf16 with P = (r+2)^3 padded voxels per cloud (halo = 0) and Cpad = channels
// rounded up to 64 (16 for the 3-channel first layer); the voxelisation or the previous layer's epilogue writes it in
// that form.  For a filter tap (dx,dy,dz) the A operand of a tile of 128 consecutive padded voxels is then simply the
// SAME matrix shifted by dx*(r+2)^2 + dy*(r+2) + dz rows, which one 2-D TMA tensor load fetches (SWIZZLE_128B / 32B,
// out-of-range rows read as zero) - no im2col buffer.  The weights are pre-packed UMMA images [tap][K block][128 rows x
// 128 B].  One tile = 128 padded voxels x all output channels (UMMA M = 128, N = C_out <= 128): warp 0 producer, warp 1
// UMMA issuer, warps 2-9 epilogue (bias, drop halo voxels).  Two output forms: fp32 in the reference's [b, c, r^3] layout,
// or - for the fused voxel branch - the same zero-padded channels-last grid the next Conv3d reads (bf16, or fp32 for the
// devoxelize input) together with per-tile GroupNorm partial sums (no atomics; conv_stats_finalize_kernel adds the tiles
// of a cloud in a fixed order), so that GroupNorm + Swish becomes one in-place pass (gn_swish_cl_kernel) and devoxelize
// gathers channels-last rows (devox_cl_kernel).
//
// Kernels, by layer (the launchers pick; INTEGRATION.md section 5 lists the switches):
//   conv3d_tc_kernel      one tile per CTA, one activation tile per tap (first version; fp32 [b, c, r^3] output path)
//   conv3d_tc3_kernel     one tile per CTA, one activation tile per filter column (three dz taps per load)
//   conv3d_tc3p_kernel    persistent, the whole filter bank resident (48 -> 48 at 24^3), one box per dx plane
//   conv3d_tc3m_kernel    persistent, streamed weights, two tiles per weight stage (48 -> 96, 96 -> 96 at 12^3)
//   conv3d_tc16_kernel    the 3 -> 48 first layer (16-channel rows, K = 16 per tap), one tile per CTA
//   conv3d_tc16p_kernel   the same, persistent and weight-stationary
//   conv3d_tc16g_kernel   the same with two tiles' epilogues in flight (default for large grids)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gldm {
using namespace tc;

namespace c3 {
constexpr int A_BYTES = 16384;              // 128 voxels x 64 channels bf16
constexpr int W_BYTES = 16384;              // one weight block in the pack
constexpr int W_SLOT = 12288;               // shared-memory slot of a weight block: rows 0..95 (co <= 96), 16384 otherwise
constexpr int STAGES = 3;                   // 3 x 28 KB: two CTAs per SM overlap each other's load latency and epilogue
constexpr int NTHREADS = 320;               // producer, UMMA issuer, 8 epilogue warps
}  // namespace c3

struct Conv3dTcParams {
  const uint8_t* w_img;    // [27][k_blocks][16384]  (rows = output channels, zero padded to 128)
  const float* bias;       // [co] or NULL
  float* y;                // [b, co, r^3] fp32
  int r, co, k_blocks, ksteps_last;   // K blocks of 64 channels per tap; UMMAs (1..4) in the last block
  long long rows;          // b * (r+2)^3
  int w_rows_bytes;        // bytes of a weight block actually needed (co rounded up to 8 rows x 128 B)
  // channels-last output (out_mode 1: bf16, 2: fp32; 0: the fp32 [b, co, r^3] layout above)
  int out_mode, out_stride;   // elements per output row (multiple of 16, >= co; channels >= co are written as zero)
  void* y_cl;                 // [rows][out_stride]; halo rows are never written (they must be zero on entry)
  double* stats;              // [b][8][2]: sum, sum of squares per GroupNorm group (8 groups), accumulated
  int batch;
  int n_acc;                  // accumulators per tile, 64 TMEM columns apart, summed by the epilogue (1, or 3: one per dz tap)
  // ceil(2^32 / (r+2)^2), ceil(2^32 / (r+2)): __umulhi(n, magic) == n / d exactly while n * d < 2^32 (n < (r+2)^3 <= 2^18)
  unsigned magic_rp2, magic_rp;
};
static void conv3d_set_magic(Conv3dTcParams& p) {
  const unsigned long long rp = (unsigned long long)p.r + 2;
  p.magic_rp2 = (unsigned)(((1ull << 32) + rp * rp - 1) / (rp * rp));
  p.magic_rp = (unsigned)(((1ull << 32) + rp - 1) / rp);
}

// L2 prefetch of a 2-D box (no shared memory involved): hides the HBM leg of a later tma_load_2d
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Half an output row (channels [H * CO / 2, (H + 1) * CO / 2) of one padded voxel) of the channels-last epilogue:
// accumulator + bias -> bf16 / fp32 row piece (the upper half also writes the padding channels up to out_stride as zero)
// and the GroupNorm(8) partial sums of those channels into st[0..7] (sums) / st[8..15] (sums of squares).  Everything is
// indexed at compile time: no running counters, no branches, the statistics never leave the registers.
template <int CO, int H>
__device__ __forceinline__ void conv3d_row_epilogue(const Conv3dTcParams& p, uint32_t taddr, const float* s_bias, bool interior,
                                                    uint8_t* yrow, float (&st)[16]) {
  constexpr int CPG = CO / 8, C_LO = H * (CO / 2);
#pragma unroll
  for (int cc = 0; cc < CO / 2; cc += 8) {
    const int c0 = C_LO + cc;
    uint32_t u[8], u1[8], u2[8];
    tmem_ld8(taddr + c0, u);
    if (p.n_acc == 3) { tmem_ld8(taddr + 64 + c0, u1); tmem_ld8(taddr + 128 + c0, u2); }
    tmem_ld_wait();
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t t = u[j];
      asm volatile("" : "+r"(t));
      float a = __uint_as_float(t);
      if (p.n_acc == 3) {
        uint32_t t1 = u1[j], t2 = u2[j];
        asm volatile("" : "+r"(t1), "+r"(t2));
        a += __uint_as_float(t1) + __uint_as_float(t2);
      }
      v[j] = interior ? a + s_bias[c0 + j] : 0.f;
      st[(C_LO + cc + j) / CPG] += v[j];
      st[8 + (C_LO + cc + j) / CPG] = fmaf(v[j], v[j], st[8 + (C_LO + cc + j) / CPG]);
    }
    if (interior) {
      if (p.out_mode == 1) {
        *reinterpret_cast<uint4*>(yrow + c0 * 2) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                                              pack_bf16(v[6], v[7]));
      } else {
        float4* dst = reinterpret_cast<float4*>(yrow + c0 * 4);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  if (H == 1 && interior) {           // padding channels of the row (a bf16 row is padded to a multiple of 64 channels)
    for (int c0 = CO; c0 < p.out_stride; c0 += 8) {
      if (p.out_mode == 1) {
        *reinterpret_cast<uint4*>(yrow + c0 * 2) = make_uint4(0, 0, 0, 0);
      } else {
        float4* dst = reinterpret_cast<float4*>(yrow + c0 * 4);
        dst[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// epilogue warps (2..9) of every kernel variant: thread <-> padded voxel row; two warps share a TMEM lane quarter and take
// half of the channels each (the epilogue, not the tensor pipe, bounds these layers: N = 48 / 96 columns per 128 rows);
// halo voxels and rows past the end are dropped
__device__ __forceinline__ void conv3d_epilogue(const Conv3dTcParams& p, uint8_t* smem, uint32_t tmem, uint64_t* acc_full,
                                                long long row0, int tid, int lane, int wid, const float* s_bias,
                                                uint32_t parity = 0, long long tile = -1) {
  // s_bias: the layer's bias staged in shared memory by the CTA (zeros when the layer has none): a global load per
  // element here cost ~1 us per 16-channel batch behind the TMA traffic - 60 % of the epilogue, which bounded the kernel
  if (tile < 0) tile = blockIdx.x;
  const int rp = p.r + 2, rp2 = rp * rp;
  {
    const int ew = wid - 2, q = wid & 3, half = ew >> 2;      // warps 2..5: lower channel half, 6..9: upper (wid & 3 = TMEM quarter)
    const long long m = row0 + q * 32 + lane;
    const int P = rp2 * rp;
    // One division per tile (32-bit whenever the grid has fewer than 2^31 rows), voxel coordinates by multiplication with
    // the precomputed reciprocals: the 128 rows of a tile touch at most two clouds (P >= 216).  The per-thread 64-bit
    // division and three 32-bit ones this replaces were a fifth of the stall samples of the first layer's epilogue, which
    // is what bounds that layer.
    const long long b0 = (p.rows < 0x7fffffffLL) ? (long long)((unsigned)row0 / (unsigned)P) : row0 / P;
    int pp = (int)(row0 - b0 * P) + q * 32 + lane;
    const long long b = b0 + (pp >= P ? 1 : 0);
    pp -= pp >= P ? P : 0;
    const int x = (int)__umulhi((unsigned)pp, p.magic_rp2), rem = pp - x * rp2;
    const int yy = (int)__umulhi((unsigned)rem, p.magic_rp), z = rem - yy * rp;
    const bool interior = m < p.rows && x >= 1 && x <= p.r && yy >= 1 && yy <= p.r && z >= 1 && z <= p.r;
    const int r3 = p.r * p.r * p.r;
    const int v = ((x - 1) * p.r + (yy - 1)) * p.r + (z - 1);
    float* yb = p.y + ((size_t)b * p.co) * r3 + v;
    mbar_wait(acc_full, parity);
    tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
    if (p.out_mode == 0) {
      if (half == 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < p.co; c0 += 16) {
          uint32_t u[16];
          tmem_ld16(taddr + c0, u);
          tmem_ld_wait();
          if (interior) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.co) yb[(size_t)(c0 + j) * r3] = __uint_as_float(u[j]) + s_bias[c0 + j];
          }
        }
      }
    } else {
      // channels-last row of this voxel + GroupNorm partial sums
      float st[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) st[i] = 0.f;
      uint8_t* yrow = reinterpret_cast<uint8_t*>(p.y_cl) + (size_t)m * p.out_stride * (p.out_mode == 1 ? 2 : 4);
      if (p.co == 48) {
        if (half == 0) conv3d_row_epilogue<48, 0>(p, taddr, s_bias, interior, yrow, st);
        else conv3d_row_epilogue<48, 1>(p, taddr, s_bias, interior, yrow, st);
      } else if (p.co == 96) {
        if (half == 0) conv3d_row_epilogue<96, 0>(p, taddr, s_bias, interior, yrow, st);
        else conv3d_row_epilogue<96, 1>(p, taddr, s_bias, interior, yrow, st);
      } else if (half == 0) {
        // other widths: one warp per quarter walks the whole row with a running group counter
        float* part = reinterpret_cast<float*>(smem) + 256 + (ew * 32 + lane) * 17;
        const int cpg = p.co >> 3, n_umma = (p.co + 15) & ~15;
        float gs = 0.f, gq = 0.f;
        int g = 0, in_g = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < p.out_stride; c0 += 16) {
          float vv[16];
          if (c0 < n_umma) {
            uint32_t u[16];
            tmem_ld16(taddr + c0, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              vv[j] = (interior && c0 + j < p.co) ? __uint_as_float(u[j]) + s_bias[c0 + j] : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) vv[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c0 + j < p.co) {
              gs += vv[j];
              gq = fmaf(vv[j], vv[j], gq);
              if (++in_g == cpg) { part[g] = gs; part[8 + g] = gq; ++g; in_g = 0; gs = 0.f; gq = 0.f; }
            }
          }
          if (interior) {
            if (p.out_mode == 1) {
              uint4* dst = reinterpret_cast<uint4*>(yrow + c0 * 2);
              dst[0] = make_uint4(pack_bf16(vv[0], vv[1]), pack_bf16(vv[2], vv[3]), pack_bf16(vv[4], vv[5]), pack_bf16(vv[6], vv[7]));
              dst[1] = make_uint4(pack_bf16(vv[8], vv[9]), pack_bf16(vv[10], vv[11]), pack_bf16(vv[12], vv[13]), pack_bf16(vv[14], vv[15]));
            } else {
              float4* dst = reinterpret_cast<float4*>(yrow + c0 * 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) dst[k] = make_float4(vv[4 * k], vv[4 * k + 1], vv[4 * k + 2], vv[4 * k + 3]);
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = part[i];
      }
      // warp totals of the 16 statistics (32 voxels) per cloud slot (the 128 rows of a CTA touch at most two clouds:
      // slot 0 = the cloud of its first row, slot 1 = the next one), then a fixed-order sum over the 8 epilogue warps:
      // the statistics are bit-reproducible (no atomics); conv_stats_finalize_kernel adds the tiles of a cloud in a fixed order
      float* wtot = reinterpret_cast<float*>(smem);                     // [8 warps][2 slots][16]
      for (int slot = 0; slot < 2; ++slot) {
        // (all but one tile in ~137 lie inside one cloud: the second reduce-scatter is skipped for them, warp-uniformly)
        if (slot == 1 && !__any_sync(0xffffffffu, interior && b != b0)) {
          if ((lane & 1) == 0) wtot[(ew * 2 + 1) * 16 + (lane >> 1)] = 0.f;
          break;
        }
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = (interior && b == b0 + slot) ? st[i] : 0.f;
        rs_step<16, 8>(a, lane); rs_step<8, 4>(a, lane); rs_step<4, 2>(a, lane); rs_step<2, 1>(a, lane);
        a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
        if ((lane & 1) == 0) wtot[(ew * 2 + slot) * 16 + (lane >> 1)] = a[0];   // index 0..7 sums, 8..15 sums of squares
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ew == 0) {
        const int slot = lane >> 4, idx = lane & 15;
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += (double)wtot[(w * 2 + slot) * 16 + idx];
        p.stats[((size_t)tile * 2 + slot) * 16 + idx] = t;                // p.stats = per-tile partials here
      }
    }
    tc_fence_before();
  }
}

template <int WSLOT>
__global__ void __launch_bounds__(c3::NTHREADS, (WSLOT <= 12288) ? 2 : 1) conv3d_tc_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                    const __grid_constant__ Conv3dTcParams p) {
  using namespace c3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // pointer arithmetic (no integer round trip) keeps the shared address space visible to the compiler: LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int STAGE_BYTES = A_BYTES + WSLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* s_bias = reinterpret_cast<float*>(bars) + 32;        // [128], 128 bytes into the barrier block
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bias[i] = (p.bias && i < p.co) ? __ldg(p.bias + i) : 0.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long row0 = (long long)blockIdx.x * 128;
  const int rp = p.r + 2, rp2 = rp * rp;
  const int n_it = 27 * p.k_blocks;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
  }
  if (wid == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
    // ---- producer: per (tap, K block) one shifted A tile (TMA tensor load) + one weight block (bulk copy)
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES, round = it / STAGES;
      const int tap = it / p.k_blocks, kb = it - tap * p.k_blocks;
      const int shift = (tap / 9 - 1) * rp2 + ((tap / 3) % 3 - 1) * rp + (tap % 3 - 1);
      if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[s], A_BYTES + p.w_rows_bytes);
        tma_load_2d(smem + s * STAGE_BYTES, &xmap, kb * 64, (int)(row0 + shift), &full[s]);
        bulk_g2s(smem + s * STAGE_BYTES + A_BYTES, p.w_img + (size_t)it * W_BYTES, p.w_rows_bytes, &full[s]);
      }
      __syncwarp();
    }
  } else if (wid == 1) {
    // ---- UMMA issuer (N = co rounded up to 16)
    const uint32_t idesc = idesc_bf16(128, (p.co + 15) & ~15);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
    const uint32_t base = smem_u32(smem);
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES;
      const int kb = it % p.k_blocks;
      mbar_wait(&full[s], (it / STAGES) & 1);
      tc_fence_after();
      const uint64_t ad = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGE_BYTES) >> 4));
      const uint64_t bd = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGE_BYTES + A_BYTES) >> 4));
      const int ks = (kb == p.k_blocks - 1) ? p.ksteps_last : 4;
      const uint32_t acc = it != 0;
      if (ks == 4) umma_bf16_block_elect<4>(tmem, ad, bd, idesc, acc);
      else if (ks == 3) { umma_bf16_block_elect<2>(tmem, ad, bd, idesc, acc); umma_bf16_block_elect<1>(tmem, ad + 4, bd + 4, idesc, 1u); }
      else if (ks == 2) umma_bf16_block_elect<2>(tmem, ad, bd, idesc, acc);
      else umma_bf16_block_elect<1>(tmem, ad, bd, idesc, acc);
      umma_commit_elect(&empty[s]);
    }
    umma_commit_elect(acc_full);
  } else {
    conv3d_epilogue(p, smem, tmem, acc_full, row0, tid, lane, wid, s_bias);
  }
  __syncthreads();
  if (wid == 1) tmem_dealloc<128>(tmem);
}

// Variant that fetches the A rows once per (dx, dy): the three dz taps of a filter column are the same rows shifted by one,
// so one 136-row TMA box serves three UMMA groups whose A descriptor starts 0 / 128 / 256 bytes into the tile (the
// 128-byte swizzle is a function of the absolute shared-memory address bits, so a descriptor start that is not a
// multiple of the 8-row swizzle atom needs nothing else - measured on the B200: with the matrix-base-offset field set to
// the row phase (bo_mode 1) the results are wrong, with the field left 0 (bo_mode 2, the default) they are exact).
// A traffic drops 3x; a stage is one A tile + three weight blocks.
namespace c3 {
constexpr int A3_ROWS = 136, A3_BYTES = 18432;      // 136 x 128 B = 17408, slot rounded to 1024
constexpr int STAGES3 = 2;
}
template <int WSLOT>
__global__ void __launch_bounds__(c3::NTHREADS, 2) conv3d_tc3_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                     const __grid_constant__ Conv3dTcParams p, int bo_mode) {
  using namespace c3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int STAGE_BYTES = A3_BYTES + 3 * WSLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES3 * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES3;
  uint64_t* acc_full = bars + 2 * STAGES3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES3 + 1);
  float* s_bias = reinterpret_cast<float*>(bars) + 32;        // [128], 128 bytes into the barrier block
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bias[i] = (p.bias && i < p.co) ? __ldg(p.bias + i) : 0.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long row0 = (long long)blockIdx.x * 128;
  const int rp = p.r + 2, rp2 = rp * rp;
  const int n_it = 9 * p.k_blocks;

  if (tid == 0) {
    for (int s = 0; s < STAGES3; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
  }
  if (wid == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES3, round = it / STAGES3;
      const int col = it / p.k_blocks, kb = it - col * p.k_blocks;       // col = dx*3 + dy
      const int shift = (col / 3 - 1) * rp2 + (col % 3 - 1) * rp - 1;     // row of the dz = -1 tap
      if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[s], A3_ROWS * 128 + 3 * p.w_rows_bytes);
        tma_load_2d(smem + s * STAGE_BYTES, &xmap, kb * 64, (int)(row0 + shift), &full[s]);
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
          bulk_g2s(smem + s * STAGE_BYTES + A3_BYTES + dz * WSLOT,
                   p.w_img + ((size_t)(col * 3 + dz) * p.k_blocks + kb) * W_BYTES, p.w_rows_bytes, &full[s]);
      }
      __syncwarp();
    }
  } else if (wid == 1) {
    const uint32_t idesc = idesc_bf16(128, (p.co + 15) & ~15);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
    const uint32_t base = smem_u32(smem);
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES3;
      const int kb = it % p.k_blocks;
      mbar_wait(&full[s], (it / STAGES3) & 1);
      tc_fence_after();
      const int ks = (kb == p.k_blocks - 1) ? p.ksteps_last : 4;
#pragma unroll 1
      for (int dz = 0; dz < 3; ++dz) {
        const uint32_t a_addr = base + s * STAGE_BYTES + dz * 128;
        const uint32_t a_hi = hi | (bo_mode == 1 ? ((uint32_t)dz << 17) : 0u);      // descriptor bits [49,52)
        const uint64_t ad = ((uint64_t)a_hi << 32) | (0x10000u | (a_addr >> 4));
        const uint64_t bd = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGE_BYTES + A3_BYTES + dz * WSLOT) >> 4));
        const uint32_t acc = (it != 0 || dz != 0) ? 1u : 0u;
        if (ks == 4) umma_bf16_block_elect<4>(tmem, ad, bd, idesc, acc);
        else if (ks == 3) { umma_bf16_block_elect<2>(tmem, ad, bd, idesc, acc); umma_bf16_block_elect<1>(tmem, ad + 4, bd + 4, idesc, 1u); }
        else if (ks == 2) umma_bf16_block_elect<2>(tmem, ad, bd, idesc, acc);
        else umma_bf16_block_elect<1>(tmem, ad, bd, idesc, acc);
      }
      umma_commit_elect(&empty[s]);
    }
    umma_commit_elect(acc_full);
  } else {
    conv3d_epilogue(p, smem, tmem, acc_full, row0, tid, lane, wid, s_bias);
  }
  __syncthreads();
  if (wid == 1) tmem_dealloc<128>(tmem);
}

// Persistent, weight-stationary variant of conv3d_tc3_kernel for layers whose whole filter bank fits beside the pipeline
// (48 -> 48: 27 x 6 KB = 162 KB): one CTA per SM loads the 27 weight blocks ONCE and walks the 128-voxel tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...  The one-tile-per-CTA kernel re-streams the bank for every tile (162 KB of
// weights against 153 KB of activations per tile, both L2 -> shared memory) and pays TMEM allocation, barrier set-up and
// pipeline fill / drain per tile; here a tile costs only its nine activation boxes, and two TMEM accumulators let the
// epilogue of tile t run under the UMMAs of tile t + 1.
namespace c3 {
constexpr int P_SCRATCH = 10240;            // epilogue partial sums [128][17] + [4][2][16] floats
}
// Activation rows are fetched once per dx PLANE: the three dy columns of a plane are the same rows shifted by r + 2, so
// one box of 130 + 2 (r + 2) rows serves nine UMMA groups whose A descriptors start (dy (r + 2) + dz) rows into it
// (SWIZZLE_128B is a function of the absolute shared-memory address: any row offset works with base-offset 0).  Three
// 23 KB loads per tile instead of nine 17 KB ones: the one-tile kernels were bound by the latency of those round trips.
__global__ void __launch_bounds__(c3::NTHREADS, 1) conv3d_tc3p_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                      const __grid_constant__ Conv3dTcParams p, int n_tiles,
                                                                      int stages, int a_rows, int a_slot) {
  using namespace c3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int n_blocks = 27 * p.k_blocks;
  uint8_t* s_w = smem;                                        // [27 * k_blocks][w_rows_bytes]
  uint8_t* s_a = s_w + (size_t)n_blocks * p.w_rows_bytes;     // [stages][a_slot]
  uint8_t* s_scr = s_a + stages * a_slot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scr + P_SCRATCH);
  uint64_t* full = bars;             // [4]
  uint64_t* empty = bars + 4;        // [4]
  uint64_t* acc_full = bars + 8;     // [2]
  uint64_t* acc_empty = bars + 10;   // [2]
  uint64_t* w_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* s_bias = reinterpret_cast<float*>(bars) + 32;        // [128], 128 bytes into the barrier block
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bias[i] = (p.bias && i < p.co) ? __ldg(p.bias + i) : 0.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int rp = p.r + 2, rp2 = rp * rp;
  const int n_it = 3 * p.k_blocks;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    mbar_init(w_full, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
  }
  if (wid == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
    // ---- producer: the filter bank once, then the activation boxes of every tile through the ring
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_full, (uint32_t)(n_blocks * p.w_rows_bytes));
      for (int k = 0; k < n_blocks; ++k)
        bulk_g2s(s_w + (size_t)k * p.w_rows_bytes, p.w_img + (size_t)k * W_BYTES, p.w_rows_bytes, w_full);
    }
    __syncwarp();
    int it = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row0 = (long long)tile * 128;
      // the boxes of this CTA's NEXT tile on their way into L2 while this one is computed: the ring (two 23 KB stages
      // beside the resident filter bank) is too shallow to cover an HBM round trip per box
      if (tile + (int)gridDim.x < n_tiles && elect_one_sync()) {
        const long long rown = (long long)(tile + gridDim.x) * 128;
        for (int i = 0; i < n_it; ++i) {
          const int dx = i / p.k_blocks, kb = i - dx * p.k_blocks;
          tma_prefetch_2d(&xmap, kb * 64, (int)(rown + (dx - 1) * rp2 - rp - 1));
        }
      }
      __syncwarp();
#pragma unroll 1
      for (int i = 0; i < n_it; ++i, ++it) {
        const int s = it % stages, round = it / stages;
        const int dx = i / p.k_blocks, kb = i - dx * p.k_blocks;
        const int shift = (dx - 1) * rp2 - rp - 1;                          // row of the (dy, dz) = (-1, -1) tap
        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[s], (uint32_t)a_rows * 128u);
          tma_load_2d(s_a + s * a_slot, &xmap, kb * 64, (int)(row0 + shift), &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (wid == 1) {
    // ---- UMMA issuer
    const uint32_t idesc = idesc_bf16(128, (p.co + 15) & ~15);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
    const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
    mbar_wait(w_full, 0);
    tc_fence_after();
    int it = 0, tl = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int buf = tl & 1;
      if (tl >= 2) { mbar_wait(&acc_empty[buf], ((tl >> 1) - 1) & 1); tc_fence_after(); }
      const uint32_t d = tmem + buf * 256;
#pragma unroll 1
      for (int i = 0; i < n_it; ++i, ++it) {
        const int s = it % stages;
        const int dx = i / p.k_blocks, kb = i - dx * p.k_blocks;
        mbar_wait(&full[s], (it / stages) & 1);
        tc_fence_after();
        const int ks = (kb == p.k_blocks - 1) ? p.ksteps_last : 4;
        // three accumulators, one per dz tap, 64 TMEM columns apart (the epilogue adds them): consecutive UMMAs never
        // accumulate into the same tile, and the three dz taps of a (dy, k) step go out behind one elect with their
        // descriptors formed by constant adds - the issuing warp, not the tensor pipe, bounded this loop
        const uint64_t a0 = ((uint64_t)hi << 32) | (0x10000u | ((a_base + s * a_slot) >> 4));
        const uint64_t b0 = ((uint64_t)hi << 32) | (0x10000u | ((w_base + (uint32_t)((dx * 9 * p.k_blocks + kb) * p.w_rows_bytes)) >> 4));
        const uint64_t b_tap = (uint64_t)((p.k_blocks * p.w_rows_bytes) >> 4), rp8 = (uint64_t)(rp * 8);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < ks)
              umma_bf16_x3_elect(d, a0 + dy * rp8 + 2 * k, b0 + 3 * dy * b_tap + 2 * k, b_tap, idesc,
                                 (dy != 0 || k != 0) ? 1u : (i != 0 ? 1u : 0u));
          }
        }
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(&acc_full[buf]);
    }
  } else {
    // ---- epilogue warps: tile t while the issuer works on tile t + 1
    int tl = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int buf = tl & 1;
      conv3d_epilogue(p, s_scr, tmem + buf * 256, &acc_full[buf], (long long)tile * 128, tid, lane, wid, s_bias,
                      (uint32_t)((tl >> 1) & 1), tile);
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      asm volatile("bar.sync 1, 256;" ::: "memory");       // the scratch sums of this tile have been consumed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 1) tmem_dealloc<512>(tmem);
}

// Persistent variant for the layers whose filter bank does NOT fit beside the pipeline (48 -> 96 and 96 -> 96 at 12^3: 324 /
// 648 KB of weight images): the one-tile-per-CTA kernel streams 36 KB of weights against 17 KB of activations per column
// step - 1.33 GB through the L2 for the 96 -> 96 layer of 64 clouds, 13.7 TB/s: the L2's limit (~12 TB/s on this part),
// not the tensor pipe's (46 % active).  Here a CTA works on TWO adjacent 128-voxel tiles per weight stage (every weight
// block fetched from L2 feeds two UMMA groups), one CTA per SM walks tile pairs blockIdx.x, blockIdx.x + gridDim.x, ...,
// and two TMEM accumulator sets let the epilogue of a pair run under the UMMAs of the next one.
namespace c3 {
constexpr int AM_BYTES = 17408;              // 136 x 128 B exactly (a multiple of 1024: SWIZZLE_128B tile base)
constexpr int STAGESM = 3;
constexpr int STAGEM_BYTES = 2 * AM_BYTES + 3 * W_SLOT;     // 71680
}
__global__ void __launch_bounds__(c3::NTHREADS, 1) conv3d_tc3m_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                      const __grid_constant__ Conv3dTcParams p, int n_tiles) {
  using namespace c3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_scr = smem + STAGESM * STAGEM_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scr + P_SCRATCH);
  uint64_t* full = bars;             // [4]
  uint64_t* empty = bars + 4;        // [4]
  uint64_t* acc_full = bars + 8;     // [2]
  uint64_t* acc_empty = bars + 10;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* s_bias = reinterpret_cast<float*>(bars) + 32;        // [128], 128 bytes into the barrier block
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bias[i] = (p.bias && i < p.co) ? __ldg(p.bias + i) : 0.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int rp = p.r + 2, rp2 = rp * rp;
  const int n_it = 9 * p.k_blocks;
  const int n_groups = (n_tiles + 1) >> 1;

  if (tid == 0) {
    for (int s = 0; s < STAGESM; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
  }
  if (wid == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
    // ---- producer: per column step the activation boxes of both tiles and the three dz weight blocks
    int it = 0;
#pragma unroll 1
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
      const long long row0 = (long long)grp * 256;
#pragma unroll 1
      for (int i = 0; i < n_it; ++i, ++it) {
        const int s = it % STAGESM, round = it / STAGESM;
        const int col = i / p.k_blocks, kb = i - col * p.k_blocks;          // col = dx*3 + dy
        const int shift = (col / 3 - 1) * rp2 + (col % 3 - 1) * rp - 1;     // row of the dz = -1 tap
        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[s], 2 * A3_ROWS * 128 + 3 * p.w_rows_bytes);
          // (rows past the end of the grid - the second tile of an odd last pair - are filled with zeros by the TMA unit)
          tma_load_2d(smem + s * STAGEM_BYTES, &xmap, kb * 64, (int)(row0 + shift), &full[s]);
          tma_load_2d(smem + s * STAGEM_BYTES + AM_BYTES, &xmap, kb * 64, (int)(row0 + 128 + shift), &full[s]);
#pragma unroll
          for (int dz = 0; dz < 3; ++dz)
            bulk_g2s(smem + s * STAGEM_BYTES + 2 * AM_BYTES + dz * W_SLOT,
                     p.w_img + ((size_t)(col * 3 + dz) * p.k_blocks + kb) * W_BYTES, p.w_rows_bytes, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (wid == 1) {
    // ---- UMMA issuer
    const uint32_t idesc = idesc_bf16(128, (p.co + 15) & ~15);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
    const uint32_t base = smem_u32(smem);
    int it = 0, gl = 0;
#pragma unroll 1
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++gl) {
      const int buf = gl & 1;
      if (gl >= 2) { mbar_wait(&acc_empty[buf], ((gl >> 1) - 1) & 1); tc_fence_after(); }
#pragma unroll 1
      for (int i = 0; i < n_it; ++i, ++it) {
        const int s = it % STAGESM;
        const int kb = i % p.k_blocks;
        mbar_wait(&full[s], (it / STAGESM) & 1);
        tc_fence_after();
        const int ks = (kb == p.k_blocks - 1) ? p.ksteps_last : 4;
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const uint32_t d = tmem + buf * 256 + t * 128;
#pragma unroll 1
          for (int dz = 0; dz < 3; ++dz) {
            const uint64_t ad = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGEM_BYTES + t * AM_BYTES + dz * 128) >> 4));
            const uint64_t bd = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGEM_BYTES + 2 * AM_BYTES + dz * W_SLOT) >> 4));
            const uint32_t acc = (i != 0 || dz != 0) ? 1u : 0u;
            if (ks == 4) umma_bf16_block_elect<4>(d, ad, bd, idesc, acc);
            else if (ks == 3) { umma_bf16_block_elect<2>(d, ad, bd, idesc, acc); umma_bf16_block_elect<1>(d, ad + 4, bd + 4, idesc, 1u); }
            else if (ks == 2) umma_bf16_block_elect<2>(d, ad, bd, idesc, acc);
            else umma_bf16_block_elect<1>(d, ad, bd, idesc, acc);
          }
        }
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(&acc_full[buf]);
    }
  } else {
    // ---- epilogue warps: the pair g while the issuer works on pair g + 1
    int gl = 0;
#pragma unroll 1
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++gl) {
      const int buf = gl & 1;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int tile = 2 * grp + t;
        if (tile < n_tiles) {
          conv3d_epilogue(p, s_scr, tmem + buf * 256 + t * 128, &acc_full[buf], (long long)tile * 128, tid, lane, wid, s_bias,
                          (uint32_t)((gl >> 1) & 1), tile);
          asm volatile("bar.sync 1, 256;" ::: "memory");       // the scratch sums of this tile have been consumed
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

// Narrow-input variant (ci <= 16: the 3 -> 48 first layer): the padded grid has 16 channels (32-byte rows, SWIZZLE_32B),
// one K = 16 UMMA per tap, A rows again fetched once per filter column.  Weight image: [27 taps][128 rows x 32 B].
namespace c3 {
constexpr int A16_BYTES = 5120;             // 136 rows x 32 B = 4352, slot rounded to 1024
constexpr int W16_SLOT = 4096, W16_BYTES = 4096;
constexpr int STAGES16 = 4;
}
__global__ void __launch_bounds__(c3::NTHREADS, 2) conv3d_tc16_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                      const __grid_constant__ Conv3dTcParams p) {
  using namespace c3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int STAGE_BYTES = A16_BYTES + 3 * W16_SLOT;
  // (the epilogue reuses the first 9.2 KB of the stage area for its partial sums: 4 stages = 68 KB, enough)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES16 * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES16;
  uint64_t* acc_full = bars + 2 * STAGES16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES16 + 1);
  float* s_bias = reinterpret_cast<float*>(bars) + 32;        // [128], 128 bytes into the barrier block
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bias[i] = (p.bias && i < p.co) ? __ldg(p.bias + i) : 0.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long row0 = (long long)blockIdx.x * 128;
  const int rp = p.r + 2, rp2 = rp * rp;
  constexpr int n_it = 9;

  if (tid == 0) {
    for (int s = 0; s < STAGES16; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
  }
  if (wid == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES16, round = it / STAGES16;
      const int shift = (it / 3 - 1) * rp2 + (it % 3 - 1) * rp - 1;       // row of the dz = -1 tap of column it = dx*3 + dy
      if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[s], A3_ROWS * 32 + 3 * p.w_rows_bytes);
        tma_load_2d(smem + s * STAGE_BYTES, &xmap, 0, (int)(row0 + shift), &full[s]);
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
          bulk_g2s(smem + s * STAGE_BYTES + A16_BYTES + dz * W16_SLOT, p.w_img + (size_t)(it * 3 + dz) * W16_BYTES,
                   p.w_rows_bytes, &full[s]);
      }
      __syncwarp();
    }
  } else if (wid == 1) {
    const uint32_t idesc = idesc_bf16(128, (p.co + 15) & ~15);
    const uint32_t hi = (256u >> 4) | (1u << 14) | ((uint32_t)SW_32 << 29);      // 8-row groups of 32-byte rows
    const uint32_t base = smem_u32(smem);
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES16;
      mbar_wait(&full[s], (it / STAGES16) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int dz = 0; dz < 3; ++dz) {
        const uint64_t ad = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGE_BYTES + dz * 32) >> 4));
        const uint64_t bd = ((uint64_t)hi << 32) | (0x10000u | ((base + s * STAGE_BYTES + A16_BYTES + dz * W16_SLOT) >> 4));
        umma_bf16_block_elect<1>(tmem, ad, bd, idesc, (it != 0 || dz != 0) ? 1u : 0u);
      }
      umma_commit_elect(&empty[s]);
    }
    umma_commit_elect(acc_full);
  } else {
    conv3d_epilogue(p, smem, tmem, acc_full, row0, tid, lane, wid, s_bias);
  }
  __syncthreads();
  if (wid == 1) tmem_dealloc<128>(tmem);
}

// Channels-last epilogue of ONE tile by a group of FOUR warps (one per TMEM lane quarter; a thread takes all channels of its
// row): two such groups work on two tiles at once.  The epilogue of the narrow first layer is a latency chain (TMEM load ->
// adds -> shuffle reduce-scatter -> barrier -> tile sums), not an instruction-count problem: with eight warps on one tile it
// took ~4.0 k cycles per tile whatever else was changed, so the way to go faster is two chains in flight.
template <int CO>
__device__ __forceinline__ void conv3d_epilogue_g4(const Conv3dTcParams& p, float* wtot, uint32_t tmem, uint64_t* acc_full,
                                                   long long row0, int lane, int q, int gw, int bar_id, const float* s_bias,
                                                   uint32_t parity, long long tile) {
  const int rp = p.r + 2, rp2 = rp * rp, P = rp2 * rp;
  const long long m = row0 + q * 32 + lane;
  const long long b0 = (p.rows < 0x7fffffffLL) ? (long long)((unsigned)row0 / (unsigned)P) : row0 / P;
  int pp = (int)(row0 - b0 * P) + q * 32 + lane;
  const long long b = b0 + (pp >= P ? 1 : 0);
  pp -= pp >= P ? P : 0;
  const int x = (int)__umulhi((unsigned)pp, p.magic_rp2), rem = pp - x * rp2;
  const int yy = (int)__umulhi((unsigned)rem, p.magic_rp), z = rem - yy * rp;
  const bool interior = m < p.rows && x >= 1 && x <= p.r && yy >= 1 && yy <= p.r && z >= 1 && z <= p.r;
  mbar_wait(acc_full, parity);
  tc_fence_after();
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
  float st[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) st[i] = 0.f;
  uint8_t* yrow = reinterpret_cast<uint8_t*>(p.y_cl) + (size_t)m * p.out_stride * (p.out_mode == 1 ? 2 : 4);
  conv3d_row_epilogue<CO, 0>(p, taddr, s_bias, interior, yrow, st);
  conv3d_row_epilogue<CO, 1>(p, taddr, s_bias, interior, yrow, st);
  // warp totals per cloud slot, then a fixed-order sum over the group's four warps (bit-reproducible, no atomics)
  for (int slot = 0; slot < 2; ++slot) {
    if (slot == 1 && !__any_sync(0xffffffffu, interior && b != b0)) {
      if ((lane & 1) == 0) wtot[(gw * 2 + 1) * 16 + (lane >> 1)] = 0.f;
      break;
    }
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (interior && b == b0 + slot) ? st[i] : 0.f;
    rs_step<16, 8>(a, lane); rs_step<8, 4>(a, lane); rs_step<4, 2>(a, lane); rs_step<2, 1>(a, lane);
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
    if ((lane & 1) == 0) wtot[(gw * 2 + slot) * 16 + (lane >> 1)] = a[0];
  }
  asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
  if (gw == 0) {
    const int slot = lane >> 4, idx = lane & 15;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) t += (double)wtot[(w * 2 + slot) * 16 + idx];
    p.stats[((size_t)tile * 2 + slot) * 16 + idx] = t;
  }
  tc_fence_before();
}

// Persistent form of the narrow first layer for large grids, built like conv3d_tc3p_kernel: the 27 K = 16 taps (27 x co x
// 32 B) stay resident, one CTA per SM walks the tiles, one activation box per dx plane (130 + 2 (r + 2) rows of 32 bytes
// serve the nine (dy, dz) taps), three accumulators per tile (one per dz) in two TMEM buffers so the epilogue of tile t runs
// under the UMMAs of tile t + 1.  The one-tile CTAs spent most of their life outside the main loop (TMEM allocation, first
// loads, epilogue): 17 % tensor pipe, 4.4 us per tile and CTA.
__global__ void __launch_bounds__(c3::NTHREADS, 1) conv3d_tc16p_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                       const __grid_constant__ Conv3dTcParams p, int n_tiles,
                                                                       int stages, int a_rows, int a_slot, int w_slot) {
  using namespace c3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;                                        // [27][w_slot]
  uint8_t* s_a = s_w + 27 * w_slot;                           // [stages][a_slot]
  uint8_t* s_scr = s_a + stages * a_slot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scr + P_SCRATCH);
  uint64_t* full = bars;             // [4]
  uint64_t* empty = bars + 4;        // [4]
  uint64_t* acc_full = bars + 8;     // [2]
  uint64_t* acc_empty = bars + 10;   // [2]
  uint64_t* w_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* s_bias = reinterpret_cast<float*>(bars) + 32;        // [128], 128 bytes into the barrier block
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bias[i] = (p.bias && i < p.co) ? __ldg(p.bias + i) : 0.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int rp = p.r + 2, rp2 = rp * rp;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    mbar_init(w_full, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
  }
  if (wid == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
    // ---- producer: the filter bank once, then three plane boxes per tile
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_full, (uint32_t)(27 * p.w_rows_bytes));
      for (int k = 0; k < 27; ++k) bulk_g2s(s_w + (size_t)k * w_slot, p.w_img + (size_t)k * W16_BYTES, p.w_rows_bytes, w_full);
    }
    __syncwarp();
    int it = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row0 = (long long)tile * 128;
      // (no L2 prefetch of the next tile here: with 32-byte rows the TMA unit's row rate, not the HBM leg, is the limit)
#pragma unroll 1
      for (int dx = 0; dx < 3; ++dx, ++it) {
        const int s = it % stages, round = it / stages;
        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[s], (uint32_t)a_rows * 32u);
          tma_load_2d(s_a + s * a_slot, &xmap, 0, (int)(row0 + (dx - 1) * rp2 - rp - 1), &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (wid == 1) {
    // ---- UMMA issuer: nine K = 16 taps per plane, three per elect (the dz taps: operand row + 1, accumulator + 64 columns)
    const uint32_t idesc = idesc_bf16(128, (p.co + 15) & ~15);
    const uint32_t hi = (256u >> 4) | (1u << 14) | ((uint32_t)SW_32 << 29);      // 8-row groups of 32-byte rows
    const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
    const uint64_t b_tap = (uint64_t)(w_slot >> 4), rp2u = (uint64_t)(rp * 2);
    mbar_wait(w_full, 0);
    tc_fence_after();
    int it = 0, tl = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int buf = tl & 1;
      if (tl >= 2) { mbar_wait(&acc_empty[buf], ((tl >> 1) - 1) & 1); tc_fence_after(); }
      const uint32_t d = tmem + buf * 256;
#pragma unroll 1
      for (int dx = 0; dx < 3; ++dx, ++it) {
        const int s = it % stages;
        mbar_wait(&full[s], (it / stages) & 1);
        tc_fence_after();
        const uint64_t a0 = ((uint64_t)hi << 32) | (0x10000u | ((a_base + s * a_slot) >> 4));
        const uint64_t b0 = ((uint64_t)hi << 32) | (0x10000u | ((w_base + (uint32_t)(dx * 9 * w_slot)) >> 4));
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
          umma_bf16_x3_elect_a<2>(d, a0 + dy * rp2u, b0 + 3 * dy * b_tap, b_tap, idesc, (dx != 0 || dy != 0) ? 1u : 0u);
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(&acc_full[buf]);
    }
  } else {
    // ---- epilogue warps: tile t while the issuer works on tile t + 1
    int tl = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int buf = tl & 1;
      conv3d_epilogue(p, s_scr, tmem + buf * 256, &acc_full[buf], (long long)tile * 128, tid, lane, wid, s_bias,
                      (uint32_t)((tl >> 1) & 1), tile);
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      asm volatile("bar.sync 1, 256;" ::: "memory");       // the scratch sums of this tile have been consumed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 1) tmem_dealloc<512>(tmem);
}

// conv3d_tc16p_kernel with TWO tiles' epilogues in flight: four single-accumulator TMEM buffers (64 columns each, all 27 taps
// of a tile accumulate into one), the eight epilogue warps as two groups of four (conv3d_epilogue_g4): group g takes the
// CTA's tiles g, g + 2, ... while the issuer runs up to four tiles ahead.
__global__ void __launch_bounds__(c3::NTHREADS, 1) conv3d_tc16g_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                       const __grid_constant__ Conv3dTcParams p, int n_tiles,
                                                                       int stages, int a_rows, int a_slot, int w_slot) {
  using namespace c3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;                                        // [27][w_slot]
  uint8_t* s_a = s_w + 27 * w_slot;                           // [stages][a_slot]
  uint8_t* s_scr = s_a + stages * a_slot;                     // [2 groups][4 warps][2 slots][16] floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scr + P_SCRATCH);
  uint64_t* full = bars;             // [4]
  uint64_t* empty = bars + 4;        // [4]
  uint64_t* acc_full = bars + 8;     // [4]
  uint64_t* acc_empty = bars + 12;   // [4]
  uint64_t* w_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  float* s_bias = reinterpret_cast<float*>(bars) + 64;        // [128], 256 bytes into the barrier block
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bias[i] = (p.bias && i < p.co) ? __ldg(p.bias + i) : 0.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int rp = p.r + 2, rp2 = rp * rp;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    mbar_init(w_full, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
  }
  if (wid == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid == 0) {
    // ---- producer: the filter bank once, then three plane boxes per tile
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_full, (uint32_t)(27 * p.w_rows_bytes));
      for (int k = 0; k < 27; ++k) bulk_g2s(s_w + (size_t)k * w_slot, p.w_img + (size_t)k * W16_BYTES, p.w_rows_bytes, w_full);
    }
    __syncwarp();
    int it = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row0 = (long long)tile * 128;
#pragma unroll 1
      for (int dx = 0; dx < 3; ++dx, ++it) {
        const int s = it % stages, round = it / stages;
        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[s], (uint32_t)a_rows * 32u);
          tma_load_2d(s_a + s * a_slot, &xmap, 0, (int)(row0 + (dx - 1) * rp2 - rp - 1), &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (wid == 1) {
    // ---- UMMA issuer: 27 K = 16 taps of a tile into one accumulator, three (the dz taps) per elect
    const uint32_t idesc = idesc_bf16(128, (p.co + 15) & ~15);
    const uint32_t hi = (256u >> 4) | (1u << 14) | ((uint32_t)SW_32 << 29);      // 8-row groups of 32-byte rows
    const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
    const uint64_t b_tap = (uint64_t)(w_slot >> 4), rp2u = (uint64_t)(rp * 2);
    mbar_wait(w_full, 0);
    tc_fence_after();
    int it = 0, tl = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int buf = tl & 3;
      if (tl >= 4) { mbar_wait(&acc_empty[buf], ((tl >> 2) - 1) & 1); tc_fence_after(); }
      const uint32_t d = tmem + buf * 64;
#pragma unroll 1
      for (int dx = 0; dx < 3; ++dx, ++it) {
        const int s = it % stages;
        mbar_wait(&full[s], (it / stages) & 1);
        tc_fence_after();
        const uint64_t a0 = ((uint64_t)hi << 32) | (0x10000u | ((a_base + s * a_slot) >> 4));
        const uint64_t b0 = ((uint64_t)hi << 32) | (0x10000u | ((w_base + (uint32_t)(dx * 9 * w_slot)) >> 4));
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
          umma_bf16_x3_same_elect_a<2>(d, a0 + dy * rp2u, b0 + 3 * dy * b_tap, b_tap, idesc, (dx != 0 || dy != 0) ? 1u : 0u);
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(&acc_full[buf]);
    }
  } else {
    // ---- epilogue: group g (warps 2 + 4 g .. 5 + 4 g, one per TMEM lane quarter) takes the CTA's tiles g, g + 2, ...
    const int ew = wid - 2, g = ew >> 2, q = wid & 3, gw = ew & 3;
    float* wtot = reinterpret_cast<float*>(s_scr) + g * 256;
    int tl = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      if ((tl & 1) != g) continue;
      const int buf = tl & 3;
      const uint32_t par = (uint32_t)((tl >> 2) & 1);
      if (p.co == 48)
        conv3d_epilogue_g4<48>(p, wtot, tmem + buf * 64, &acc_full[buf], (long long)tile * 128, lane, q, gw, 1 + g, s_bias, par, tile);
      else
        conv3d_epilogue_g4<32>(p, wtot, tmem + buf * 64, &acc_full[buf], (long long)tile * 128, lane, q, gw, 1 + g, s_bias, par, tile);
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");       // the group's scratch sums have been consumed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 1) tmem_dealloc<256>(tmem);
}

// fp32 [b, c <= 16, r^3] -> bf16 zero-padded grid with 16 channels per row; one thread per padded voxel
__global__ void __launch_bounds__(256) cl_pad16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int c, int r,
                                                       long long rows) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  const int rp = r + 2, rp2 = rp * rp, P = rp2 * rp, r3 = r * r * r;
  const long long b = m / P;
  const int pp = (int)(m - b * P);
  const int xx = pp / rp2, yy = (pp / rp) % rp, zz = pp % rp;
  const bool interior = xx >= 1 && xx <= r && yy >= 1 && yy <= r && zz >= 1 && zz <= r;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.f;
  if (interior) {
    const float* xb = x + ((size_t)b * c) * r3 + ((xx - 1) * r + (yy - 1)) * r + (zz - 1);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < c) v[j] = __ldg(xb + (size_t)j * r3);
  }
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)m * 16);
  dst[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  dst[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}

// Conv3d weight fp32 [co][ci <= 16][27] -> [27 taps][128 rows x 32 B] (SWIZZLE_32B rows = output channels)
__global__ void __launch_bounds__(128) conv3d_weight_image16_kernel(const float* __restrict__ w, uint8_t* __restrict__ img, int co,
                                                                    int ci) {
  const int row = threadIdx.x, tap = blockIdx.x;
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = (row < co && k < ci) ? w[((size_t)row * ci + k) * 27 + tap] : 0.f;
#pragma unroll
  for (int chunk = 0; chunk < 2; ++chunk)
    *reinterpret_cast<uint4*>(img + (size_t)tap * c3::W16_BYTES + swz_off<32>(row, chunk)) =
        make_uint4(pack_bf16(v[8 * chunk], v[8 * chunk + 1]), pack_bf16(v[8 * chunk + 2], v[8 * chunk + 3]),
                   pack_bf16(v[8 * chunk + 4], v[8 * chunk + 5]), pack_bf16(v[8 * chunk + 6], v[8 * chunk + 7]));
}

// fp32 [b, c, r^3] -> bf16 channels-last zero-padded grid [b * (r+2)^3][cpad]; one thread per (padded voxel, 8 channels)
__global__ void __launch_bounds__(256) cl_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int c, int cpad,
                                                     int r, long long rows) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  if (m >= rows) return;
  const int rp = r + 2, rp2 = rp * rp, P = rp2 * rp, r3 = r * r * r;
  const long long b = m / P;
  const int pp = (int)(m - b * P);
  const int xx = pp / rp2, yy = (pp / rp) % rp, zz = pp % rp;
  const bool interior = xx >= 1 && xx <= r && yy >= 1 && yy <= r && zz >= 1 && zz <= r;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (interior) {
    const float* xb = x + ((size_t)b * c) * r3 + ((xx - 1) * r + (yy - 1)) * r + (zz - 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = chunk * 8 + j;
      if (ch < c) v[j] = __ldg(xb + (size_t)ch * r3);
    }
  }
  *reinterpret_cast<uint4*>(out + (size_t)m * cpad + chunk * 8) =
      make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// stats[b][g][moment] = sum over the CTAs covering cloud b of their slot partials.  One warp per statistic: lane l adds
// the CTAs l, l + 32, ... in order, then a fixed xor tree - the same summation order on every run (bit-reproducible), and
// 32 loads in flight per statistic instead of a serial walk over the ~140 CTAs of a 24^3 cloud.
__global__ void __launch_bounds__(512) conv_stats_finalize_kernel(const double* __restrict__ part, int P, long long rows,
                                                                  double* __restrict__ stats) {
  const int b = blockIdx.x, idx = threadIdx.x >> 5, lane = threadIdx.x & 31;       // idx 0..7 sums, 8..15 sums of squares
  const long long r_lo = (long long)b * P, r_hi = r_lo + P - 1;
  const long long c_lo = r_lo / 128, c_hi = r_hi / 128;
  double t = 0.0;
  for (long long cc = c_lo + lane; cc <= c_hi; cc += 32)
    t += part[((size_t)cc * 2 + (b - (int)((cc * 128) / P))) * 16 + idx];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  (void)rows;
  if (lane == 0) stats[((size_t)b * 8 + (idx & 7)) * 2 + (idx >> 3)] = t;
}
// the same for per-block partials laid out [b][nblk][16] (SIMT first conv) and [b][nblk][c] (SE squeeze sums)
__global__ void block_partials_finalize_kernel(const double* __restrict__ part, int nblk, int width, int remap,
                                               double* __restrict__ out) {
  const int b = blockIdx.x, i = threadIdx.x;
  if (i >= width) return;
  double t = 0.0;
  for (int k = 0; k < nblk; k += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = k + u < nblk ? part[((size_t)b * nblk + k + u) * width + i] : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) t += v[u];
  }
  if (remap) out[((size_t)b * 8 + (i & 7)) * 2 + (i >> 3)] = t;      // [16] -> stats[b][group][moment]
  else out[(size_t)b * width + i] = t;
}

// GroupNorm(8) + Swish in place on the interior rows of a channels-last padded grid, from the statistics the Conv3d
// epilogue accumulated (R/../pvcnn/modules/pvconv.py:52-57); optionally the per-channel sums of the result for the
// SE squeeze (se.py:18-19).  Thread <-> (row lane, 8 channels); a block covers `rpb` padded voxels of one cloud.
template <bool F32>
__global__ void __launch_bounds__(256, F32 ? 3 : 4) gn_swish_cl_kernel(void* __restrict__ y, const double* __restrict__ stats,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          int c, int stride, int r, float eps, double* __restrict__ se_sum,
                                                          int rpb) {
  __shared__ float s_red[2048];
  const int b = blockIdx.y, tid = threadIdx.x;
  // only the 8-channel chunks that hold channels get a thread (the bf16 row stride is padded to 64 / 128 channels)
  const int nchunk = (c + 7) >> 3, rl = 256 / nchunk, cw = nchunk * 8;
  const int chunk = tid % nchunk, rsub = tid / nchunk;
  const int rp = r + 2, rp2 = rp * rp, P = rp2 * rp, cpg = c >> 3;
  const float inv_rp2 = 1.0f / (float)rp2, inv_rp = 1.0f / (float)rp;
  const double cnt = (double)cpg * r * r * r;
  // y = x * A[ch] + B[ch]: one thread per channel does the double-precision mean / variance / rsqrt of its group and the
  // fold with the affine; every thread then picks up its eight channels from shared memory.  (Each thread used to derive
  // its eight pairs itself: two double divisions and a double square root per channel were 40 % of the instructions the
  // whole pass executed.)
  __shared__ __align__(16) float s_ab[2][128];
  if (tid < 128) {
    float a = 0.f, bb = 0.f;
    if (tid < c) {
      const double* st = stats + ((size_t)b * 8 + tid / cpg) * 2;
      const double mean = st[0] / cnt;
      const double var = fmax(st[1] / cnt - mean * mean, 0.0);
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      a = __ldg(gamma + tid) * rstd;
      bb = __ldg(beta + tid) - (float)mean * a;
    }
    s_ab[0][tid] = a; s_ab[1][tid] = bb;
  }
  __syncthreads();
  float A[8], B[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    A[j] = s_ab[0][chunk * 8 + j]; B[j] = s_ab[1][chunk * 8 + j]; acc[j] = 0.f;
  }
  const bool active = rsub < rl;        // 256 is not a multiple of every chunk count (6, 12 chunks for fp32 rows)
  const int p_end = active ? min(P, (int)(blockIdx.x + 1) * rpb) : 0;
  // four rows per thread and iteration: all loads are issued before the first value is used (one 16-byte load in flight per
  // thread left the pass latency bound at 40 % of the HBM rate); the rows are finished in ascending order, so the SE sums
  // accumulate exactly as before
  constexpr int U = 4;
  for (int pp0 = blockIdx.x * rpb + rsub; pp0 < p_end; pp0 += U * rl) {
    bool ok[U];
    uint4 raw[U][F32 ? 2 : 1];
    uint4* ptr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = pp0 + u * rl;
      // exact for these sizes: (pp + 0.5) / rp2 is at least 0.5 / rp2 away from an integer, far more than the fp32 rounding
      const int x = (int)(((float)pp + 0.5f) * inv_rp2), rem = pp - x * rp2;
      const int yy = (int)(((float)rem + 0.5f) * inv_rp), z = rem - yy * rp;
      ok[u] = pp < p_end && !(x < 1 || x > r || yy < 1 || yy > r || z < 1 || z > r);
      ptr[u] = F32 ? reinterpret_cast<uint4*>(reinterpret_cast<float*>(y) + ((size_t)b * P + pp) * stride + chunk * 8)
                   : reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + ((size_t)b * P + pp) * stride + chunk * 8);
      if (ok[u]) {
        raw[u][0] = ptr[u][0];
        if (F32) raw[u][F32 ? 1 : 0] = ptr[u][1];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      float v[8];
      if (F32) {
        const uint4 lo = raw[u][0], hi = raw[u][F32 ? 1 : 0];
        v[0] = __uint_as_float(lo.x); v[1] = __uint_as_float(lo.y); v[2] = __uint_as_float(lo.z); v[3] = __uint_as_float(lo.w);
        v[4] = __uint_as_float(hi.x); v[5] = __uint_as_float(hi.y); v[6] = __uint_as_float(hi.z); v[7] = __uint_as_float(hi.w);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(v[j], A[j], B[j]);
          v[j] = __fdividef(t, 1.0f + __expf(-t));          // fast-math exp / divide: ~1e-6 relative, far below the bf16 operands
          acc[j] += v[j];
        }
        ptr[u][0] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
        ptr[u][1] = make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
      } else {
        const uint32_t w[4] = {raw[u][0].x, raw[u][0].y, raw[u][0].z, raw[u][0].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
          v[2 * k] = __low2float(h);
          v[2 * k + 1] = __high2float(h);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(v[j], A[j], B[j]), h = 0.5f * t;
          float th;                                          // x * sigmoid(x) = h (1 + tanh h): one MUFU op, bf16 output
          asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
          v[j] = fmaf(h, th, h);
          acc[j] += v[j];
        }
        *ptr[u] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
    }
  }
  if (se_sum) {
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[rsub * cw + chunk * 8 + j] = acc[j];
    }
    __syncthreads();
    if (tid < c) {
      float t = 0.f;
      for (int k = 0; k < rl; ++k) t += s_red[k * cw + tid];
      se_sum[((size_t)b * gridDim.x + blockIdx.x) * c + tid] = (double)t;     // per-block partial (summed in order later)
    }
  }
}

// SE excite from channel sums: gate = sigmoid(W2 swish(W1 (sum * inv_count)))   (se.py:10-21); one block per cloud
__global__ void __launch_bounds__(128) se_gate_sum_kernel(const double* __restrict__ sum, float inv_count,
                                                          const float* __restrict__ w1, const float* __restrict__ w2, int c,
                                                          int cr, float* __restrict__ gate) {
  extern __shared__ float s_se[];   // [c] means, [cr] hidden
  float* s_m = s_se;
  float* s_h = s_se + c;
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < c; k += blockDim.x) s_m[k] = (float)(sum[(size_t)b * c + k] * (double)inv_count);
  __syncthreads();
  for (int j = threadIdx.x; j < cr; j += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < c; ++k) a = fmaf(w1[j * c + k], s_m[k], a);
    s_h[j] = a * (1.0f / (1.0f + expf(-a)));
  }
  __syncthreads();
  for (int o = threadIdx.x; o < c; o += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < cr; ++k) a = fmaf(w2[o * cr + k], s_h[k], a);
    gate[(size_t)b * c + o] = 1.0f / (1.0f + expf(-a));
  }
}

// trilinear devoxelize of (grid * gate) + point branch from a channels-last padded grid
// (R/../pvcnn/modules/functional/src/trilinear_devox/trilinear_devox.cu:20-84 for the corner weights / indices).
// Thread <-> (point, 8 channels): each corner is one 16- or 32-byte row segment.
template <bool F32>
__global__ void __launch_bounds__(128) devox_cl_kernel(const float* __restrict__ coords, const void* __restrict__ grid,
                                                       const float* __restrict__ gate, const float* __restrict__ point,
                                                       int c, int stride, int n, int r, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rp = r + 2, rp2 = rp * rp, P = rp2 * rp;
  const float* cb = coords + (size_t)b * 3 * n;
  const float x = cb[i], y = cb[i + n], z = cb[i + 2 * n];
  const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  const float xd1 = x - xl, yd1 = y - yl, zd1 = z - zl;
  const float xd0 = 1.0f - xd1, yd0 = 1.0f - yd1, zd0 = 1.0f - zd1;
  float wgt[8];
  wgt[0] = __fmul_rn(__fmul_rn(xd0, yd0), zd0); wgt[1] = __fmul_rn(__fmul_rn(xd0, yd0), zd1);
  wgt[2] = __fmul_rn(__fmul_rn(xd0, yd1), zd0); wgt[3] = __fmul_rn(__fmul_rn(xd0, yd1), zd1);
  wgt[4] = __fmul_rn(__fmul_rn(xd1, yd0), zd0); wgt[5] = __fmul_rn(__fmul_rn(xd1, yd0), zd1);
  wgt[6] = __fmul_rn(__fmul_rn(xd1, yd1), zd0); wgt[7] = __fmul_rn(__fmul_rn(xd1, yd1), zd1);
  const int xh = xd1 > 0 ? rp2 : 0, yh = yd1 > 0 ? rp : 0, zh = zd1 > 0 ? 1 : 0;
  int id[8];
  id[0] = ((int)xl + 1) * rp2 + ((int)yl + 1) * rp + ((int)zl + 1);
  id[1] = id[0] + zh; id[2] = id[0] + yh; id[3] = id[2] + zh;
  id[4] = id[0] + xh; id[5] = id[4] + zh; id[6] = id[4] + yh; id[7] = id[6] + zh;
  const int c0 = blockIdx.y * 8;
  float g[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = (gate && c0 + j < c) ? __ldg(gate + (size_t)b * c + c0 + j) : 1.f;
  const int order[8] = {1, 0, 2, 3, 4, 5, 6, 7};     // accumulation order of the fp32 kernel (devox_gate_add_kernel)
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const int k = order[kk];
    float f[8];
    if (F32) {
      const float4* ptr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grid) + ((size_t)b * P + id[k]) * stride + c0);
      const float4 lo = __ldg(ptr), hi = __ldg(ptr + 1);
      f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = lo.w; f[4] = hi.x; f[5] = hi.y; f[6] = hi.z; f[7] = hi.w;
    } else {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(grid) + ((size_t)b * P + id[k]) * stride + c0));
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
        f[2 * q] = __low2float(h);
        f[2 * q + 1] = __high2float(h);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = __fmul_rn(f[j], g[j]);
      acc[j] = kk == 0 ? __fmul_rn(wgt[k], t) : __fmaf_rn(wgt[k], t, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (c0 + j < c) {
      const size_t o = ((size_t)b * c + c0 + j) * n + i;
      out[o] = acc[j] + (point ? point[o] : 0.f);
    }
  }
}

// Conv3d weight fp32 [co][ci][27] -> images [27][k_blocks][128 rows x 128 B] (rows = output channels)
__global__ void __launch_bounds__(256) conv3d_weight_image_kernel(const float* __restrict__ w, uint8_t* __restrict__ img, int co,
                                                                  int ci, int k_blocks) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;   // one warp per output channel
  const int tap = blockIdx.y;
  if (row >= 128) return;
  for (int chunk = lane; chunk < k_blocks * 8; chunk += 32) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = chunk * 8 + j;
      v[j] = (row < co && k < ci) ? w[((size_t)row * ci + k) * 27 + tap] : 0.f;
    }
    uint8_t* dst = img + ((size_t)tap * k_blocks + (chunk >> 3)) * c3::A_BYTES + swz_off<128>(row, chunk & 7);
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                                pack_bf16(v[6], v[7]));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace gldm

using namespace gldm;

extern "C" long long gldm_conv3d_tc_weight_bytes(int ci) { return ci > 0 ? 27LL * ((ci + 63) / 64) * c3::A_BYTES : -1; }
extern "C" long long gldm_conv3d_tc_grid_bytes(int b, int ci, int r) {
  if (b < 0 || ci <= 0 || r <= 0) return -1;
  return (long long)b * (r + 2) * (r + 2) * (r + 2) * (((ci + 63) / 64) * 64) * 2;
}

extern "C" int gldm_conv3d_tc_pack_weight(const float* w, int co, int ci, void* img, void* stream) {
  GLDM_REQUIRE(w && img, "conv3d_tc_pack_weight: null pointer");
  GLDM_REQUIRE(co > 0 && co <= 128 && ci > 0, "conv3d_tc_pack_weight: co <= 128");
  conv3d_weight_image_kernel<<<dim3(16, 27), 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<uint8_t*>(img), co, ci,
                                                                             (ci + 63) / 64);
  return check_launch("conv3d_weight_image_kernel");
}

static int launch_conv3d(const void* x_cl, const void* w_img, const float* bias, int b, int ci, int co, int r, float* y,
                         int out_mode, int out_stride, void* y_cl, double* stats, cudaStream_t s) {
  const int cpad = ((ci + 63) / 64) * 64, kb = cpad / 64;
  const long long P = (long long)(r + 2) * (r + 2) * (r + 2), rows = (long long)b * P;
  // the epilogue's voxel indexing: a 128-row tile touches at most two clouds, coordinates by 32-bit reciprocal multiplication
  GLDM_REQUIRE(r >= 4 && r <= 62, "conv3d (tensor cores): resolution %d outside [4, 62]", r);
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("conv3d_k3_tc: cuTensorMapEncodeTiled is not available from the driver");
    return GLDM_ECUDA;
  }
  CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)cpad, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cpad * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_cl), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("conv3d_k3_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return GLDM_ECUDA;
  }
  Conv3dTcParams p;
  p.w_img = reinterpret_cast<const uint8_t*>(w_img);
  p.bias = bias; p.y = y; p.r = r; p.co = co; p.k_blocks = kb;
  const int last_valid = ci - (kb - 1) * 64;                   // channels in the last K block
  p.ksteps_last = (last_valid + 15) / 16;
  p.rows = rows;
  p.w_rows_bytes = ((co + 7) / 8) * 1024;
  p.out_mode = out_mode; p.out_stride = out_stride; p.y_cl = y_cl; p.stats = stats; p.batch = b; p.n_acc = 1;
  conv3d_set_magic(p);
  static SmemOptIn attr_s, attr_b;
  const int smem_small = c3::STAGES * (c3::A_BYTES + c3::W_SLOT) + 1024 + 768;
  const int smem_big = c3::STAGES * (c3::A_BYTES + c3::W_BYTES) + 1024 + 768;
  if (int rc = opt_in_smem(attr_s, conv3d_tc_kernel<c3::W_SLOT>, smem_small, "conv3d_tc_kernel")) return rc;
  if (int rc = opt_in_smem(attr_b, conv3d_tc_kernel<c3::W_BYTES>, smem_big, "conv3d_tc_kernel (wide)")) return rc;
  const unsigned grid = (unsigned)((rows + 127) / 128);
  // one A tile per filter column (conv3d_tc3_kernel) by default; GLDM_CONV3D_TAPS3=0 selects the one-tile-per-tap kernel
  static int taps3 = -1;
  if (taps3 < 0) { const char* ev = getenv("GLDM_CONV3D_TAPS3"); taps3 = ev ? atoi(ev) : 2; }
  if (taps3 > 0 && p.w_rows_bytes <= c3::W_SLOT) {
    CUtensorMap map3;
    const cuuint32_t box3[2] = {64, (cuuint32_t)c3::A3_ROWS};
    const CUresult cr3 = enc(&map3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_cl), gdim, gstride, box3, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr3 != CUDA_SUCCESS) {
      set_error("conv3d_k3_tc: cuTensorMapEncodeTiled (136-row box) failed (%d)", (int)cr3);
      return GLDM_ECUDA;
    }
    // the whole filter bank resident beside a 2- or 3-deep activation ring: persistent, weight-stationary kernel
    static int persist = -1;
    if (persist < 0) { const char* ev = getenv("GLDM_CONV3D_PERSISTENT"); persist = ev ? atoi(ev) : 1; }
    const int bank = 27 * kb * p.w_rows_bytes;
    const int fixed = c3::P_SCRATCH + 768 + 1024;
    const int a_rows = 130 + 2 * (r + 2);                               // one dx plane: dy in {-1,0,1} x dz in {-1,0,1}
    const int a_slot = ((a_rows * 128 + 1023) / 1024) * 1024;
    const int stages_p = (bank + 3 * a_slot + fixed <= 232448) ? 3 : 2;
    if (persist && out_mode != 0 && a_rows <= 256 && ((co + 15) & ~15) <= 64 && bank + stages_p * a_slot + fixed <= 232448) {
      CUtensorMap mapp;
      const cuuint32_t boxp[2] = {64, (cuuint32_t)a_rows};
      const CUresult crp = enc(&mapp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_cl), gdim, gstride, boxp, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (crp != CUDA_SUCCESS) {
        set_error("conv3d_k3_tc: cuTensorMapEncodeTiled (%d-row box) failed (%d)", a_rows, (int)crp);
        return GLDM_ECUDA;
      }
      const int smem_p = bank + stages_p * a_slot + fixed;
      static SmemOptIn attr_p;
      if (int rc = opt_in_smem(attr_p, conv3d_tc3p_kernel, 232448, "conv3d_tc3p_kernel")) return rc;
      const int n_tiles = (int)grid;
      p.n_acc = 3;
      conv3d_tc3p_kernel<<<min(n_tiles, kNumSMs), c3::NTHREADS, smem_p, s>>>(mapp, p, n_tiles, stages_p, a_rows, a_slot);
      return check_launch("conv3d_tc3p_kernel");
    }
    // filter bank too large to stay resident: two tiles per weight stage, persistent (the L2 -> shared-memory weight stream
    // is what bounds these layers)
    // GLDM_CONV3D_MULTI: 0 = never, 1 (default) = from two tiles per SM on, 2 = always (tests); read per call
    const char* evm = getenv("GLDM_CONV3D_MULTI");
    const int multi = evm ? atoi(evm) : 1;
    if (multi && out_mode != 0 && (multi == 2 || grid >= 2 * (unsigned)kNumSMs)) {
      const int smem_m = c3::STAGESM * c3::STAGEM_BYTES + c3::P_SCRATCH + 768 + 1024;
      static SmemOptIn attr_m;
      if (int rc = opt_in_smem(attr_m, conv3d_tc3m_kernel, smem_m, "conv3d_tc3m_kernel")) return rc;
      const int n_tiles = (int)grid, n_groups = (n_tiles + 1) / 2;
      p.n_acc = 1;
      conv3d_tc3m_kernel<<<min(n_groups, kNumSMs), c3::NTHREADS, smem_m, s>>>(map3, p, n_tiles);
      return check_launch("conv3d_tc3m_kernel");
    }
    const int smem3 = c3::STAGES3 * (c3::A3_BYTES + 3 * c3::W_SLOT) + 1024 + 768;
    static SmemOptIn attr3;
    if (int rc = opt_in_smem(attr3, conv3d_tc3_kernel<c3::W_SLOT>, smem3, "conv3d_tc3_kernel")) return rc;
    conv3d_tc3_kernel<c3::W_SLOT><<<grid, c3::NTHREADS, smem3, s>>>(map3, p, taps3);
    return check_launch("conv3d_tc3_kernel");
  }
  if (p.w_rows_bytes <= c3::W_SLOT) conv3d_tc_kernel<c3::W_SLOT><<<grid, c3::NTHREADS, smem_small, s>>>(map, p);
  else conv3d_tc_kernel<c3::W_BYTES><<<grid, c3::NTHREADS, smem_big, s>>>(map, p);
  return check_launch("conv3d_tc_kernel");
}

static int launch_cl_pad(const float* x, int b, int c, int r, void* out, cudaStream_t s) {
  const int cpad = ((c + 63) / 64) * 64;
  const long long P = (long long)(r + 2) * (r + 2) * (r + 2), rows = (long long)b * P;
  cl_pad_kernel<<<dim3((unsigned)((rows + 255) / 256), cpad / 8), 256, 0, s>>>(x, reinterpret_cast<__nv_bfloat16*>(out), c, cpad,
                                                                               r, rows);
  return check_launch("cl_pad_kernel");
}

/* x f32[b,ci,r^3] -> y f32[b,co,r^3]; scratch: gldm_conv3d_tc_grid_bytes(b, ci, r) bytes (256-byte aligned) */
extern "C" int gldm_conv3d_k3_tc(const float* x, const void* w_img, const float* bias, int b, int ci, int co, int r,
                                 void* scratch, float* y, void* stream) {
  GLDM_REQUIRE(b <= 0 || (x && w_img && y && scratch), "conv3d_k3_tc: null pointer");
  GLDM_REQUIRE(b >= 0 && ci >= 16 && co > 0 && co <= 128 && r > 0, "conv3d_k3_tc: need 16 <= ci, co <= 128");
  if (b == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_cl_pad(x, b, ci, r, scratch, s);
  if (rc) return rc;
  return launch_conv3d(scratch, w_img, bias, b, ci, co, r, y, 0, 0, nullptr, nullptr, s);
}

extern "C" int gldm_cl_pad(const float* x, int b, int c, int r, void* out_cl, void* stream) {
  GLDM_REQUIRE(b <= 0 || (x && out_cl), "cl_pad: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && r > 0, "cl_pad: bad sizes");
  if (b == 0) return GLDM_OK;
  return launch_cl_pad(x, b, c, r, out_cl, (cudaStream_t)stream);
}

static int gn_blocks(int b, int r, int* rpb_out) {
  const int P = (r + 2) * (r + 2) * (r + 2);
  int rpb = 1024;                                  // padded voxels per block: keep >= 4 blocks per SM in flight
  while (rpb > 128 && (long long)((P + rpb - 1) / rpb) * b < 592) rpb >>= 1;
  if (rpb_out) *rpb_out = rpb;
  return (P + rpb - 1) / rpb;
}

extern "C" long long gldm_voxel_ws_bytes(int b, int c, int r) {
  if (b < 0 || c <= 0 || r <= 0) return -1;
  const long long P = (long long)(r + 2) * (r + 2) * (r + 2);
  const long long n_cta = (b * P + 127) / 128, nblk = ((long long)r * r * r + 255) / 256;
  long long d = n_cta * 32;
  if ((long long)b * nblk * 16 > d) d = (long long)b * nblk * 16;
  const long long se = (long long)b * gn_blocks(b > 0 ? b : 1, r, nullptr) * c;
  if (se > d) d = se;
  return d * 8 + 256;
}

extern "C" int gldm_conv3d_tc_cl(const void* x_cl, const void* w_img, const float* bias, int b, int ci, int co, int r,
                                 void* y_cl, int out_fp32, int out_stride, double* stats, void* ws, void* stream) {
  GLDM_REQUIRE(b <= 0 || (x_cl && w_img && y_cl && stats && ws), "conv3d_tc_cl: null pointer");
  GLDM_REQUIRE(b >= 0 && ci >= 16 && co > 0 && co <= 128 && r > 0, "conv3d_tc_cl: need 16 <= ci, co <= 128");
  GLDM_REQUIRE(co % 8 == 0, "conv3d_tc_cl: GroupNorm(8) statistics need co % 8 == 0");
  GLDM_REQUIRE(out_stride >= co && out_stride % 16 == 0 && out_stride <= 128, "conv3d_tc_cl: out_stride must be a multiple of 16 in [co, 128]");
  if (b == 0) return GLDM_OK;
  int rc = launch_conv3d(x_cl, w_img, bias, b, ci, co, r, nullptr, out_fp32 ? 2 : 1, out_stride, y_cl,
                         reinterpret_cast<double*>(ws), (cudaStream_t)stream);
  if (rc) return rc;
  const int P = (r + 2) * (r + 2) * (r + 2);
  conv_stats_finalize_kernel<<<b, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(ws), P, (long long)b * P, stats);
  return check_launch("conv_stats_finalize_kernel");
}

extern "C" int gldm_gn_swish_cl(void* y_cl, int is_fp32, int stride, const double* stats, const float* gamma,
                                const float* beta, int b, int c, int r, float eps, double* se_sum, void* ws, void* stream) {
  GLDM_REQUIRE(b <= 0 || (y_cl && stats && gamma && beta), "gn_swish_cl: null pointer");
  GLDM_REQUIRE(b <= 0 || !se_sum || ws, "gn_swish_cl: the SE sums need the workspace");
  GLDM_REQUIRE(b >= 0 && c > 0 && c % 8 == 0 && c <= 128 && r > 0, "gn_swish_cl: bad sizes");
  GLDM_REQUIRE(stride >= c && stride % 8 == 0 && stride <= 128, "gn_swish_cl: bad row stride");
  if (b == 0) return GLDM_OK;
  int rpb;
  const int gx = gn_blocks(b, r, &rpb);
  dim3 grid(gx, b);
  cudaStream_t s = (cudaStream_t)stream;
  double* part = se_sum ? reinterpret_cast<double*>(ws) : nullptr;
  if (is_fp32) gn_swish_cl_kernel<true><<<grid, 256, 0, s>>>(y_cl, stats, gamma, beta, c, stride, r, eps, part, rpb);
  else gn_swish_cl_kernel<false><<<grid, 256, 0, s>>>(y_cl, stats, gamma, beta, c, stride, r, eps, part, rpb);
  int rc = check_launch("gn_swish_cl_kernel");
  if (rc || !se_sum) return rc;
  block_partials_finalize_kernel<<<b, 128, 0, s>>>(part, gx, c, 0, se_sum);
  return check_launch("block_partials_finalize_kernel");
}

/* narrow-input Conv3d (ci <= 16) on the tensor cores, channels-last output + statistics:
 * w_img: 27 * 4096 bytes from gldm_conv3d_tc16_pack_weight; scratch: b * (r+2)^3 * 32 bytes (256-byte aligned) */
extern "C" long long gldm_conv3d_tc16_weight_bytes(void) { return 27LL * c3::W16_BYTES; }
extern "C" int gldm_conv3d_tc16_pack_weight(const float* w, int co, int ci, void* img, void* stream) {
  GLDM_REQUIRE(w && img, "conv3d_tc16_pack_weight: null pointer");
  GLDM_REQUIRE(co > 0 && co <= 128 && ci > 0 && ci <= 16, "conv3d_tc16_pack_weight: need ci <= 16, co <= 128");
  conv3d_weight_image16_kernel<<<27, 128, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<uint8_t*>(img), co, ci);
  return check_launch("conv3d_weight_image16_kernel");
}
extern "C" int gldm_conv3d_tc16_cl(const float* x, const void* w_img, const float* bias, int b, int ci, int co, int r,
                                   void* scratch, void* y_cl, int out_stride, double* stats, void* ws, void* stream) {
  // x == NULL: `scratch` already holds the zero-padded 16-channel bf16 grid (gldm_voxelize_fused_cl wrote it)
  GLDM_REQUIRE(b <= 0 || (w_img && scratch && y_cl && stats && ws), "conv3d_tc16_cl: null pointer");
  GLDM_REQUIRE(b >= 0 && ci > 0 && ci <= 16 && co > 0 && co <= 128 && co % 8 == 0 && r >= 4 && r <= 62, "conv3d_tc16_cl: bad sizes");
  GLDM_REQUIRE(out_stride >= co && out_stride % 16 == 0 && out_stride <= 128, "conv3d_tc16_cl: bad out_stride");
  if (b == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const long long P = (long long)(r + 2) * (r + 2) * (r + 2), rows = (long long)b * P;
  int rc = GLDM_OK;
  if (x) {
    cl_pad16_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(x, reinterpret_cast<__nv_bfloat16*>(scratch), ci, r, rows);
    rc = check_launch("cl_pad16_kernel");
    if (rc) return rc;
  }
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("conv3d_tc16_cl: cuTensorMapEncodeTiled is not available from the driver");
    return GLDM_ECUDA;
  }
  CUtensorMap map;
  const cuuint64_t gdim[2] = {16, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {32};
  const cuuint32_t box[2] = {16, (cuuint32_t)c3::A3_ROWS};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, scratch, gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("conv3d_tc16_cl: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return GLDM_ECUDA;
  }
  Conv3dTcParams p = {};
  p.w_img = reinterpret_cast<const uint8_t*>(w_img);
  p.bias = bias; p.y = nullptr; p.r = r; p.co = co; p.k_blocks = 1; p.ksteps_last = 1;
  p.rows = rows;
  p.w_rows_bytes = ((co + 7) / 8) * 256;
  p.out_mode = 1; p.out_stride = out_stride; p.y_cl = y_cl; p.stats = reinterpret_cast<double*>(ws); p.batch = b; p.n_acc = 1;
  conv3d_set_magic(p);
  // large grids: persistent, weight-stationary kernel (GLDM_CONV3D_PERSISTENT16: 0 = never, 1 (default) = from two tiles per
  // SM on, 2 = always, for tests; read per call)
  {
    const char* ev = getenv("GLDM_CONV3D_PERSISTENT16");
    const int mode = ev ? atoi(ev) : 1;
    const int n_tiles = (int)((rows + 127) / 128), a_rows = 130 + 2 * (r + 2);
    if (mode && a_rows <= 256 && ((co + 15) & ~15) <= 64 && (mode == 2 || n_tiles >= 2 * kNumSMs)) {
      CUtensorMap mapp;
      const cuuint32_t boxp[2] = {16, (cuuint32_t)a_rows};
      const CUresult crp = enc(&mapp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, scratch, gdim, gstride, boxp, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (crp != CUDA_SUCCESS) {
        set_error("conv3d_tc16_cl: cuTensorMapEncodeTiled (%d-row box) failed (%d)", a_rows, (int)crp);
        return GLDM_ECUDA;
      }
      const int w_slot = (p.w_rows_bytes + 1023) & ~1023, a_slot = (a_rows * 32 + 1023) & ~1023, stages_p = 4;
      const int smem_p = 27 * w_slot + stages_p * a_slot + c3::P_SCRATCH + 768 + 1024;
      static SmemOptIn attr_p;
      if (int rc2 = opt_in_smem(attr_p, conv3d_tc16p_kernel, 27 * 4096 + 4 * 8192 + c3::P_SCRATCH + 768 + 1024, "conv3d_tc16p_kernel"))
        return rc2;
      // two tiles' epilogues in flight (widths with a compile-time epilogue); GLDM_CONV3D_TC16_GROUPS=1: one tile at a time
      const char* evg = getenv("GLDM_CONV3D_TC16_GROUPS");
      if ((co == 48 || co == 32) && !(evg && atoi(evg) == 1)) {
        static SmemOptIn attr_g;
        if (int rc2 = opt_in_smem(attr_g, conv3d_tc16g_kernel, 27 * 4096 + 4 * 8192 + c3::P_SCRATCH + 768 + 1024, "conv3d_tc16g_kernel"))
          return rc2;
        p.n_acc = 1;
        conv3d_tc16g_kernel<<<min(n_tiles, kNumSMs), c3::NTHREADS, smem_p, s>>>(mapp, p, n_tiles, stages_p, a_rows, a_slot, w_slot);
      } else {
        p.n_acc = 3;
        conv3d_tc16p_kernel<<<min(n_tiles, kNumSMs), c3::NTHREADS, smem_p, s>>>(mapp, p, n_tiles, stages_p, a_rows, a_slot, w_slot);
      }
      rc = check_launch("conv3d_tc16p_kernel");
      if (rc) return rc;
      conv_stats_finalize_kernel<<<b, 512, 0, s>>>(reinterpret_cast<const double*>(ws), (int)P, rows, stats);
      return check_launch("conv_stats_finalize_kernel");
    }
  }
  const int smem16 = c3::STAGES16 * (c3::A16_BYTES + 3 * c3::W16_SLOT) + 1024 + 768;
  static SmemOptIn attr16;
  if (int rc2 = opt_in_smem(attr16, conv3d_tc16_kernel, smem16, "conv3d_tc16_kernel")) return rc2;
  conv3d_tc16_kernel<<<(unsigned)((rows + 127) / 128), c3::NTHREADS, smem16, s>>>(map, p);
  rc = check_launch("conv3d_tc16_kernel");
  if (rc) return rc;
  conv_stats_finalize_kernel<<<b, 512, 0, s>>>(reinterpret_cast<const double*>(ws), (int)P, rows, stats);
  return check_launch("conv_stats_finalize_kernel");
}

extern "C" int gldm_block_partials_to_stats(const double* part, int b, int nblk, double* stats, void* stream) {
  GLDM_REQUIRE(b <= 0 || (part && stats), "block_partials_to_stats: null pointer");
  GLDM_REQUIRE(b >= 0 && nblk > 0, "block_partials_to_stats: bad sizes");
  if (b == 0) return GLDM_OK;
  block_partials_finalize_kernel<<<b, 16, 0, (cudaStream_t)stream>>>(part, nblk, 16, 1, stats);
  return check_launch("block_partials_finalize_kernel");
}

extern "C" int gldm_se_gate_sum(const double* sum, int count, const float* w1, const float* w2, int b, int c, int cr,
                                float* gate, void* stream) {
  GLDM_REQUIRE(b <= 0 || (sum && w1 && w2 && gate), "se_gate_sum: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && cr > 0 && count > 0, "se_gate_sum: bad sizes");
  if (b == 0) return GLDM_OK;
  se_gate_sum_kernel<<<b, 128, sizeof(float) * (c + cr), (cudaStream_t)stream>>>(sum, 1.0f / (float)count, w1, w2, c, cr, gate);
  return check_launch("se_gate_sum_kernel");
}

extern "C" int gldm_devox_cl(const float* coords, const void* grid_cl, int is_fp32, int stride, const float* gate,
                             const float* point, int b, int c, int n, int r, float* out, void* stream) {
  GLDM_REQUIRE(b <= 0 || (coords && grid_cl && out), "devox_cl: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && n > 0 && r > 0, "devox_cl: bad sizes");
  GLDM_REQUIRE(stride % 8 == 0 && stride >= ((c + 7) / 8) * 8, "devox_cl: bad row stride");
  if (b == 0) return GLDM_OK;
  dim3 g(ceil_div(n, 128), ceil_div(c, 8), b);
  cudaStream_t s = (cudaStream_t)stream;
  if (is_fp32) devox_cl_kernel<true><<<g, 128, 0, s>>>(coords, grid_cl, gate, point, c, stride, n, r, out);
  else devox_cl_kernel<false><<<g, 128, 0, s>>>(coords, grid_cl, gate, point, c, stride, n, r, out);
  return check_launch("devox_cl_kernel");
}
