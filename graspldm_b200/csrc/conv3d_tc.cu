// Tensor-core (tcgen05 / TMEM) Conv3d k3 p1 of the PVConv voxel branch (R/../pvcnn/modules/pvconv.py:48-67),
// bf16 operands / fp32 accumulation, as an implicit GEMM over a zero-padded channels-last voxel grid.
//
// The input grid is first rewritten (cl_pad_kernel) as X[b * P + p][Cpad] bf16 with P = (r+2)^3 padded voxels per
// cloud (halo = 0) and Cpad = channels rounded up to 64.  For a filter tap (dx,dy,dz) the A operand of a tile of
// 128 consecutive padded voxels is then simply the SAME matrix shifted by dx*(r+2)^2 + dy*(r+2) + dz rows, which one
// 2-D TMA tensor load fetches (SWIZZLE_128B, out-of-range rows read as zero) - no im2col buffer.  The weights are
// pre-packed UMMA images [tap][K block][128 rows x 128 B] streamed with 1-D bulk copies.  One CTA = 128 padded voxels
// x all output channels (UMMA M = 128, N = C_out <= 128): warp 0 producer, warp 1 UMMA issuer, warps 2-5 epilogue
// (bias, drop halo voxels, fp32 store into the reference's [b, c, r^3] layout for GroupNorm / SE / devoxelize).
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gldm {
using namespace tc;

namespace c3 {
constexpr int A_BYTES = 16384;              // 128 voxels x 64 channels bf16
constexpr int W_BYTES = 16384;              // one weight block in the pack
constexpr int W_SLOT = 12288;               // shared-memory slot of a weight block: rows 0..95 (co <= 96), 16384 otherwise
constexpr int STAGES = 3;                   // 3 x 28 KB: two CTAs per SM overlap each other's load latency and epilogue
constexpr int NTHREADS = 192;
}  // namespace c3

struct Conv3dTcParams {
  const uint8_t* w_img;    // [27][k_blocks][16384]  (rows = output channels, zero padded to 128)
  const float* bias;       // [co] or NULL
  float* y;                // [b, co, r^3] fp32
  int r, co, k_blocks, ksteps_last;   // K blocks of 64 channels per tap; UMMAs (1..4) in the last block
  long long rows;          // b * (r+2)^3
  int w_rows_bytes;        // bytes of a weight block actually needed (co rounded up to 8 rows x 128 B)
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
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
    // ---- epilogue: thread <-> padded voxel row; halo voxels and rows past the end are dropped
    const int q = wid & 3;
    const long long m = row0 + q * 32 + lane;
    const int P = rp2 * rp;
    const long long b = m / P;
    const int pp = (int)(m - b * P);
    const int x = pp / rp2, yy = (pp / rp) % rp, z = pp % rp;
    const bool interior = m < p.rows && x >= 1 && x <= p.r && yy >= 1 && yy <= p.r && z >= 1 && z <= p.r;
    const int r3 = p.r * p.r * p.r;
    const int v = ((x - 1) * p.r + (yy - 1)) * p.r + (z - 1);
    float* yb = p.y + ((size_t)b * p.co) * r3 + v;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < p.co; c0 += 16) {
      uint32_t u[16];
      tmem_ld16(taddr + c0, u);
      tmem_ld_wait();
      if (interior) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < p.co) yb[(size_t)(c0 + j) * r3] = __uint_as_float(u[j]) + (p.bias ? __ldg(p.bias + c0 + j) : 0.f);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (wid == 1) tmem_dealloc<128>(tmem);
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

/* x f32[b,ci,r^3] -> y f32[b,co,r^3]; scratch: gldm_conv3d_tc_grid_bytes(b, ci, r) bytes (256-byte aligned) */
extern "C" int gldm_conv3d_k3_tc(const float* x, const void* w_img, const float* bias, int b, int ci, int co, int r,
                                 void* scratch, float* y, void* stream) {
  GLDM_REQUIRE(x && w_img && y && scratch, "conv3d_k3_tc: null pointer");
  GLDM_REQUIRE(b >= 0 && ci >= 16 && co > 0 && co <= 128 && r > 0, "conv3d_k3_tc: need 16 <= ci, co <= 128");
  if (b == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int cpad = ((ci + 63) / 64) * 64, kb = cpad / 64;
  const long long P = (long long)(r + 2) * (r + 2) * (r + 2), rows = (long long)b * P;
  cl_pad_kernel<<<dim3((unsigned)((rows + 255) / 256), cpad / 8), 256, 0, s>>>(x, reinterpret_cast<__nv_bfloat16*>(scratch), ci,
                                                                               cpad, r, rows);
  int rc = check_launch("cl_pad_kernel");
  if (rc) return rc;
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
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, scratch, gdim, gstride, box, estr,
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
  static bool attr = false;
  const int smem_small = c3::STAGES * (c3::A_BYTES + c3::W_SLOT) + 1024 + 256;
  const int smem_big = c3::STAGES * (c3::A_BYTES + c3::W_BYTES) + 1024 + 256;
  if (!attr) {
    cudaFuncSetAttribute(conv3d_tc_kernel<c3::W_SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_small);
    cudaFuncSetAttribute(conv3d_tc_kernel<c3::W_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_big);
    attr = true;
  }
  const unsigned grid = (unsigned)((rows + 127) / 128);
  if (p.w_rows_bytes <= c3::W_SLOT) conv3d_tc_kernel<c3::W_SLOT><<<grid, c3::NTHREADS, smem_small, s>>>(map, p);
  else conv3d_tc_kernel<c3::W_BYTES><<<grid, c3::NTHREADS, smem_big, s>>>(map, p);
  return check_launch("conv3d_tc_kernel");
}
