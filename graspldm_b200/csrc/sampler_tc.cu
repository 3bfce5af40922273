// Tensor-core (tcgen05 / TMEM) persistent sampler: the whole T-step reverse-diffusion loop of the fpc latent
// denoiser in ONE launch, GEMMs on the 5th-generation tensor cores with bf16 operands and fp32 accumulation.
//
//   TimeConditionedResNet1D.forward   R/models/modules/resnets.py:558-616
//   GaussianDiffusion1D.sample        R/models/diffusion/gaussian_diffusion.py:232-277
//   DDPM / DDIM step                  diffusers (restated in oracle/schedulers.py)
//
// Orientation.  A CTA owns NS = 16 samples for all steps: N = L*NS = 64 GEMM columns (column r = l*16 + s).
// The WEIGHTS are the UMMA A operand (M = 128 output channels per tile, K-major, swizzled images streamed
// L2 -> shared memory with 1-D bulk copies through an 8-stage mbarrier ring) and the ACTIVATIONS are the B
// operand (K-major [column][channel] bf16, SWIZZLE_128B, one 96-row slab per 64 channels with 16 zero halo rows
// on both sides), so a k=3 convolution is three accumulating UMMA groups whose B descriptor start address is
// shifted by -16 / 0 / +16 rows - no im2col.  Accumulators, the fp32 residual stream and the FiLM vectors live in
// TMEM (lane = channel, column = r).  Warp roles: warps 0-7 epilogue (GroupNorm / FiLM / SiLU / LayerNorm / linear
// attention / scheduler update; warp-group g owns samples 8g..8g+7), warp 8 weight producer, warp 9 UMMA issuer.
#include <math.h>

#include "common.cuh"
#include "resnet_layout.cuh"
#include "tc_common.cuh"

namespace gldm {
using namespace tc;

namespace stc {
constexpr int L = 4, NS = 16, NCOL = L * NS, HALO = 16, BROWS = NCOL + 2 * HALO;
constexpr int SLAB = BROWS * 128;          // bytes per 64-channel K-block of the B operand
constexpr int BSLABS = 4;                  // up to 256 channels
constexpr int CHUNK = 16384, STAGES = 8;   // weight ring
constexpr int EMB = 16;
constexpr int MAXJOBS = 40;
constexpr uint32_t T_ACC = 0, T_RES = 192, T_FILM = 320;   // TMEM column map (512 allocated)
constexpr int NCOMPUTE = 256, NTHREADS = 320;

// shared memory map (bytes, from a 1024-aligned base)
constexpr int SM_B = 0;                                  // B operand             49152
constexpr int SM_U = SM_B + BSLABS * SLAB;               // FiLM operand (u)       2048
constexpr int SM_RING = SM_U + 2048;                     // weight ring          131072
constexpr int SM_SCR = SM_RING + STAGES * CHUNK;         // per-warp scratch 8 x 256 floats = 8192
constexpr int SM_XCH = SM_SCR + 8 * 256 * 4;             // cross-warp exchange 2 WG x 4 warps x 64 floats = 2048
constexpr int SM_INEMB = SM_XCH + 2 * 4 * 64 * 4;        // in_emb [16][3][16] floats = 3072
constexpr int SM_X = SM_INEMB + NS * 3 * EMB * 4;        // sampler state [16][4] floats = 256
constexpr int SM_BAR = SM_X + NS * L * 4;                // mbarriers
constexpr int SM_TOTAL = SM_BAR + 256;
}  // namespace stc

struct TcJob {
  uint32_t a_off, bytes;                      // image location in the pack
  uint16_t mtiles, taps, kpt, a_swb, film_tiles, pad;
};

struct TcParams {
  GldmResNetCfg cfg;
  ResNetLayout lay;
  const float* W;          // raw fp32 blob (per-channel parameters are read from here)
  const uint8_t* pack;     // bf16 UMMA images
  TcJob jobs[stc::MAXJOBS];
  int n_jobs;
  int mode;                // 0 sampler, 1 single evaluation (per-sample timestep)
  int n, gpo;
  const float* x_in;       // [n][L]
  const float* z_cond;     // [n_obj][R][cond_dim]
  const float* te;         // time embedding table [n_steps][EMB] (mode 0) or [n][EMB] (mode 1)
  int n_steps;
  const float* coef;       // [n_steps][8]
  int sched_kind, clip;
  const float* noise;
  unsigned long long seed;
  float* x_out;
  float* x_all;
};

// ------------------------------------------------------------------------------------------------
// job table (host): the order of the GEMMs of one network evaluation and where their weight images live
// ------------------------------------------------------------------------------------------------
static int swb_for(int kpt) { return kpt >= 64 ? 128 : kpt >= 32 ? 64 : 32; }
static int pad16(int k) { return (k + 15) & ~15; }

static uint32_t job_bytes(const TcJob& j) {
  const int nkb = (j.kpt * 2 + j.a_swb - 1) / j.a_swb;
  return (uint32_t)j.mtiles * j.taps * nkb * 128 * j.a_swb + (uint32_t)j.film_tiles * 4096;
}

static int build_jobs(const GldmResNetCfg& c, TcJob* jobs, uint32_t* total_bytes) {
  int n = 0;
  uint32_t off = 0;
  auto add = [&](int cout, int cin, int taps, int film_c) {
    TcJob j = {};
    j.mtiles = (uint16_t)((cout + 127) / 128);
    j.taps = (uint16_t)taps;
    j.kpt = (uint16_t)pad16(cin);
    j.a_swb = (uint16_t)swb_for(j.kpt);
    j.film_tiles = (uint16_t)(film_c ? 2 * ((film_c + 127) / 128) : 0);
    j.a_off = off;
    j.bytes = job_bytes(j);
    off += (j.bytes + 1023) & ~1023u;
    jobs[n++] = j;
  };
  for (int s = 0; s < c.n_stages; ++s) {
    const int ch = c.ch[s], cn = c.ch[s + 1];
    for (int rb = 0; rb < 2; ++rb) {
      add(ch, ch, 3, ch);   // block1 (+ FiLM tiles)
      add(ch, ch, 3, 0);    // block2
    }
    add(384, ch, 1, 0);     // to_qkv
    add(ch, 128, 1, 0);     // to_out
    add(cn, ch, 3, 0);      // stage conv
  }
  const int cl = c.ch[c.n_stages];
  add(cl, cl, 3, cl);
  add(cl, cl, 3, 0);
  *total_bytes = off;
  return n;
}

static int check_tc_cfg(const GldmResNetCfg* c) {
  int rc = check_cfg(c);
  if (rc) return rc;
  if (!(c->L == 4 && c->emb_dim == 16 && c->time_cond && c->n_stages == 4 && c->groups == 4 && c->cond_ch <= 3)) {
    set_error("sampler_tc: this build covers the fpc latent denoiser (L=4, emb 16, 4 stages, 4 groups)");
    return GLDM_ENOSUP;
  }
  for (int s = 0; s < c->n_stages; ++s)
    if (c->ch[s] > 128) {
      set_error("sampler_tc: stage width %d > 128", c->ch[s]);
      return GLDM_ENOSUP;
    }
  return GLDM_OK;
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 matrix -> bf16 UMMA image [mtile][tap][kblock][128 rows x SWB bytes] (swizzled)
// element (m, tap, k) of the source is src[(row0 + m) * (cin * taps) + k * taps + tap]
// ------------------------------------------------------------------------------------------------
__global__ void pack_image_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int rows_valid, int row0,
                                  int cin, int taps, int mtiles, int kpt, int a_swb, int standardize) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= mtiles * 128) return;
  const bool valid = m < rows_valid;
  const int K = cin * taps;
  const float* w = src + (size_t)(row0 + m) * K;
  float mean = 0.f, rstd = 1.f;
  if (valid && standardize) {   // resnets.py:85-91: per output channel over (cin, k), biased variance, eps 1e-5
    float s = 0.f;
    for (int i = lane; i < K; i += 32) s += w[i];
    mean = warp_sum(s) / (float)K;
    float q = 0.f;
    for (int i = lane; i < K; i += 32) { const float d = w[i] - mean; q = fmaf(d, d, q); }
    rstd = rsqrtf(warp_sum(q) / (float)K + 1e-5f);
  }
  const int epr = a_swb / 2, nkb = (kpt + epr - 1) / epr;
  const int t = m >> 7, mr = m & 127;
  for (int tap = 0; tap < taps; ++tap)
    for (int k = lane; k < nkb * epr; k += 32) {
      float v = 0.f;
      if (valid && k < cin) v = (w[k * taps + tap] - mean) * rstd;
      const int kb = k / epr, kk = k % epr;
      uint32_t off = (a_swb == 128) ? swz_off<128>(mr, kk >> 3) : (a_swb == 64) ? swz_off<64>(mr, kk >> 3)
                                                                                : swz_off<32>(mr, kk >> 3);
      const size_t blk = ((size_t)(t * taps + tap) * nkb + kb) * 128 * a_swb;
      *reinterpret_cast<__nv_bfloat16*>(dst + blk + off + (kk & 7) * 2) = __float2bfloat16(v);
    }
}

// time embedding table: te[i][:] = time_mlp(t_i)   (resnets.py:44-56, 517-522); one warp per entry
__global__ void time_embed_kernel(const float* __restrict__ W, ResNetLayout lay, int fh, int emb,
                                  const int* __restrict__ ts, int count, float* __restrict__ te) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= count) return;
  __shared__ float s_f[8][40], s_h[8][64];
  float* f = s_f[threadIdx.x >> 5];
  float* h = s_h[threadIdx.x >> 5];
  const float tf = (float)ts[i];
  const int fd = 2 * fh + 1;
  for (int j = lane; j < fd; j += 32) {
    float v = tf;
    if (j > 0) {
      const int q = (j - 1) % fh;
      const float a = __fmul_rn(__fmul_rn(__fmul_rn(tf, __ldg(W + lay.tm_freq + q)), 2.0f), 3.14159274101257324f);
      v = (j - 1 < fh) ? sinf(a) : cosf(a);
    }
    f[j] = v;
  }
  __syncwarp();
  for (int e = lane; e < emb; e += 32) {
    float a = __ldg(W + lay.tm_b1 + e);
    for (int j = 0; j < fd; ++j) a = fmaf(__ldg(W + lay.tm_w1 + e * fd + j), f[j], a);
    h[e] = 0.5f * a * (1.0f + erff(a * 0.70710678118654752440f));
  }
  __syncwarp();
  for (int e = lane; e < emb; e += 32) {
    float a = __ldg(W + lay.tm_b2 + e);
    for (int j = 0; j < emb; ++j) a = fmaf(__ldg(W + lay.tm_w2 + e * emb + j), h[j], a);
    te[(size_t)i * emb + e] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// device helpers of the epilogue warps
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ void wg_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// reduce-scatter step: NOUT pairs (i, i + NOUT); lanes with bit `OFF` set keep the upper element
template <int OFF, int NOUT, int NV>
__device__ __forceinline__ void rs_step(float (&a)[NV], int lane) {
  const bool hi = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const float send = hi ? a[i] : a[i + NOUT];
    const float keep = hi ? a[i + NOUT] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}
// sum of a[idx] over the 32 lanes lands in lane idx (a[0])
__device__ __forceinline__ float reduce_scatter32(float (&a)[32], int lane) {
  rs_step<16, 16>(a, lane);
  rs_step<8, 8>(a, lane);
  rs_step<4, 4>(a, lane);
  rs_step<2, 2>(a, lane);
  rs_step<1, 1>(a, lane);
  return a[0];
}

struct Ctx {
  uint32_t tmem;        // TMEM base with this warp's lane quarter in the lane field
  int lane, q, g, ch;   // lane, quarter (warp & 3), warp-group (sample half), channel within a 128-tile
  uint8_t* smem;
  float* scr;           // per-warp scratch (256 floats)
  float* xch;           // per-warp-group exchange (4 warps x 64 floats)
  uint32_t xoff[8];     // swizzled 16-byte-chunk offset (+ element offset) of this thread's channel for row&7 = j
};

// 32 values of this thread's channel of accumulator/residual tile at column base `col`: v[l*8 + j], j = sample - 8g
__device__ __forceinline__ void tm_load32(const Ctx& c, uint32_t col, float (&v)[32]) {
  uint32_t r[4][8];
#pragma unroll
  for (int l = 0; l < 4; ++l) tmem_ld8(c.tmem + col + l * 16 + c.g * 8, r[l]);
  tmem_ld_wait();
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[l * 8 + j] = __uint_as_float(r[l][j]);
}
__device__ __forceinline__ void tm_store32(const Ctx& c, uint32_t col, const float (&v)[32]) {
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(v[l * 8 + j]);
    tmem_st8(c.tmem + col + l * 16 + c.g * 8, r);
  }
  tmem_st_wait();
}
__device__ __forceinline__ void tm_load8(const Ctx& c, uint32_t col, float (&v)[8]) {
  uint32_t r[8];
  tmem_ld8(c.tmem + col, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

// write this thread's channel (tile t) of the next B operand: rows HALO + l*16 + 8g + j
__device__ __forceinline__ void write_b(const Ctx& c, int t, const float (&v)[32], bool valid) {
  if (!valid) return;
  const int chan = t * 128 + c.ch;
  uint8_t* base = c.smem + stc::SM_B + (chan >> 6) * stc::SLAB + (stc::HALO + c.g * 8) * 128;
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<__nv_bfloat16*>(base + (l * 16 + j) * 128 + c.xoff[j]) = __float2bfloat16(v[l * 8 + j]);
}

// GroupNorm statistics of one tile: per sample j the mean / rstd over (channels of the group x 4 positions).
// GL = lanes per group inside a warp (1, 8, 16, 32); PAIR: the group spans two warps (64 channels).
template <int GL, bool PAIR>
__device__ __forceinline__ void gn_stats(const Ctx& c, const float (&v)[32], float inv_count, float (&mean)[8],
                                         float (&rstd)[8]) {
  float a[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int l = 0; l < 4; ++l) { s += v[l * 8 + j]; q = fmaf(v[l * 8 + j], v[l * 8 + j], q); }
    a[j] = s;
    a[8 + j] = q;
  }
  const int lane = c.lane;
  if (GL == 32) {
    rs_step<16, 8>(a, lane); rs_step<8, 4>(a, lane); rs_step<4, 2>(a, lane); rs_step<2, 1>(a, lane);
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
    // lane holds the total of index (lane >> 1) & 15
    if (!PAIR) {
      if ((lane & 1) == 0) c.scr[lane >> 1] = a[0];
      __syncwarp();
    } else {
      if ((lane & 1) == 0) c.xch[c.q * 64 + (lane >> 1)] = a[0];
      wg_sync(c.g);
    }
  } else if (GL == 16) {
    rs_step<8, 8>(a, lane); rs_step<4, 4>(a, lane); rs_step<2, 2>(a, lane); rs_step<1, 1>(a, lane);
    c.scr[(lane & 16) + (lane & 15)] = a[0];      // [half][idx]
    __syncwarp();
  } else if (GL == 8) {
    rs_step<4, 8>(a, lane); rs_step<2, 4>(a, lane); rs_step<1, 2>(a, lane);
    // lane holds indices 2*(lane&7) + {0,1}
    c.scr[(lane >> 3) * 16 + 2 * (lane & 7)] = a[0];
    c.scr[(lane >> 3) * 16 + 2 * (lane & 7) + 1] = a[1];
    __syncwarp();
  }
  float tot[16];
  if (GL == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) tot[i] = a[i];
  } else {
    const float* src = PAIR ? c.xch + c.q * 64 : c.scr + (GL == 32 ? 0 : (GL == 16 ? (lane & 16) : (lane >> 3) * 16));
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(src + i);
      tot[i] = t4.x; tot[i + 1] = t4.y; tot[i + 2] = t4.z; tot[i + 3] = t4.w;
    }
    if (PAIR) {
      const float* src2 = c.xch + (c.q ^ 1) * 64;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(src2 + i);
        tot[i] += t4.x; tot[i + 1] += t4.y; tot[i + 2] += t4.z; tot[i + 3] += t4.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = tot[j] * inv_count;
    const float var = fmaxf(tot[8 + j] * inv_count - m * m, 0.f);
    mean[j] = m;
    rstd[j] = rsqrtf(var + 1e-5f);
  }
  if (GL != 1) {
    if (PAIR) wg_sync(c.g); else __syncwarp();   // scratch may be rewritten by the next call
  }
}

template <bool PAIR_UNUSED = false>
__device__ __forceinline__ void gn_stats_dispatch(const Ctx& c, int ch_total, const float (&v)[32], float (&mean)[8],
                                                  float (&rstd)[8]) {
  const int cg = ch_total >> 2;                 // channels per group (4 groups)
  const float inv = 1.0f / (float)(cg * 4);
  if (cg >= 64) gn_stats<32, true>(c, v, inv, mean, rstd);
  else if (cg == 32) gn_stats<32, false>(c, v, inv, mean, rstd);
  else if (cg == 16) gn_stats<16, false>(c, v, inv, mean, rstd);
  else if (cg == 8) gn_stats<8, false>(c, v, inv, mean, rstd);
  else gn_stats<1, false>(c, v, inv, mean, rstd);
}

// LayerNorm over the channel axis (one 128-lane tile; lanes >= c hold zeros): per row mean / rstd.
// All four warps of the warp-group call this; out: mr[32] = mean, rs[32] = rstd for rows l*8 + j.
__device__ __forceinline__ void ln_stats(const Ctx& c, int ch_total, const float (&v)[32], float (&mr)[32],
                                         float (&rs)[32]) {
  float a[32], b[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = v[i]; b[i] = v[i] * v[i]; }
  const float s = reduce_scatter32(a, c.lane);      // lane r: row r
  const float q = reduce_scatter32(b, c.lane);
  c.xch[c.q * 64 + c.lane] = s;
  c.xch[c.q * 64 + 32 + c.lane] = q;
  wg_sync(c.g);
  float ts = 0.f, tq = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) { ts += c.xch[w * 64 + c.lane]; tq += c.xch[w * 64 + 32 + c.lane]; }
  const float inv = 1.0f / (float)ch_total;
  const float m = ts * inv;
  const float r = rsqrtf(fmaxf(tq * inv - m * m, 0.f) + 1e-5f);
  c.scr[c.lane] = m;
  c.scr[32 + c.lane] = r;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 m4 = *reinterpret_cast<const float4*>(c.scr + i);
    const float4 r4 = *reinterpret_cast<const float4*>(c.scr + 32 + i);
    mr[i] = m4.x; mr[i + 1] = m4.y; mr[i + 2] = m4.z; mr[i + 3] = m4.w;
    rs[i] = r4.x; rs[i + 1] = r4.y; rs[i + 2] = r4.z; rs[i + 3] = r4.w;
  }
  wg_sync(c.g);     // xch / scr reusable
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(stc::NTHREADS, 1) sampler_tc_kernel(const __grid_constant__ TcParams p) {
  using namespace stc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* b_ready = bars + 2 * STAGES;
  uint64_t* acc_ready = bars + 2 * STAGES + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
  float* s_inemb = reinterpret_cast<float*>(smem + SM_INEMB);
  float* s_x = reinterpret_cast<float*>(smem + SM_X);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * NS;
  const GldmResNetCfg& cfg = p.cfg;
  const ResNetLayout& lay = p.lay;
  const float* W = p.W;
  const int R = cfg.cond_ch;
  const int n_steps = (p.mode == 0) ? p.n_steps : 1;

  // ---- one-time setup
  for (int i = tid; i < (SM_RING) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(b_ready, NCOMPUTE);
    mbar_init(acc_ready, 1);
    fence_barrier_init();
  }
  if (wid == 9) tmem_alloc<512>(tmem_slot);
  // conditioning embedding SiLU(Linear(z_cond))  (resnets.py:531-533,596), once per launch
  for (int idx = tid; idx < NS * R * EMB; idx += NTHREADS) {
    const int e = idx % EMB, r = (idx / EMB) % R, s = idx / (EMB * R);
    float a = 0.f;
    if (s0 + s < p.n) {
      const int obj = (s0 + s) / p.gpo;
      const float* z = p.z_cond + ((size_t)obj * R + r) * cfg.cond_dim;
      const float* w = W + lay.in_w + (size_t)e * cfg.cond_dim;
      a = __ldg(W + lay.in_b + e);
      for (int j = 0; j < cfg.cond_dim; ++j) a = fmaf(__ldg(w + j), __ldg(z + j), a);
      a = a / (1.0f + expf(-a));
    }
    s_inemb[(s * 3 + r) * EMB + e] = a;
  }
  if (tid < NS * L) {
    const int s = tid >> 2, l = tid & 3;
    const float v = (s0 + s < p.n) ? __ldg(p.x_in + (size_t)(s0 + s) * L + l) : 0.f;
    s_x[tid] = v;
    if (p.mode == 0 && p.x_all && s0 + s < p.n) p.x_all[(size_t)(s0 + s) * L + l] = v;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (wid == 8) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int step = 0; step < n_steps; ++step)
        for (int j = 0; j < p.n_jobs; ++j) {
          const TcJob& job = p.jobs[j];
          for (uint32_t off = 0; off < job.bytes; off += CHUNK, ++it) {
            const uint32_t s = it % STAGES, round = it / STAGES;
            if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
            const uint32_t sz = min((uint32_t)CHUNK, job.bytes - off);
            mbar_arrive_expect_tx(&full[s], sz);
            bulk_g2s(smem + SM_RING + s * CHUNK, p.pack + job.a_off + off, sz, &full[s]);
          }
        }
    }
  } else if (wid == 9) {
    // =========================== UMMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc64 = idesc_bf16(128, NCOL), idesc16 = idesc_bf16(128, 16);
      const uint32_t b_base = smem_u32(smem + SM_B), u_base = smem_u32(smem + SM_U), ring = smem_u32(smem + SM_RING);
      uint32_t it = 0, jobn = 0;
      for (int step = 0; step < n_steps; ++step)
        for (int j = 0; j < p.n_jobs; ++j, ++jobn) {
          const TcJob& job = p.jobs[j];
          mbar_wait(b_ready, jobn & 1);
          tc_fence_after();
          const uint32_t a_swb = job.a_swb, blk = 128 * a_swb, epr = a_swb / 2;
          const uint32_t nkb = (job.kpt + epr - 1) / epr, ksteps = min(epr, (uint32_t)job.kpt) / 16;
          const uint32_t a_layout = a_swb == 128 ? SW_128 : a_swb == 64 ? SW_64 : SW_32;
          uint32_t off = 0, sbase = 0;
          auto next_chunk = [&]() {
            if (off != 0) { umma_commit(&empty[(it - 1) % STAGES]); }
            const uint32_t s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            tc_fence_after();
            sbase = ring + s * CHUNK;
            ++it;
          };
          for (uint32_t t = 0; t < job.mtiles; ++t)
            for (uint32_t tap = 0; tap < job.taps; ++tap)
              for (uint32_t kb = 0; kb < nkb; ++kb) {
                if ((off & (CHUNK - 1)) == 0) next_chunk();
                const uint32_t a_addr = sbase + (off & (CHUNK - 1));
                const uint32_t row_off = (job.taps == 3 ? tap : 1u) * HALO * 128;
                for (uint32_t ks = 0; ks < ksteps; ++ks) {
                  const uint32_t k = kb * epr + ks * 16;
                  const uint64_t ad = smem_desc(a_addr + ks * 32, 8 * a_swb, a_layout);
                  const uint64_t bd = smem_desc(b_base + (k >> 6) * SLAB + row_off + (k & 63) * 2, 1024, SW_128);
                  umma_bf16(tmem_base + T_ACC + t * NCOL, ad, bd, idesc64, (tap | kb | ks) != 0);
                }
                off += blk;
              }
          for (uint32_t f = 0; f < job.film_tiles; ++f) {
            if ((off & (CHUNK - 1)) == 0) next_chunk();
            const uint64_t ad = smem_desc(sbase + (off & (CHUNK - 1)), 256, SW_32);
            const uint64_t bd = smem_desc(u_base, 1024, SW_128);
            umma_bf16(tmem_base + T_FILM + f * 16, ad, bd, idesc16, 0);
            off += 4096;
          }
          umma_commit(&empty[(it - 1) % STAGES]);
          umma_commit(acc_ready);
        }
    }
  } else {
    // =========================== epilogue warps ===========================
    Ctx c;
    c.lane = lane; c.q = wid & 3; c.g = wid >> 2; c.ch = c.q * 32 + lane;
    c.tmem = tmem_base + ((uint32_t)(c.q * 32) << 16);
    c.smem = smem;
    c.scr = reinterpret_cast<float*>(smem + SM_SCR) + wid * 256;
    c.xch = reinterpret_cast<float*>(smem + SM_XCH) + c.g * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) c.xoff[j] = ((uint32_t)((((c.ch & 63) >> 3) ^ j) << 4)) + (c.ch & 7) * 2;
    uint32_t jobn = 0;
    auto arrive_b = [&]() {
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(b_ready);
    };
    auto wait_acc = [&]() {
      mbar_wait(acc_ready, jobn & 1);
      ++jobn;
      tc_fence_after();
    };
    const int sgl = c.g * 8;   // first sample of this warp-group inside the CTA

    for (int step = 0; step < n_steps; ++step) {
      // ---- u[s][e] = sum_r silu(time_emb[e] + in_emb[s][r][e])  -> FiLM GEMM operand (bf16)
      {
        const int s = tid >> 4, e = tid & 15;
        const int ti = (p.mode == 0) ? step : min(s0 + s, p.n - 1);
        const float te = __ldg(p.te + (size_t)ti * EMB + e);
        float a = 0.f;
        for (int r = 0; r < R; ++r) { const float z = te + s_inemb[(s * 3 + r) * EMB + e]; a += z / (1.0f + __expf(-z)); }
        *reinterpret_cast<__nv_bfloat16*>(smem + SM_U + swz_off<128>(s, e >> 3) + (e & 7) * 2) = __float2bfloat16(a);
      }
      // ---- init_conv: Conv1d(1 -> ch0, k7, p3) on the state -> residual stream tile 0 and B operand
      {
        const int c0 = cfg.ch[0];
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
        if (c.ch < c0) {
          float w7[7];
#pragma unroll
          for (int t = 0; t < 7; ++t) w7[t] = __ldg(W + lay.init_w + c.ch * 7 + t);
          const float b = __ldg(W + lay.init_b + c.ch);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float xs[4];
#pragma unroll
            for (int l = 0; l < 4; ++l) xs[l] = s_x[(sgl + j) * 4 + l];
#pragma unroll
            for (int l = 0; l < 4; ++l) {
              float a = b;
#pragma unroll
              for (int t = 0; t < 7; ++t) {
                const int ll = l + t - 3;
                if (ll >= 0 && ll < 4) a = fmaf(w7[t], xs[ll], a);
              }
              v[l * 8 + j] = a;
            }
          }
        }
        tm_store32(c, T_RES, v);
        write_b(c, 0, v, c.ch < c0);
      }
      arrive_b();

      // ---- stages
      for (int st = 0; st <= cfg.n_stages; ++st) {
        const bool fin = (st == cfg.n_stages);
        const int ch = cfg.ch[st];
        const int nt = (ch + 127) >> 7;
        const StageOff& so = lay.st[fin ? 0 : st];
        for (int rb = 0; rb < (fin ? 1 : 2); ++rb) {
          const RbOff& o = fin ? lay.fin : so.rb[rb];
          // ======== block1: conv -> GN -> FiLM -> SiLU
          wait_acc();
          for (int t = 0; t < nt; ++t) {
            const int chan = t * 128 + c.ch;
            const bool valid = chan < ch;
            float v[32], mean[8], rstd[8];
            tm_load32(c, T_ACC + t * NCOL, v);
            const float bias = valid ? __ldg(W + o.p1_b + chan) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += bias;
            gn_stats_dispatch(c, ch, v, mean, rstd);
            float fs[8], fh[8];
            tm_load8(c, T_FILM + t * 16 + c.g * 8, fs);
            tm_load8(c, T_FILM + (nt + t) * 16 + c.g * 8, fh);
            const float ga = valid ? __ldg(W + o.n1_w + chan) : 0.f, be = valid ? __ldg(W + o.n1_b + chan) : 0.f;
            const float cs = valid ? (float)R * __ldg(W + o.mlp_b + chan) + (float)R : 0.f;
            const float chh = valid ? (float)R * __ldg(W + o.mlp_b + ch + chan) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float a = rstd[j] * ga, b = be - mean[j] * a;
              const float sc = fs[j] + cs, sh = fh[j] + chh;
#pragma unroll
              for (int l = 0; l < 4; ++l) {
                float y = fmaf(v[l * 8 + j], a, b);
                y = fmaf(y, sc, sh);
                v[l * 8 + j] = silu_fast(y);
              }
            }
            write_b(c, t, v, valid);
          }
          arrive_b();
          // ======== block2: conv -> GN -> SiLU, + residual
          wait_acc();
          const bool to_attn = !fin && rb == 1;
          float fc_part[32];
          if (fin) {
#pragma unroll
            for (int i = 0; i < 32; ++i) fc_part[i] = 0.f;
          }
          for (int t = 0; t < nt; ++t) {
            const int chan = t * 128 + c.ch;
            const bool valid = chan < ch;
            float v[32], mean[8], rstd[8], res[32];
            tm_load32(c, T_ACC + t * NCOL, v);
            const float bias = valid ? __ldg(W + o.p2_b + chan) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += bias;
            gn_stats_dispatch(c, ch, v, mean, rstd);
            tm_load32(c, T_RES + t * NCOL, res);
            const float ga = valid ? __ldg(W + o.n2_w + chan) : 0.f, be = valid ? __ldg(W + o.n2_b + chan) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float a = rstd[j] * ga, b = be - mean[j] * a;
#pragma unroll
              for (int l = 0; l < 4; ++l) {
                const float y = fmaf(v[l * 8 + j], a, b);
                v[l * 8 + j] = valid ? silu_fast(y) + res[l * 8 + j] : 0.f;
              }
            }
            if (fin) {
              const float wfc = valid ? __ldg(W + lay.fc_w + chan) : 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) fc_part[i] = fmaf(wfc, v[i], fc_part[i]);
            } else {
              tm_store32(c, T_RES + t * NCOL, v);
              if (!to_attn) {
                write_b(c, t, v, valid);
              } else {
                // PreNorm LayerNorm (resnets.py:104-124) -> qkv operand
                float mr[32], rs[32];
                ln_stats(c, ch, v, mr, rs);
                const float gg = valid ? __ldg(W + so.ln_g + chan) : 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (v[i] - mr[i]) * rs[i] * gg;
                write_b(c, t, v, valid);
              }
            }
          }
          if (fin) {
            // ======== final_conv (1x1 -> 1 channel) + scheduler update
            const float part = reduce_scatter32(fc_part, lane);     // lane r: row r = l*8 + j of this warp's 32 channels
            c.xch[c.q * 64 + lane] = part;
            wg_sync(c.g);
            if (c.q == 0) {
              float eps = __ldg(W + lay.fc_b);
#pragma unroll
              for (int w = 0; w < 4; ++w) eps += c.xch[w * 64 + lane];
              const int l = lane >> 3, j = lane & 7, s = sgl + j;
              const bool ok = s0 + s < p.n;
              if (p.mode == 0) {
                const float* cf = p.coef + (size_t)step * 8;
                const float x = s_x[s * 4 + l];
                float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(__ldg(cf + 0), eps)), __ldg(cf + 1));
                if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
                float prev;
                if (p.sched_kind == GLDM_SCHED_DDPM) {
                  prev = __fadd_rn(__fmul_rn(__ldg(cf + 2), x0), __fmul_rn(__ldg(cf + 3), x));
                  const float sg = __ldg(cf + 4);
                  if (sg > 0.f && ok) {
                    const float z = p.noise ? __ldg(p.noise + ((size_t)step * p.n + s0 + s) * L + l)
                                            : philox_normal(p.seed, (unsigned)(s0 + s), (unsigned)step, (unsigned)l);
                    prev = __fadd_rn(prev, __fmul_rn(sg, z));
                  }
                } else {
                  prev = __fadd_rn(__fmul_rn(__ldg(cf + 2), x0), __fmul_rn(__ldg(cf + 3), eps));
                }
                s_x[s * 4 + l] = prev;
                if (p.x_all && ok) p.x_all[((size_t)(step + 1) * p.n + s0 + s) * L + l] = prev;
              } else {
                s_x[s * 4 + l] = eps;
              }
            }
            wg_sync(c.g);
          } else {
            arrive_b();
          }
        }
        if (fin) break;
        // ======== linear attention (resnets.py:211-235): qkv -> core -> out operand
        wait_acc();
        {
          float kk[32], e[32];
          tm_load32(c, T_ACC + 1 * NCOL, kk);
#pragma unroll
          for (int j = 0; j < 8; ++j) {   // softmax over the 4 positions (dim=-1)
            const float m = fmaxf(fmaxf(kk[j], kk[8 + j]), fmaxf(kk[16 + j], kk[24 + j]));
            float s = 0.f;
#pragma unroll
            for (int l = 0; l < 4; ++l) { kk[l * 8 + j] = __expf(kk[l * 8 + j] - m); s += kk[l * 8 + j]; }
            const float inv = __fdividef(1.0f, s);
#pragma unroll
            for (int l = 0; l < 4; ++l) kk[l * 8 + j] *= inv;
          }
          tm_load32(c, T_ACC + 0 * NCOL, e);
#pragma unroll
          for (int i = 0; i < 32; ++i) e[i] = __expf(fminf(e[i], 80.f));   // softmax over d: normalised by Z below
          // lane sums over d (this warp = one head): A[j][n'][n] = sum_d k[n'][j] e[n][j], Z[n][j] = sum_d e[n][j]
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            float pr[32];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
              for (int n1 = 0; n1 < 4; ++n1)
#pragma unroll
                for (int n = 0; n < 4; ++n) pr[jj * 16 + n1 * 4 + n] = kk[n1 * 8 + 2 * b + jj] * e[n * 8 + 2 * b + jj];
            c.scr[b * 32 + lane] = reduce_scatter32(pr, lane);
          }
          {
            float z[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = e[i];
            c.scr[128 + lane] = reduce_scatter32(z, lane);
          }
          __syncwarp();
          float vv[32], o[32];
          tm_load32(c, T_ACC + 2 * NCOL, vv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float A[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 t4 = *reinterpret_cast<const float4*>(c.scr + (j >> 1) * 32 + (j & 1) * 16 + i);
              A[i] = t4.x; A[i + 1] = t4.y; A[i + 2] = t4.z; A[i + 3] = t4.w;
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              float acc = 0.f;
#pragma unroll
              for (int n1 = 0; n1 < 4; ++n1) acc = fmaf(vv[n1 * 8 + j], A[n1 * 4 + n], acc);
              const float zinv = __fdividef(0.17677669529663687f, c.scr[128 + n * 8 + j]);   // scale 32^-0.5 / Z
              o[n * 8 + j] = acc * zinv;
            }
          }
          __syncwarp();
          write_b(c, 0, o, true);
        }
        arrive_b();
        // ======== to_out: conv(128 -> ch) + bias -> LayerNorm -> + residual
        wait_acc();
        {
          const bool valid = c.ch < ch;
          float v[32], mr[32], rs[32], res[32];
          tm_load32(c, T_ACC, v);
          const float bias = valid ? __ldg(W + so.out_b + c.ch) : 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += bias;
          ln_stats(c, ch, v, mr, rs);
          tm_load32(c, T_RES, res);
          const float gg = valid ? __ldg(W + so.out_g + c.ch) : 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = valid ? (v[i] - mr[i]) * rs[i] * gg + res[i] : 0.f;
          // the residual stream is replaced by the stage conv below; only the operand is needed
          write_b(c, 0, v, valid);
        }
        arrive_b();
        // ======== stage conv: Conv1d(ch -> cn, k3) + bias -> new residual stream
        wait_acc();
        {
          const int cn = cfg.ch[st + 1];
          const int ntn = (cn + 127) >> 7;
          for (int t = 0; t < ntn; ++t) {
            const int chan = t * 128 + c.ch;
            const bool valid = chan < cn;
            float v[32];
            tm_load32(c, T_ACC + t * NCOL, v);
            const float bias = valid ? __ldg(W + so.down_b + chan) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = valid ? v[i] + bias : 0.f;
            tm_store32(c, T_RES + t * NCOL, v);
            write_b(c, t, v, valid);
          }
        }
        arrive_b();
      }
    }
    // ---- outputs
    wg_sync(c.g);
    if (c.q == 0) {
      const int l = lane >> 3, j = lane & 7, s = sgl + j;
      if (s0 + s < p.n) p.x_out[(size_t)(s0 + s) * L + l] = s_x[s * 4 + l];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 9) tmem_dealloc<512>(tmem_base);
}

static int fill_tc(TcParams& p, const GldmResNetCfg* cfg, const float* raw, const void* pack) {
  int rc = check_tc_cfg(cfg);
  if (rc) return rc;
  GLDM_REQUIRE(raw && pack, "sampler_tc: null weights");
  p.cfg = *cfg;
  make_layout(*cfg, p.lay);
  p.W = raw;
  p.pack = reinterpret_cast<const uint8_t*>(pack);
  uint32_t total;
  p.n_jobs = build_jobs(*cfg, p.jobs, &total);
  return GLDM_OK;
}

static int launch_tc(const TcParams& p, cudaStream_t s) {
  static bool attr = false;
  const int smem = stc::SM_TOTAL + 1024;
  if (!attr) {
    cudaFuncSetAttribute(sampler_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  sampler_tc_kernel<<<ceil_div(p.n, stc::NS), stc::NTHREADS, smem, s>>>(p);
  return check_launch("sampler_tc_kernel");
}

}  // namespace gldm

using namespace gldm;

extern "C" long long gldm_sampler_tc_pack_bytes(const GldmResNetCfg* cfg) {
  if (check_tc_cfg(cfg) != GLDM_OK) return -1;
  TcJob jobs[stc::MAXJOBS];
  uint32_t total = 0;
  build_jobs(*cfg, jobs, &total);
  return (long long)total;
}

extern "C" int gldm_sampler_tc_prepare(const GldmResNetCfg* cfg, const float* raw, void* pack, void* stream) {
  int rc = check_tc_cfg(cfg);
  if (rc) return rc;
  GLDM_REQUIRE(raw && pack, "sampler_tc_prepare: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  ResNetLayout l;
  make_layout(*cfg, l);
  TcJob jobs[stc::MAXJOBS];
  uint32_t total = 0;
  const int nj = build_jobs(*cfg, jobs, &total);
  uint8_t* dst = reinterpret_cast<uint8_t*>(pack);
  cudaMemsetAsync(dst, 0, total, s);
  int launches = 0, ji = 0;
  auto pack_main = [&](const TcJob& j, int src_off, int cout, int cin, int standardize) {
    pack_image_kernel<<<ceil_div(j.mtiles * 128, 8), 256, 0, s>>>(raw + src_off, dst + j.a_off, cout, 0, cin, j.taps,
                                                                  j.mtiles, j.kpt, j.a_swb, standardize);
    ++launches;
  };
  auto pack_film = [&](const TcJob& j, int mlp_off, int ch) {
    const int nkb = (j.kpt * 2 + j.a_swb - 1) / j.a_swb;
    uint8_t* f = dst + j.a_off + (size_t)j.mtiles * j.taps * nkb * 128 * j.a_swb;
    const int ct = (ch + 127) / 128;
    for (int half = 0; half < 2; ++half)
      for (int t = 0; t < ct; ++t) {
        const int row0 = half * ch + t * 128;
        pack_image_kernel<<<ceil_div(128, 8), 256, 0, s>>>(raw + mlp_off, f + (size_t)(half * ct + t) * 4096,
                                                           min(128, ch - t * 128), row0, cfg->emb_dim, 1, 1, 16, 32, 0);
        ++launches;
      }
  };
  for (int st = 0; st < cfg->n_stages; ++st) {
    const int ch = cfg->ch[st], cn = cfg->ch[st + 1];
    for (int rb = 0; rb < 2; ++rb) {
      const RbOff& o = l.st[st].rb[rb];
      pack_main(jobs[ji], o.p1_w, ch, ch, 1);
      pack_film(jobs[ji], o.mlp_w, ch);
      ++ji;
      pack_main(jobs[ji], o.p2_w, ch, ch, 1);
      ++ji;
    }
    pack_main(jobs[ji++], l.st[st].qkv_w, 384, ch, 0);
    pack_main(jobs[ji++], l.st[st].out_w, ch, 128, 0);
    pack_main(jobs[ji++], l.st[st].down_w, cn, ch, 0);
  }
  const int cl = cfg->ch[cfg->n_stages];
  pack_main(jobs[ji], l.fin.p1_w, cl, cl, 1);
  pack_film(jobs[ji], l.fin.mlp_w, cl);
  ++ji;
  pack_main(jobs[ji++], l.fin.p2_w, cl, cl, 1);
  if (ji != nj) {
    set_error("sampler_tc_prepare: job table mismatch");
    return GLDM_EINVAL;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("sampler_tc_prepare: %s", cudaGetErrorString(e));
    return GLDM_ECUDA;
  }
  count_launch(launches);
  return GLDM_OK;
}

static int run_time_embed(const TcParams& p, const int* ts_dev, int count, float* te, cudaStream_t s) {
  time_embed_kernel<<<ceil_div(count, 8), 256, 0, s>>>(p.W, p.lay, p.cfg.fourier_half, p.cfg.emb_dim, ts_dev, count, te);
  return check_launch("time_embed_kernel");
}

extern "C" int gldm_sampler_run_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x_T,
                                   const float* z_obj, int n, int grasps_per_obj, int n_steps,
                                   const int* timesteps_host, const float* coef_host, int sched_kind, int clip_sample,
                                   const float* noise, unsigned long long seed, float* x_out, float* x_all,
                                   void* stream) {
  TcParams p = {};
  int rc = fill_tc(p, cfg, raw, pack);
  if (rc) return rc;
  GLDM_REQUIRE(x_T && z_obj && x_out && timesteps_host && coef_host, "sampler_run_tc: null pointer");
  GLDM_REQUIRE(n >= 0 && grasps_per_obj > 0 && n_steps > 0, "sampler_run_tc: bad sizes");
  GLDM_REQUIRE(sched_kind == GLDM_SCHED_DDPM || sched_kind == GLDM_SCHED_DDIM, "sampler_run_tc: bad scheduler");
  if (n == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  void* scratch = nullptr;
  const size_t cb = sizeof(float) * 8 * (size_t)n_steps, tb = sizeof(int) * (size_t)n_steps,
               eb = sizeof(float) * stc::EMB * (size_t)n_steps;
  if (cudaMallocAsync(&scratch, cb + tb + eb + 256, s) != cudaSuccess) {
    set_error("sampler_run_tc: cudaMallocAsync failed");
    return GLDM_ECUDA;
  }
  float* d_coef = reinterpret_cast<float*>(scratch);
  float* d_te = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + cb);
  int* d_ts = reinterpret_cast<int*>(reinterpret_cast<char*>(scratch) + cb + eb);
  cudaMemcpyAsync(d_coef, coef_host, cb, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_ts, timesteps_host, tb, cudaMemcpyHostToDevice, s);
  p.mode = 0; p.n = n; p.gpo = grasps_per_obj; p.x_in = x_T; p.z_cond = z_obj; p.te = d_te;
  p.n_steps = n_steps; p.coef = d_coef; p.sched_kind = sched_kind; p.clip = clip_sample;
  p.noise = noise; p.seed = seed; p.x_out = x_out; p.x_all = x_all;
  rc = run_time_embed(p, d_ts, n_steps, d_te, s);
  if (rc == GLDM_OK) rc = launch_tc(p, s);
  cudaFreeAsync(scratch, s);
  return rc;
}

extern "C" int gldm_denoiser_forward_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x,
                                        const int* t, const float* z_cond, int n, float* eps, void* stream) {
  TcParams p = {};
  int rc = fill_tc(p, cfg, raw, pack);
  if (rc) return rc;
  GLDM_REQUIRE(x && t && z_cond && eps, "denoiser_forward_tc: null pointer");
  GLDM_REQUIRE(n >= 0, "denoiser_forward_tc: bad n");
  if (n == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* d_te = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&d_te), sizeof(float) * stc::EMB * (size_t)n, s) != cudaSuccess) {
    set_error("denoiser_forward_tc: cudaMallocAsync failed");
    return GLDM_ECUDA;
  }
  p.mode = 1; p.n = n; p.gpo = 1; p.x_in = x; p.z_cond = z_cond; p.te = d_te; p.n_steps = 1; p.x_out = eps;
  rc = run_time_embed(p, t, n, d_te, s);
  if (rc == GLDM_OK) rc = launch_tc(p, s);
  cudaFreeAsync(d_te, s);
  return rc;
}
