// Tensor-core (tcgen05 / TMEM) ResNet1D machine.  L = 4: the persistent sampler - the whole T-step reverse-diffusion
// loop of the fpc latent denoiser in ONE launch; L = 16: the grasp decoder trunk + heads.  GEMMs on the
// 5th-generation tensor cores with bf16 operands and fp32 accumulation.
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
constexpr int NCOL = 64;                   // GEMM columns per CTA = L * NS
constexpr int BSLABS = 4;                  // up to 256 channels (64 per K block)
constexpr int CHUNK = 32768;               // weight ring stage (two 16 KB operand blocks)
constexpr int MAXJOBS = 32, MAXCHUNKS = 224, MAXOPS = 400;
// TMEM column map of one sample set (256 columns): accumulator tiles 0..2, the FiLM tiles alias tile 2 (the only
// 3-tile job, to_qkv, has none), residual stream tile 0.  Residual tile 1 (channels 128..255, final block only) lives
// in shared memory as bf16.
constexpr uint32_t T_ACC = 0, T_FILM = 128, T_RES = 192, T_SET = 256;
constexpr int NCOMPUTE = 256;                  // epilogue warps 0..7
constexpr int NTHREADS = 384;                  // + warp 8 weight-ring producer, warps 9 / 10 UMMA issuers, warp 11 idle
}  // namespace stc

// Two instantiations of the same machine:
//   L = 4  : latent denoiser (sampler).  16 samples per CTA, column = position * 16 + sample; the k=3 taps are row
//            shifts of one operand buffer with 16-row zero halos; FiLM / time embedding width 16.
//   L = 16 : grasp decoder trunk (ResNet1D, R/models/grasp_vae.py:401-436).  4 samples per CTA, column = sample * 16 +
//            position; a one-position shift is not a whole 8-row swizzle group, so the epilogue writes three shifted
//            copies of the operand instead (one per tap); embedding width 64.
template <int L_, int NSETS_>
struct Tr {
  static constexpr int L = L_, NS = stc::NCOL / L_, NSETS = NSETS_;
  static constexpr int EMB = (L_ == 4) ? 16 : 64;
  static constexpr int HALO = (L_ == 4) ? 16 : 0;
  static constexpr int BROWS = stc::NCOL + 2 * HALO;
  static constexpr int SLAB = BROWS * 128;                       // bytes per 64-channel K block of the B operand
  static constexpr int COPIES = (L_ == 4) ? 1 : 3;               // operand copies (one per tap for L = 16)
  static constexpr int B_BYTES = COPIES * stc::BSLABS * SLAB;    // operand buffer of one sample set
  static constexpr int STAGES = (L_ == 4 && NSETS_ == 1) ? 4 : 2;   // weight ring depth (x 32 KB)
  static constexpr int SCR = (L_ == 4) ? 256 : 576;              // per-warp scratch floats
  // shared memory map (bytes, from a 1024-aligned base); per-set regions are NSETS consecutive copies
  static constexpr int SM_B = 0;
  static constexpr int SM_U = SM_B + NSETS * B_BYTES;               // FiLM operand (u): 16 rows x 128 B per set
  static constexpr int SM_RING = SM_U + NSETS * 2048;
  static constexpr int SM_RES1 = SM_RING + STAGES * stc::CHUNK;     // residual tile 1: [32 values][256 threads] bf16 per set
  static constexpr int SM_SCR = SM_RES1 + NSETS * 16384;
  static constexpr int SM_XCH = SM_SCR + 8 * SCR * 4;               // cross-warp exchange 2 WG x 4 warps x 64 floats
  static constexpr int SM_INEMB = SM_XCH + 2 * 4 * 64 * 4;          // in_emb [NS][3][EMB] floats = 3072 per set
  static constexpr int SM_X = SM_INEMB + NSETS * 3072;              // state / trunk input [NS][L] floats = 256 per set
  static constexpr int SM_BAR = SM_X + NSETS * 256;                 // mbarriers
  static constexpr int SM_CHUNKS = SM_BAR + 256;                    // weight chunk table {pack offset, bytes}
  static constexpr int SM_JOBS = SM_CHUNKS + stc::MAXCHUNKS * 8;    // job table copy
  static constexpr int SM_OPS = SM_JOBS + stc::MAXJOBS * 64;        // UMMA op table, 16 bytes per block of <= 4 UMMAs
  static constexpr int SM_OPBEG = SM_OPS + stc::MAXOPS * 16;        // first op of every (job, set)
  static constexpr int SM_TOTAL = SM_OPBEG + (2 * stc::MAXJOBS + 2) * 2 + 16;
};

// epilogue recipe of a job (what the epilogue warps do with its accumulator)
enum : uint16_t {
  E_GN = 1,         // + bias, GroupNorm * gamma + beta
  E_FILM = 2,       // * (scale + 1) + shift   (FiLM tiles of the same job)
  E_SILU = 4,
  E_LN = 8,         // + bias, channel LayerNorm * g           (to_out)
  E_ADDRES = 16,    // + residual stream
  E_STORERES = 32,  // result becomes the residual stream
  E_LNNEXT = 64,    // operand of the next job = LayerNorm(result) * g2   (PreNorm of the attention)
  E_FINAL = 128,    // no operand: final_conv dot product (weights at o_g2) + scheduler update
  E_ATTN = 256,     // linear-attention core on the q, k, v tiles
};
struct TcJob {
  uint32_t a_off, bytes;                      // image location in the pack
  uint16_t mtiles, taps, kpt, a_swb, film_tiles, flags;
  int ch;                                     // valid output channels
  int o_bias, o_gamma, o_beta, o_mlpb, o_g, o_g2;   // float offsets into the raw blob (-1: unused)
  int o_mlpw;                                 // FiLM projection Linear(emb -> 2 ch) weight [2 ch][emb] (-1: no FiLM)
  int o_film;                                 // float offset of this layer's [A (ch) | B (ch)] pair in a FiLM table row
};
static_assert(sizeof(TcJob) <= 64, "job table slot");

struct TcParams {
  GldmResNetCfg cfg;
  ResNetLayout lay;
  const float* W;          // raw fp32 blob (per-channel parameters are read from here)
  const uint8_t* pack;     // bf16 UMMA images
  TcJob jobs[stc::MAXJOBS];
  int n_jobs;
  int mode;                // 0 sampler, 1 single evaluation (per-sample timestep), 2 decoder (L = 16)
  int n, gpo;
  const float* x_in;       // [n][L]
  const float* z_cond;     // [n_obj][R][cond_dim]
  const float* te;         // time embedding table [n_steps][EMB] (mode 0) or [n][EMB] (mode 1); unused in mode 2
  int n_steps;
  const float* coef;       // [n_steps][8]
  int sched_kind, clip;
  const float* noise;
  unsigned long long seed;
  float* x_out;
  float* x_all;
  const float* head;       // decoder: in_w[L][D] in_b[L] tmrp_w[6][L] tmrp_b[6] cls_w[L] cls_b[1]
  int D;
  float* tmrp;             // [n][6]
  float* logit;            // [n]
  long long* prof;         // development aid: per-job clock stamps of CTA 0 (NULL in production)
  // row-major kernel: FiLM hoisted out of the sample loop.  film[(obj * n_steps + step) * film_stride + o_film + ...]
  // (mode 1: one row per sample) holds, per FiLM layer, A = gamma * S and B = beta * S + H with S / H the summed
  // multi-channel scale (+1 each) / shift of resnets.py:163-175: GroupNorm affine and FiLM as ONE multiply-add.
  const float* film;
  int film_stride;
  const float* cls_emb;    // [n_obj][EMB] added to the time embedding (class-conditioned denoiser) or NULL
};

// ------------------------------------------------------------------------------------------------
// job table (host): the order of the GEMMs of one network evaluation and where their weight images live
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int swb_for(int kpt) { return kpt >= 64 ? 128 : kpt >= 32 ? 64 : 32; }
__host__ __device__ inline int pad16(int k) { return (k + 15) & ~15; }

// FiLM projection tiles: [128 rows x emb] with K = emb (16 -> SWIZZLE_32B 4 KB tiles, 64 -> SWIZZLE_128B 16 KB tiles)
static uint32_t film_tile_bytes(int emb) { return 128u * swb_for(pad16(emb)); }
static uint32_t main_bytes(const TcJob& j) {
  const int nkb = (j.kpt * 2 + j.a_swb - 1) / j.a_swb;
  return (uint32_t)j.mtiles * j.taps * nkb * 128 * j.a_swb;
}
// blocks of a job are stored in descending size so that none straddles a ring stage: FiLM tiles first when they
// are larger than the main blocks
static bool film_first(const TcJob& j, int emb) { return j.film_tiles && film_tile_bytes(emb) > 128u * j.a_swb; }
static uint32_t job_bytes(const TcJob& j, int emb) { return main_bytes(j) + (uint32_t)j.film_tiles * film_tile_bytes(emb); }

static int build_jobs(const GldmResNetCfg& c, TcJob* jobs, uint32_t* total_bytes, int* film_stride = nullptr) {
  ResNetLayout l;
  make_layout(c, l);
  int n = 0, film_off = 0, o_mlpw_next = -1;
  uint32_t off = 0;
  auto add = [&](int cout, int cin, int taps, int film_c, uint16_t flags, int o_bias, int o_gamma, int o_beta,
                 int o_mlpb, int o_g, int o_g2) {
    TcJob j = {};
    j.mtiles = (uint16_t)((cout + 127) / 128);
    j.taps = (uint16_t)taps;
    j.kpt = (uint16_t)pad16(cin);
    j.a_swb = (uint16_t)swb_for(j.kpt);
    j.film_tiles = (uint16_t)(film_c ? 2 * ((film_c + 127) / 128) : 0);
    j.flags = flags;
    j.ch = cout;
    j.o_bias = o_bias; j.o_gamma = o_gamma; j.o_beta = o_beta; j.o_mlpb = o_mlpb; j.o_g = o_g; j.o_g2 = o_g2;
    j.o_mlpw = -1; j.o_film = -1;
    if (film_c) { j.o_mlpw = o_mlpw_next; j.o_film = film_off; film_off += 2 * film_c; }
    j.a_off = off;
    j.bytes = job_bytes(j, c.emb_dim);
    off += (j.bytes + 1023) & ~1023u;
    jobs[n++] = j;
  };
  auto add_rb = [&](const RbOff& o, int ch, uint16_t extra2, int o_g2) {
    o_mlpw_next = o.mlp_w;
    add(ch, ch, 3, ch, E_GN | E_FILM | E_SILU, o.p1_b, o.n1_w, o.n1_b, o.mlp_b, -1, -1);
    add(ch, ch, 3, 0, (uint16_t)(E_GN | E_SILU | E_ADDRES | extra2), o.p2_b, o.n2_w, o.n2_b, -1, -1, o_g2);
  };
  for (int s = 0; s < c.n_stages; ++s) {
    const int ch = c.ch[s], cn = c.ch[s + 1];
    const StageOff& so = l.st[s];
    add_rb(so.rb[0], ch, E_STORERES, -1);
    add_rb(so.rb[1], ch, E_STORERES | E_LNNEXT, so.ln_g);
    add(384, ch, 1, 0, E_ATTN, -1, -1, -1, -1, -1, -1);                              // to_qkv
    add(ch, 128, 1, 0, E_LN | E_ADDRES, so.out_b, -1, -1, -1, so.out_g, -1);          // to_out
    add(cn, ch, 3, 0, E_STORERES, so.down_b, -1, -1, -1, -1, -1);                     // stage conv
  }
  const int cl = c.ch[c.n_stages];
  add_rb(l.fin, cl, E_FINAL, l.fc_w);
  *total_bytes = off;
  if (film_stride) *film_stride = film_off;
  return n;
}

static int check_tc_cfg(const GldmResNetCfg* c) {
  int rc = check_cfg(c);
  if (rc) return rc;
  const bool l4 = c->L == 4 && c->emb_dim == 16 && c->time_cond;     // fpc latent denoiser
  const bool l16 = c->L == 16 && c->emb_dim == 64;                    // ppc latent denoiser, grasp decoder trunk
  if (!((l4 || l16) && c->n_stages == 4 && c->groups == 4 && c->cond_ch <= 3 && c->cond_dim <= 1024)) {
    set_error("resnet_tc: this build covers the latent denoisers (L=4 / emb 16, L=16 / emb 64, time conditioned) and the "
              "grasp decoder trunk (L=16, emb 64), 4 stages, 4 groups");
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
// block order: tile-major [t][tap][kb], or - when `interleave` - [tap][kb][t] so that the two UMMA issuers, which own
// the even / odd output tiles, walk the weight stream side by side
__global__ void pack_image_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int rows_valid, int row0,
                                  int cin, int taps, int mtiles, int kpt, int a_swb, int standardize, int interleave) {
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
      const size_t bi = interleave ? ((size_t)(tap * nkb + kb) * mtiles + t) : ((size_t)(t * taps + tap) * nkb + kb);
      const size_t blk = bi * 128 * a_swb;
      *reinterpret_cast<__nv_bfloat16*>(dst + blk + off + (kk & 7) * 2) = __float2bfloat16(v);
    }
}

// time embedding table: te[i][:] = time_mlp(t_i)   (resnets.py:44-56, 517-522); one warp per entry
__global__ void time_embed_kernel(const float* __restrict__ W, ResNetLayout lay, int fh, int emb,
                                  const int* __restrict__ ts, const float* __restrict__ tsf, int count,
                                  float* __restrict__ te) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= count) return;
  __shared__ float s_f[8][40], s_h[8][64];
  float* f = s_f[threadIdx.x >> 5];
  float* h = s_h[threadIdx.x >> 5];
  const float tf = tsf ? tsf[i] : (float)ts[i];      // continuous time for the elucidated sampler
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
// x * sigmoid(x) = 0.5 x (1 + tanh(x / 2)): one MUFU op per value (the epilogues are MUFU-throughput sensitive)
__device__ __forceinline__ float silu_fast(float x) {
  float t;
  const float h = 0.5f * x;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// packed fp32 pairs (FFMA2 / FADD2 / FMUL2): the epilogues are issue-slot bound, one instruction per two values
__device__ __forceinline__ float2 ld2(const float (&v)[32], int i) { return make_float2(v[2 * i], v[2 * i + 1]); }
__device__ __forceinline__ void st2(float (&v)[32], int i, float2 x) { v[2 * i] = x.x; v[2 * i + 1] = x.y; }
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 silu_fast2(float2 x) {
  const float2 h = __fmul2_rn(x, bc2(0.5f));
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
  return __ffma2_rn(h, t, h);
}

__device__ __forceinline__ void wg_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

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
  uint8_t* bbase;       // B operand buffer of the current sample set
  float* scr;           // per-warp scratch (256 floats)
  float* xch;           // per-warp-group exchange (4 warps x 64 floats)
  uint32_t xoff[8];     // swizzled 16-byte-chunk offset (+ element offset) of this thread's channel for row&7 = j
};

// 32 values of this thread's channel of an accumulator / residual tile at column base `col`.
//   L = 4 : v[l*8 + j]  <-> column l*16 + 8g + j   (j = sample - 8g)
//   L = 16: v[jj*16 + l] <-> column 32g + jj*16 + l (jj = sample - 2g)
// tm_issue* start the asynchronous loads; tm_wait() + tm_use*() make the registers safe to read.
template <int L>
__device__ __forceinline__ void tm_issue32(const Ctx& c, uint32_t col, uint32_t (&r)[32]) {
  if (L == 4) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      uint32_t(&q)[8] = *reinterpret_cast<uint32_t(*)[8]>(&r[l * 8]);
      tmem_ld8(c.tmem + col + l * 16 + c.g * 8, q);
    }
  } else {
    tmem_ld32(c.tmem + col + c.g * 32, r);
  }
}
__device__ __forceinline__ void tm_issue8(const Ctx& c, uint32_t col, uint32_t (&r)[8]) { tmem_ld8(c.tmem + col, r); }
__device__ __forceinline__ void tm_issue4(const Ctx& c, uint32_t col, uint32_t (&r)[4]) { tmem_ld4(c.tmem + col, r); }
__device__ __forceinline__ void tm_wait() { tmem_ld_wait(); }
template <int N>
__device__ __forceinline__ void tm_use(const uint32_t (&r)[N], float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    uint32_t x = r[i];
    asm volatile("" : "+r"(x));     // ordered after the tcgen05.wait::ld above (volatile asm keeps program order)
    v[i] = __uint_as_float(x);
  }
}
template <int L>
__device__ __forceinline__ void tm_load32(const Ctx& c, uint32_t col, float (&v)[32]) {
  uint32_t r[32];
  tm_issue32<L>(c, col, r);
  tm_wait();
  tm_use(r, v);
}
template <int L>
__device__ __forceinline__ void tm_store32(const Ctx& c, uint32_t col, const float (&v)[32]) {
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(v[l * 8 + j]);
    tmem_st8(c.tmem + col + (L == 4 ? l * 16 + c.g * 8 : c.g * 32 + l * 8), r);
  }
  tmem_st_wait();
}

// write this thread's channel (tile t) of the next B operand.
//   L = 4 : rows HALO + l*16 + 8g + j of the single (halo-padded) buffer
//   L = 16: row 32g + i of the centre copy, and the rows one position later / earlier (same sample) of the tap-0 /
//           tap-2 copies: B_tap[(s, l)] = x[s][l + tap - 1]; the never-written boundary rows stay zero
template <int L>
__device__ __forceinline__ void write_b(const Ctx& c, int t, const float (&v)[32], bool valid) {
  if (!valid) return;
  using T = Tr<L, 1>;
  const int chan = t * 128 + c.ch;
  if (L == 4) {
    uint8_t* base = c.bbase + (chan >> 6) * T::SLAB + (T::HALO + c.g * 8) * 128;
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<__nv_bfloat16*>(base + (l * 16 + j) * 128 + c.xoff[j]) = __float2bfloat16(v[l * 8 + j]);
  } else {
    uint8_t* base = c.bbase + (chan >> 6) * T::SLAB + (c.g * 32) * 128;
    constexpr int COPY = stc::BSLABS * T::SLAB;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const __nv_bfloat16 h = __float2bfloat16(v[i]);
      *reinterpret_cast<__nv_bfloat16*>(base + COPY + i * 128 + c.xoff[i & 7]) = h;
      if ((i & 15) != 15) *reinterpret_cast<__nv_bfloat16*>(base + (i + 1) * 128 + c.xoff[(i + 1) & 7]) = h;
      if ((i & 15) != 0) *reinterpret_cast<__nv_bfloat16*>(base + 2 * COPY + (i - 1) * 128 + c.xoff[(i - 1) & 7]) = h;
    }
  }
}

// GroupNorm statistics of one tile: per sample j the mean / rstd over (channels of the group x 4 positions).
// GL = lanes per group inside a warp (1, 8, 16, 32); PAIR: the group spans two warps (64 channels).
template <int GL, bool PAIR>
__device__ __forceinline__ void gn_stats(const Ctx& c, const float (&v)[32], float inv_count, float (&mean)[8],
                                         float (&rstd)[8]) {
  float a[16];
#pragma unroll
  for (int jp = 0; jp < 4; ++jp) {      // sample pair (2jp, 2jp + 1): v[l*8 + 2jp], v[l*8 + 2jp + 1] are register pairs
    float2 s = ld2(v, jp), q = __fmul2_rn(s, s);
#pragma unroll
    for (int l = 1; l < 4; ++l) {
      const float2 x = ld2(v, l * 4 + jp);
      s = __fadd2_rn(s, x);
      q = __ffma2_rn(x, x, q);
    }
    a[2 * jp] = s.x; a[2 * jp + 1] = s.y;
    a[8 + 2 * jp] = q.x; a[8 + 2 * jp + 1] = q.y;
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

// L = 16 (two samples per warp-group, 16 positions each): statistics over (channels of the group x 16 positions).
// Groups are cg = c/4 >= 4 consecutive channels: butterfly over min(cg, 32) lanes, + the partner warp for cg = 64.
__device__ __forceinline__ void gn_stats16(const Ctx& c, int cg, const float (&v)[32], float (&mean)[8], float (&rstd)[8]) {
  float a[4];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    float2 s = ld2(v, jj * 8), q = __fmul2_rn(s, s);
#pragma unroll
    for (int lp = 1; lp < 8; ++lp) {
      const float2 x = ld2(v, jj * 8 + lp);
      s = __fadd2_rn(s, x);
      q = __ffma2_rn(x, x, q);
    }
    a[jj] = s.x + s.y;
    a[2 + jj] = q.x + q.y;
  }
  const int gl = cg < 32 ? cg : 32;
  for (int o = 1; o < gl; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
  }
  if (cg > 32) {
    if (c.lane == 0) *reinterpret_cast<float4*>(c.xch + c.q * 64) = make_float4(a[0], a[1], a[2], a[3]);
    wg_sync(c.g);
    const float4 o4 = *reinterpret_cast<const float4*>(c.xch + (c.q ^ 1) * 64);
    a[0] += o4.x; a[1] += o4.y; a[2] += o4.z; a[3] += o4.w;
    wg_sync(c.g);
  }
  const float inv = 1.0f / (float)(cg * 16);
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const float m = a[jj] * inv;
    mean[jj] = m;
    rstd[jj] = rsqrtf(fmaxf(a[2 + jj] * inv - m * m, 0.f) + 1e-5f);
  }
}

template <int L>
__device__ __forceinline__ void gn_stats_dispatch(const Ctx& c, int ch_total, const float (&v)[32], float (&mean)[8],
                                                  float (&rstd)[8]) {
  const int cg = ch_total >> 2;                 // channels per group (4 groups)
  if (L == 16) {
    gn_stats16(c, cg, v, mean, rstd);
    return;
  }
  const float inv = 1.0f / (float)(cg * 4);
  if (cg >= 64) gn_stats<32, true>(c, v, inv, mean, rstd);
  else if (cg == 32) gn_stats<32, false>(c, v, inv, mean, rstd);
  else if (cg == 16) gn_stats<16, false>(c, v, inv, mean, rstd);
  else if (cg == 8) gn_stats<8, false>(c, v, inv, mean, rstd);
  else gn_stats<1, false>(c, v, inv, mean, rstd);
}

// LayerNorm over the channel axis (one 128-lane tile; lanes >= c hold zeros): per row mean / rstd.
// All four warps of the warp-group call this; out: mr[32] = -mean, rs[32] = rstd for rows l*8 + j.
__device__ __forceinline__ void ln_stats(const Ctx& c, int ch_total, const float (&v)[32], float (&mr)[32],
                                         float (&rs)[32]) {
  float a[32], b[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = ld2(v, i);
    st2(a, i, x);
    st2(b, i, __fmul2_rn(x, x));
  }
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
  c.scr[c.lane] = -m;           // callers apply v * (rstd * g) + (-mean) * (rstd * g)
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

// v = (v - mean) * rstd * g as one packed FMA per value pair
__device__ __forceinline__ void ln_apply(const Ctx& c, int ch_total, float g, float (&v)[32]) {
  float mr[32], rs[32];
  ln_stats(c, ch_total, v, mr, rs);
  const float2 g2 = bc2(g);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 A = __fmul2_rn(ld2(rs, i), g2);
    st2(v, i, __ffma2_rn(ld2(v, i), A, __fmul2_rn(ld2(mr, i), A)));
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int L, int NSETS>
__global__ void __launch_bounds__(stc::NTHREADS, 1) resnet_tc_kernel(const __grid_constant__ TcParams p) {
  using namespace stc;
  using T = Tr<L, NSETS>;
  constexpr int NS = T::NS, EMB = T::EMB, STAGES = T::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // pointer arithmetic (no integer round trip) keeps the shared address space visible to the compiler: LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T::SM_BAR);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + 4;            // [STAGES]
  uint64_t* b_ready = bars + 8;          // [NSETS]
  uint64_t* acc_ready = bars + 10;       // [NSETS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  TcJob* s_jobs = reinterpret_cast<TcJob*>(smem + T::SM_JOBS);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int cta_s0 = blockIdx.x * NS * NSETS;      // first sample of this CTA; set k holds samples cta_s0 + k*NS ...
  const GldmResNetCfg& cfg = p.cfg;
  const ResNetLayout& lay = p.lay;
  const float* W = p.W;
  const int R = cfg.cond_ch;
  const int n_steps = (p.mode == 0) ? p.n_steps : 1;
  const int n_jobs = p.n_jobs;

  // ---- one-time setup
  for (int i = tid; i < (T::SM_RING) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < n_jobs * (int)(sizeof(TcJob) / 4); i += NTHREADS)
    reinterpret_cast<uint32_t*>(s_jobs)[i] = reinterpret_cast<const uint32_t*>(p.jobs)[i];
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }   // released by both issuers
    for (int k = 0; k < NSETS; ++k) { mbar_init(&b_ready[k], NCOMPUTE / 32); mbar_init(&acc_ready[k], 2); }
    fence_barrier_init();
  }
  if (wid == 0) tmem_alloc<512>(tmem_slot);
  // conditioning embedding SiLU(Linear(z_cond))  (resnets.py:531-533,596), once per launch
  for (int idx = tid; idx < NSETS * NS * R * EMB; idx += NTHREADS) {
    const int e = idx % EMB, r = (idx / EMB) % R, s = idx / (EMB * R);     // s: sample within the CTA
    float a = 0.f;
    if (cta_s0 + s < p.n) {
      const int obj = (cta_s0 + s) / p.gpo;
      const float* z = p.z_cond + ((size_t)obj * R + r) * cfg.cond_dim;
      const float* w = W + lay.in_w + (size_t)e * cfg.cond_dim;
      a = __ldg(W + lay.in_b + e);
      for (int j = 0; j < cfg.cond_dim; ++j) a = fmaf(__ldg(w + j), __ldg(z + j), a);
      a = a / (1.0f + expf(-a));
    }
    reinterpret_cast<float*>(smem + T::SM_INEMB + (s / NS) * 3072)[((s % NS) * 3 + r) * EMB + e] = a;
  }
  if (tid < NSETS * NS * L) {
    const int s = tid / L, l = tid % L;                                    // s: sample within the CTA
    float v = 0.f;
    if (cta_s0 + s < p.n) {
      if (p.mode != 2) {   // denoiser: the latent itself
        v = __ldg(p.x_in + (size_t)(cta_s0 + s) * L + l);
      } else {   // decoder in_layer: Linear(D -> L)   (grasp_vae.py:419)
        v = __ldg(p.head + L * p.D + l);
        for (int d = 0; d < p.D; ++d) v = fmaf(__ldg(p.head + l * p.D + d), __ldg(p.x_in + (size_t)(cta_s0 + s) * p.D + d), v);
      }
    }
    reinterpret_cast<float*>(smem + T::SM_X + (s / NS) * 256)[(s % NS) * L + l] = v;
    if (p.mode == 0 && p.x_all && cta_s0 + s < p.n) p.x_all[(size_t)(cta_s0 + s) * L + l] = v;
  }
  // ---- weight chunk table and UMMA op table of one step: for every job, for every sample set (the order in which
  //      the epilogue warps hand operands over).  The ring stage of a chunk is a static function of its position.
  uint2* chunk_tab = reinterpret_cast<uint2*>(smem + T::SM_CHUNKS);
  uint4* ops = reinterpret_cast<uint4*>(smem + T::SM_OPS);
  uint16_t* op_begin = reinterpret_cast<uint16_t*>(smem + T::SM_OPBEG);   // index = job * NSETS + set
  // op.w bits: [0,9) TMEM column, [9,12) UMMAs in the block (1/2/4), 12 accumulate-first, 13 first block of a chunk,
  //            14 FiLM tile (N = 16, operand u), 15 first block of the job, [16,19) ring stage, 19 ring padding (no UMMA),
  //            20 owner (which of the two issuer warps executes it; both walk every op for the ring bookkeeping)
  // Built by the whole CTA: thread 0 only lays out the prefix sums (first op / first chunk of every (job, set) entry), then
  // every op and chunk is written from a closed form of its index.  (A serial build by one thread cost more than the
  // network evaluation itself in the single-evaluation launches - the decoder runs 4 samples per CTA.)
  uint32_t* ent_chunk0 = reinterpret_cast<uint32_t*>(smem + T::SM_SCR);                 // [n_jobs * NSETS + 1] first chunk of an entry
  {
    const uint32_t f_swb = swb_for(pad16(EMB)), f_bytes = 128u * f_swb;
    if (tid == 0) {
      uint32_t nops = 0, ncp = 0;
      for (int j = 0; j < n_jobs; ++j) {
        const TcJob& job = p.jobs[j];
        const uint32_t nkb = job.a_swb == 128 ? (uint32_t)job.kpt >> 6 : 1u;
        const uint32_t n_e = job.film_tiles + job.mtiles * job.taps * nkb, c_e = (job.bytes + CHUNK - 1) / CHUNK;
        for (int set = 0; set < NSETS; ++set) {
          op_begin[j * NSETS + set] = (uint16_t)nops;
          ent_chunk0[j * NSETS + set] = ncp;
          nops += n_e;
          ncp += c_e;
        }
      }
      ent_chunk0[n_jobs * NSETS] = ncp;
      // a step must span a whole number of ring revolutions: pad with 16-byte dummy chunks consumed by no-op entries
      // of the last (job, set)
      while (ncp % STAGES) {
        ops[nops++] = make_uint4(0, 0, 0, (1u << 13) | (1u << 19) | ((ncp % STAGES) << 16));
        chunk_tab[ncp++] = make_uint2(p.jobs[0].a_off, 16u);
      }
      op_begin[n_jobs * NSETS] = (uint16_t)nops;
      op_begin[n_jobs * NSETS + 1] = (uint16_t)ncp;      // chunks per step
    }
    __syncthreads();
    const uint32_t ring_a = smem_u32(smem + T::SM_RING);
    const uint32_t f_hi = ((8u * f_swb) >> 4) | (1u << 14) | ((f_swb == 128 ? (uint32_t)SW_128 : (uint32_t)SW_32) << 29);
    for (int e = 0; e < n_jobs * NSETS; ++e) {
      const int j = e / NSETS, set = e % NSETS;
      const TcJob& job = p.jobs[j];
      const uint32_t op0 = op_begin[e], chunk_base = ent_chunk0[e];
      const uint32_t a_swb = job.a_swb, blk = a_swb << 7, nkb = a_swb == 128 ? (uint32_t)job.kpt >> 6 : 1u;
      const uint32_t nfilm = job.film_tiles, nblk = job.mtiles * job.taps * nkb;
      for (uint32_t c = tid; c * CHUNK < job.bytes; c += NTHREADS)
        chunk_tab[chunk_base + c] = make_uint2(job.a_off + c * CHUNK, min((uint32_t)CHUNK, job.bytes - c * CHUNK));
      if ((uint32_t)tid >= nfilm + nblk) continue;
      const uint32_t b_base = smem_u32(smem + T::SM_B + set * T::B_BYTES), u_base = smem_u32(smem + T::SM_U + set * 2048);
      const uint32_t a_hi = ((8u * a_swb) >> 4) | (1u << 14) |
                            ((a_swb == 128 ? (uint32_t)SW_128 : a_swb == 64 ? (uint32_t)SW_64 : (uint32_t)SW_32) << 29);
      const bool ffirst = nfilm && f_bytes > blk;           // blocks in descending size (see film_first)
      for (uint32_t i = tid; i < nfilm + nblk; i += NTHREADS) {
        const bool is_film = ffirst ? i < nfilm : i >= nblk;
        uint32_t off, hi, b_addr, col, ks, acc, owner = 0;
        if (is_film) {
          const uint32_t f = ffirst ? i : i - nblk;
          off = (ffirst ? 0u : nblk * blk) + f * f_bytes;
          hi = f_hi; b_addr = u_base; col = T_FILM + f * 16; ks = f_swb >> 5; acc = 0;
        } else {
          const uint32_t bi = ffirst ? i - nfilm : i;
          off = (ffirst ? nfilm * f_bytes : 0u) + bi * blk;
          uint32_t t, tap, kb;
          if (job.mtiles >= 2) { t = bi % job.mtiles; kb = (bi / job.mtiles) % nkb; tap = bi / (job.mtiles * nkb); }
          else { t = 0; kb = bi % nkb; tap = bi / nkb; }
          const uint32_t tsel = job.taps == 3 ? tap : 1u;
          b_addr = (L == 4) ? b_base + tsel * (T::HALO * 128) + kb * T::SLAB : b_base + tsel * (BSLABS * T::SLAB) + kb * T::SLAB;
          hi = a_hi; col = T_ACC + t * NCOL; ks = a_swb >> 5; acc = (tap | kb) != 0 ? 1u : 0u;
          owner = (job.mtiles >= 2 && (t & 1u)) ? 1u : 0u;      // issued by the second issuer warp
        }
        const uint32_t stage = (chunk_base + off / CHUNK) % STAGES;
        const uint32_t a_addr = ring_a + stage * CHUNK + (off % CHUNK);
        const uint32_t w = (col + set * T_SET) | (ks << 9) | (acc << 12) | ((off % CHUNK == 0 ? 1u : 0u) << 13) |
                           ((is_film ? 1u : 0u) << 14) | ((off == 0 ? 1u : 0u) << 15) | (stage << 16) | (owner << 20);
        ops[op0 + i] = make_uint4(0x10000u | (a_addr >> 4), 0x10000u | (b_addr >> 4), hi, w);
      }
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cps = op_begin[n_jobs * NSETS + 1];
  const int wid_u = __shfl_sync(0xffffffffu, wid, 0);
  if (wid_u >= 8) {
  // register re-distribution (per 4-warp group): the service warps need few registers, the epilogue warps many
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (wid_u == 8) {
    // =========================== weight ring producer (dedicated warp) ===========================
    uint32_t used = 0, par = 0;
#pragma unroll 1
    for (int step = 0; step < n_steps; ++step)
#pragma unroll 1
      for (uint32_t ci = 0; ci < cps; ++ci) {
        const uint32_t s = ci % STAGES;
        if ((used >> s) & 1u) mbar_wait(&empty[s], ((par >> s) & 1u) ^ 1u);
        used |= 1u << s;
        par ^= 1u << s;
        const uint2 e = chunk_tab[ci];
        bulk_g2s_elect(smem + T::SM_RING + s * CHUNK, p.pack + e.x, e.y, &full[s]);
      }
  } else if (wid_u == 9 || wid_u == 10) {
    // =========================== UMMA issuers (dedicated warps; even / odd output tiles) ===========================
    const uint32_t my_owner = wid_u == 10 ? 1u : 0u;
    const uint32_t idesc64 = idesc_bf16(128, NCOL), idesc16 = idesc_bf16(128, 16);
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
    uint32_t full_par = 0, jobn = 0;
#pragma unroll 1
    for (int step = 0; step < n_steps; ++step)
#pragma unroll 1
      for (int j = 0; j < n_jobs; ++j, ++jobn)
#pragma unroll 1
        for (int set = 0; set < NSETS; ++set) {
          const uint32_t o0 = op_begin[j * NSETS + set], o1 = op_begin[j * NSETS + set + 1];
          const bool rec = p.prof && blockIdx.x == 0 && step == 1 && lane == 0 && set == 0;
          long long* pr = p.prof + 8 * j;
          if (rec) pr[2] = clock64();
          mbar_wait(&b_ready[set], jobn & 1);
          tc_fence_after();
          if (rec) pr[3] = clock64();
          uint32_t prev_stage = 0;
#pragma unroll 1
          for (uint32_t i = o0; i < o1; ++i) {
            uint4 op = ops[i];
            op.x = __shfl_sync(0xffffffffu, op.x, 0); op.y = __shfl_sync(0xffffffffu, op.y, 0);
            op.z = __shfl_sync(0xffffffffu, op.z, 0); op.w = __shfl_sync(0xffffffffu, op.w, 0);
            const uint32_t stage = (op.w >> 16) & 7u;
            if (op.w & (1u << 13)) {                       // first block of a weight chunk
              if (!(op.w & (1u << 15))) umma_commit_elect(&empty[prev_stage]);   // previous chunk free once its UMMAs retire
              mbar_wait(&full[stage], (full_par >> stage) & 1u);
              full_par ^= 1u << stage;
              tc_fence_after();
              prev_stage = stage;
            }
            if (op.w & (1u << 19)) continue;               // ring padding entry
            if (((op.w >> 20) & 1u) != my_owner) continue;  // the other issuer's tile
            const uint64_t ad = ((uint64_t)op.z << 32) | op.x;
            const uint64_t bd = ((uint64_t)b_hi << 32) | op.y;
            const uint32_t d = tmem_base + (op.w & 0x1FFu), acc = (op.w >> 12) & 1u, ks = (op.w >> 9) & 7u;
            const uint32_t idesc = (op.w & (1u << 14)) ? idesc16 : idesc64;
            if (ks == 4) umma_bf16_block_elect<4>(d, ad, bd, idesc, acc);
            else if (ks == 2) umma_bf16_block_elect<2>(d, ad, bd, idesc, acc);
            else umma_bf16_block_elect<1>(d, ad, bd, idesc, acc);
          }
          if (rec) pr[4] = clock64();
          umma_commit_elect(&empty[prev_stage]);
          umma_commit_elect(&acc_ready[set]);
          if (rec) pr[5] = clock64();
        }
  }
  } else {
    // =========================== epilogue warps ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    Ctx c;
    c.lane = lane; c.q = wid & 3; c.g = wid >> 2; c.ch = c.q * 32 + lane;
    c.smem = smem;
    c.scr = reinterpret_cast<float*>(smem + T::SM_SCR) + wid * T::SCR;
    c.xch = reinterpret_cast<float*>(smem + T::SM_XCH) + c.g * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) c.xoff[j] = ((uint32_t)((((c.ch & 63) >> 3) ^ j) << 4)) + (c.ch & 7) * 2;
    constexpr int NSW = NS / 2;          // samples per warp-group
    const int sgl = c.g * NSW;           // first sample of this warp-group inside a set
    uint32_t jobn = 0;                   // hand-offs per set so far (parity of b_ready / acc_ready)
    int prof_step = -1;

    // select the sample set the following code works on
    float* s_inemb = nullptr;
    float* s_x = nullptr;
    __nv_bfloat16* s_res1 = nullptr;
    int s0 = 0;
    auto select_set = [&](int set) {
      c.tmem = tmem_base + (uint32_t)set * T_SET + ((uint32_t)(c.q * 32) << 16);
      c.bbase = smem + T::SM_B + set * T::B_BYTES;
      s_inemb = reinterpret_cast<float*>(smem + T::SM_INEMB + set * 3072);
      s_x = reinterpret_cast<float*>(smem + T::SM_X + set * 256);
      s_res1 = reinterpret_cast<__nv_bfloat16*>(smem + T::SM_RES1 + set * 16384);
      s0 = cta_s0 + set * NS;
    };
    auto handoff = [&](int set) {        // operand of the next job of `set` is complete: one arrival per warp
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_ready[set]);
    };

#pragma unroll 1
    for (int step = 0; step < n_steps; ++step) {
      prof_step = step;
      if (p.prof && blockIdx.x == 0 && tid == 0 && (step == 1 || step == 2)) p.prof[320 + step - 1] = clock64();
#pragma unroll 1
      for (int set = 0; set < NSETS; ++set) {
        select_set(set);
        // ---- u[s][e] = sum_r silu(time_emb[e] + in_emb[s][r][e])  -> FiLM GEMM operand (bf16), one (s, e) per thread
        {
          const int s = tid / EMB, e = tid % EMB;
          float te = 0.f;
          if (cfg.time_cond) {
            const int ti = (p.mode == 0) ? step : min(s0 + s, p.n - 1);
            te = __ldg(p.te + (size_t)ti * EMB + e);
            if (p.cls_emb) te += __ldg(p.cls_emb + (size_t)(min(s0 + s, p.n - 1) / p.gpo) * EMB + e);
          }
          float a = 0.f;
          for (int r = 0; r < R; ++r) { const float z = te + s_inemb[(s * 3 + r) * EMB + e]; a += z / (1.0f + __expf(-z)); }
          *reinterpret_cast<__nv_bfloat16*>(smem + T::SM_U + set * 2048 + swz_off<128>(s, e >> 3) + (e & 7) * 2) =
              __float2bfloat16(a);
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
            if (L == 4) {
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
            } else {
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                float xs[16];
#pragma unroll
                for (int l = 0; l < 16; ++l) xs[l] = s_x[(sgl + jj) * 16 + l];
#pragma unroll
                for (int l = 0; l < 16; ++l) {
                  float a = b;
#pragma unroll
                  for (int t = 0; t < 7; ++t) {
                    const int ll = l + t - 3;
                    if (ll >= 0 && ll < 16) a = fmaf(w7[t], xs[ll], a);
                  }
                  v[jj * 16 + l] = a;
                }
              }
            }
          }
          tm_store32<L>(c, T_RES, v);
          write_b<L>(c, 0, v, c.ch < c0);
        }
        handoff(set);
      }

#pragma unroll 1
      for (int j = 0; j < n_jobs; ++j, ++jobn) {
        const TcJob job = s_jobs[j];
        // ---- per-channel parameters of this job (shared by the sample sets), fetched while the UMMAs run
        const int flags = job.flags, ch = job.ch, nt = job.mtiles;
        float pbias[2], pgam[2], pbet[2], pcs[2], pch[2], pg[2], pg2[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int chan = t * 128 + c.ch;
          const bool valid = t < nt && chan < ch && !(flags & E_ATTN);
          pbias[t] = (valid && job.o_bias >= 0) ? __ldg(W + job.o_bias + chan) : 0.f;
          pgam[t] = (valid && job.o_gamma >= 0) ? __ldg(W + job.o_gamma + chan) : 0.f;
          pbet[t] = (valid && job.o_beta >= 0) ? __ldg(W + job.o_beta + chan) : 0.f;
          pcs[t] = (valid && job.o_mlpb >= 0) ? (float)R * __ldg(W + job.o_mlpb + chan) + (float)R : 0.f;
          pch[t] = (valid && job.o_mlpb >= 0) ? (float)R * __ldg(W + job.o_mlpb + ch + chan) : 0.f;
          pg[t] = (valid && job.o_g >= 0) ? __ldg(W + job.o_g + chan) : 0.f;
          pg2[t] = (valid && job.o_g2 >= 0) ? __ldg(W + job.o_g2 + chan) : 0.f;
        }
#pragma unroll 1
        for (int set = 0; set < NSETS; ++set) {
          select_set(set);
          // ---- wait for the accumulator of (job, set); the tensor core meanwhile works on the other set
          {
            const bool rec = p.prof && blockIdx.x == 0 && tid == 0 && prof_step == 1 && set == 0;
            const long long t0 = rec ? clock64() : 0;
            mbar_wait(&acc_ready[set], jobn & 1);
            if (rec) { p.prof[8 * j] = t0; p.prof[8 * j + 1] = clock64(); }
            tc_fence_after();
          }
          if (flags & E_ATTN) {
            // ======== linear attention (resnets.py:211-235): qkv -> core -> operand of to_out.  Warp = head, lane = d.
            float kk[32], e[32];
            tm_load32<L>(c, T_ACC + 1 * NCOL, kk);
            if (L == 4) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {   // softmax over the 4 positions (dim=-1)
                const float m = fmaxf(fmaxf(kk[jj], kk[8 + jj]), fmaxf(kk[16 + jj], kk[24 + jj]));
                float sum = 0.f;
#pragma unroll
                for (int l = 0; l < 4; ++l) { kk[l * 8 + jj] = __expf(kk[l * 8 + jj] - m); sum += kk[l * 8 + jj]; }
                const float inv = __fdividef(1.0f, sum);
#pragma unroll
                for (int l = 0; l < 4; ++l) kk[l * 8 + jj] *= inv;
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {   // softmax over the 16 positions
                float m = kk[jj * 16];
#pragma unroll
                for (int l = 1; l < 16; ++l) m = fmaxf(m, kk[jj * 16 + l]);
                float sum = 0.f;
#pragma unroll
                for (int l = 0; l < 16; ++l) { kk[jj * 16 + l] = __expf(kk[jj * 16 + l] - m); sum += kk[jj * 16 + l]; }
                const float inv = __fdividef(1.0f, sum);
#pragma unroll
                for (int l = 0; l < 16; ++l) kk[jj * 16 + l] *= inv;
              }
            }
            tm_load32<L>(c, T_ACC + 0 * NCOL, e);
#pragma unroll
            for (int i = 0; i < 32; ++i) e[i] = __expf(fminf(e[i], 80.f));   // softmax over d: normalised by Z below
            // lane sums over d: A[s][n'][n] = sum_d k[n'][s] e[n][s], Z[n][s] = sum_d e[n][s]
            constexpr int ZOFF = (L == 4) ? 128 : 512;
            if (L == 4) {
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
            } else {
#pragma unroll 1
              for (int b = 0; b < 16; ++b) {        // sample b >> 3, rows n' = 2 (b & 7) + {0, 1}, all 16 columns n
                const int jj = b >> 3, n1 = 2 * (b & 7);
                float k0 = 0.f, k1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {      // select k[jj][n1], k[jj][n1 + 1] without dynamic register indexing
                  k0 = (i == jj * 16 + n1) ? kk[i] : k0;
                  k1 = (i == jj * 16 + n1 + 1) ? kk[i] : k1;
                }
                float pr[32];
#pragma unroll
                for (int n = 0; n < 16; ++n) {
                  const float en = jj ? e[16 + n] : e[n];
                  pr[n] = k0 * en;
                  pr[16 + n] = k1 * en;
                }
                c.scr[b * 32 + lane] = reduce_scatter32(pr, lane);    // A[jj][n1 + (lane >> 4)][lane & 15]
              }
            }
            {
              float z[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) z[i] = e[i];
              c.scr[ZOFF + lane] = reduce_scatter32(z, lane);
            }
            __syncwarp();
            float vv[32], o[32];
            tm_load32<L>(c, T_ACC + 2 * NCOL, vv);
            if (L == 4) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                float A[16];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 t4 = *reinterpret_cast<const float4*>(c.scr + (jj >> 1) * 32 + (jj & 1) * 16 + i);
                  A[i] = t4.x; A[i + 1] = t4.y; A[i + 2] = t4.z; A[i + 3] = t4.w;
                }
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                  float acc = 0.f;
#pragma unroll
                  for (int n1 = 0; n1 < 4; ++n1) acc = fmaf(vv[n1 * 8 + jj], A[n1 * 4 + n], acc);
                  const float zinv = __fdividef(0.17677669529663687f, c.scr[ZOFF + n * 8 + jj]);   // scale 32^-0.5 / Z
                  o[n * 8 + jj] = acc * zinv;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = 0.f;
#pragma unroll
              for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                  const float vn = vv[jj * 16 + n1];
                  const float* Ar = c.scr + (jj * 8 + (n1 >> 1)) * 32 + (n1 & 1) * 16;     // A[jj][n1][0..15]
#pragma unroll
                  for (int n = 0; n < 16; n += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(Ar + n);
                    o[jj * 16 + n] = fmaf(vn, t4.x, o[jj * 16 + n]);
                    o[jj * 16 + n + 1] = fmaf(vn, t4.y, o[jj * 16 + n + 1]);
                    o[jj * 16 + n + 2] = fmaf(vn, t4.z, o[jj * 16 + n + 2]);
                    o[jj * 16 + n + 3] = fmaf(vn, t4.w, o[jj * 16 + n + 3]);
                  }
                }
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] *= __fdividef(0.17677669529663687f, c.scr[ZOFF + i]);
            }
            __syncwarp();
            write_b<L>(c, 0, o, true);
      } else {
            // ======== generic epilogue: bias, GroupNorm / LayerNorm, FiLM, SiLU, residual, PreNorm, operand write
            float fc_part[32];
            if (flags & E_FINAL) {
#pragma unroll
              for (int i = 0; i < 32; ++i) fc_part[i] = 0.f;
            }
#pragma unroll 1
            for (int t = 0; t < nt; ++t) {
              const bool valid = t * 128 + c.ch < ch;
              // a warp whose 32 channels all lie beyond the layer width has nothing to contribute unless the epilogue
              // has a warp-group wide reduction (LayerNorm, final conv)
              if (t * 128 + c.q * 32 >= ch && !(flags & (E_LN | E_LNNEXT | E_FINAL))) continue;
              const float bias = t ? pbias[1] : pbias[0], gam = t ? pgam[1] : pgam[0], bet = t ? pbet[1] : pbet[0];
              const float cs = t ? pcs[1] : pcs[0], chh = t ? pch[1] : pch[0], g1 = t ? pg[1] : pg[0], g2 = t ? pg2[1] : pg2[0];
              uint32_t rv[32], rr[32], rs8[8], rh8[8];
              tm_issue32<L>(c, T_ACC + t * NCOL, rv);
              if ((flags & E_ADDRES) && t == 0) tm_issue32<L>(c, T_RES, rr);
              if (flags & E_FILM) {      // FiLM tile columns = samples: 8 per warp-group (L = 4) or all 4 of the CTA (L = 16)
                tm_issue8(c, T_FILM + t * 16 + (L == 4 ? c.g * 8 : 0), rs8);
                tm_issue8(c, T_FILM + (nt + t) * 16 + (L == 4 ? c.g * 8 : 0), rh8);
              }
              tm_wait();
              float v[32];
              tm_use(rv, v);
              {
                const float2 b2 = bc2(bias);
#pragma unroll
                for (int i = 0; i < 16; ++i) st2(v, i, __fadd2_rn(ld2(v, i), b2));
              }
              if (flags & E_GN) {
                float mean[8], rstd[8];
                gn_stats_dispatch<L>(c, ch, v, mean, rstd);
                float fs[8], fh[8];
                if (flags & E_FILM) {
                  tm_use(rs8, fs);
                  tm_use(rh8, fh);
                  if (L == 16) {      // this warp-group's two samples are columns 2g, 2g + 1
                    fs[0] = c.g ? fs[2] : fs[0]; fs[1] = c.g ? fs[3] : fs[1];
                    fh[0] = c.g ? fh[2] : fh[0]; fh[1] = c.g ? fh[3] : fh[1];
                  }
                }
                // GroupNorm affine and FiLM folded into one multiply-add per value: v * A + B
                float A[NSW], Bc[NSW];
#pragma unroll
                for (int jj = 0; jj < NSW; ++jj) {
                  const float a = rstd[jj] * gam, b = fmaf(-mean[jj], a, bet);
                  const float sc = (flags & E_FILM) ? fs[jj] + cs : 1.f, sh = (flags & E_FILM) ? fh[jj] + chh : 0.f;
                  A[jj] = a * sc;
                  Bc[jj] = fmaf(b, sc, sh);
                }
                if (L == 4) {
#pragma unroll
                  for (int jp = 0; jp < 4; ++jp) {
                    const float2 A2 = make_float2(A[2 * jp], A[2 * jp + 1]), B2 = make_float2(Bc[2 * jp], Bc[2 * jp + 1]);
#pragma unroll
                    for (int l = 0; l < 4; ++l) st2(v, l * 4 + jp, __ffma2_rn(ld2(v, l * 4 + jp), A2, B2));
                  }
                } else {
#pragma unroll
                  for (int jj = 0; jj < 2; ++jj) {
                    const float2 A2 = bc2(A[jj]), B2 = bc2(Bc[jj]);
#pragma unroll
                    for (int lp = 0; lp < 8; ++lp) st2(v, jj * 8 + lp, __ffma2_rn(ld2(v, jj * 8 + lp), A2, B2));
                  }
                }
              }
              if (flags & E_SILU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) st2(v, i, silu_fast2(ld2(v, i)));
              }
              if (flags & E_LN) ln_apply(c, ch, g1, v);
              if (flags & E_ADDRES) {
                if (t == 0) {
                  float res[32];
                  tm_use(rr, res);
#pragma unroll
                  for (int i = 0; i < 16; ++i) st2(v, i, __fadd2_rn(ld2(v, i), ld2(res, i)));
                } else {               // channels 128..255 (final block only): bf16 copy in shared memory
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] += __bfloat162float(s_res1[i * 256 + tid]);
                }
              }
              if (!valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
              }
              if (flags & E_STORERES) {
                if (t == 0) {
                  tm_store32<L>(c, T_RES, v);
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) s_res1[i * 256 + tid] = __float2bfloat16(v[i]);
                }
              }
              if (flags & E_LNNEXT) ln_apply(c, ch, g2, v);
              if (flags & E_FINAL) {
#pragma unroll
                for (int i = 0; i < 16; ++i) st2(fc_part, i, __ffma2_rn(bc2(g2), ld2(v, i), ld2(fc_part, i)));
              } else {
                write_b<L>(c, t, v, valid);
              }
            }
            if (flags & E_FINAL) {
              // ======== final_conv (1x1 -> 1 channel) + scheduler update (L = 4) / trunk output (L = 16)
              const float part = reduce_scatter32(fc_part, lane);     // lane r: value r of this warp's 32 channels
              c.xch[c.q * 64 + lane] = part;
              wg_sync(c.g);
              if (c.q == 0) {
                float eps = __ldg(W + lay.fc_b);
#pragma unroll
                for (int w = 0; w < 4; ++w) eps += c.xch[w * 64 + lane];
                {
                  // value index of this lane: l*8 + jj (L = 4) or jj*16 + l (L = 16)
                  const int l = (L == 4) ? lane >> 3 : lane & 15, jj = (L == 4) ? lane & 7 : lane >> 4, s = sgl + jj;
                  const bool ok = s0 + s < p.n;
                  if (p.mode == 0) {
                    const float* cf = p.coef + (size_t)step * 8;
                    const float x = s_x[s * L + l];
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
                    s_x[s * L + l] = prev;
                    if (p.x_all && ok) p.x_all[((size_t)(step + 1) * p.n + s0 + s) * L + l] = prev;
                  } else {
                    s_x[s * L + l] = eps;       // network output: denoiser eps (mode 1) / decoder trunk features (mode 2)
                  }
                }
              }
              wg_sync(c.g);
            }
          }
          if (j + 1 < n_jobs) handoff(set);        // operand of job j+1 written: the issuers take over for this set
        }
      }
    }
    // ---- outputs
#pragma unroll 1
    for (int set = 0; set < NSETS; ++set) {
      select_set(set);
      wg_sync(c.g);
      if (p.mode != 2) {
        if (c.q == 0) {
          const int l = (L == 4) ? lane >> 3 : lane & 15, jj = (L == 4) ? lane & 7 : lane >> 4, s = sgl + jj;
          if (s0 + s < p.n) p.x_out[(size_t)(s0 + s) * L + l] = s_x[s * L + l];
        }
      } else {
        // decoder heads: tmrp = Linear(L -> 6), class_logits = Linear(L -> 1)   (grasp_vae.py:428-430)
        const float* hw = p.head + L * p.D + L;   // tmrp_w [6][L], tmrp_b [6], cls_w [L], cls_b [1]
        const int tl = tid & 127;
        if (tl < NSW * 7) {
          const int jj = tl / 7, o = tl - jj * 7, s = sgl + jj;
          if (s0 + s < p.n) {
            const float* x = s_x + s * 16;
            if (o < 6) {
              float a = __ldg(hw + 6 * L + o);
              for (int l = 0; l < L; ++l) a = fmaf(__ldg(hw + o * L + l), x[l], a);
              p.tmrp[(size_t)(s0 + s) * 6 + o] = a;
            } else {
              float a = __ldg(hw + 6 * L + 6 + L);
              for (int l = 0; l < L; ++l) a = fmaf(__ldg(hw + 6 * L + 6 + l), x[l], a);
              p.logit[s0 + s] = a;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem_base);
}

#include "sampler_rows.cuh"

// ------------------------------------------------------------------------------------------------
// FiLM table of the row-major kernel.  The conditioning embedding of a ResnetBlock depends only on (object, step):
//   latent_emb[r] = time_mlp(t) + SiLU(Linear(z_pc[obj][r]))                      resnets.py:590-603
//   scale | shift = sum_r Linear(SiLU(latent_emb[r])), x * (scale + R) + shift     resnets.py:163-175, 191-196
// so one vector per (object, step) serves all grasps of the object (the channel-major kernel rebuilds it per sample
// and step on the tensor cores).  Folded with the GroupNorm affine of the block: A = gamma * S, B = beta * S + H.
// One block per object (mode 1: per sample); thread = one FiLM channel, its two projection rows live in registers
// while the block walks the steps.
// ------------------------------------------------------------------------------------------------
struct FilmJobs { int n; int ch[12], o_mlpw[12], o_mlpb[12], o_gamma[12], o_beta[12], o_film[12]; };

template <int EMB>
__global__ void __launch_bounds__(256) film_table_kernel(const float* __restrict__ W, ResNetLayout lay, FilmJobs fj, int R,
                                                         int cond_dim, const float* __restrict__ z_cond, int z_div,
                                                         const float* __restrict__ te, int te_per_block, int n_steps,
                                                         int stride, const float* __restrict__ cls_emb,
                                                         float* __restrict__ film) {
  constexpr int TS = 32;                        // steps per tile
  __shared__ float s_ie[4 * EMB];
  __shared__ float s_u[TS][EMB];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int idx = tid; idx < R * EMB; idx += 256) {
    const int e = idx % EMB, r = idx / EMB;
    const float* z = z_cond + ((size_t)(b / z_div) * R + r) * cond_dim;
    const float* w = W + lay.in_w + (size_t)e * cond_dim;
    float a = __ldg(W + lay.in_b + e);
    for (int j = 0; j < cond_dim; ++j) a = fmaf(__ldg(w + j), __ldg(z + j), a);
    s_ie[idx] = a / (1.0f + expf(-a));
  }
  int total = 0;
  for (int j = 0; j < fj.n; ++j) total += fj.ch[j];
  for (int t0 = 0; t0 < n_steps; t0 += TS) {
    const int nt = min(TS, n_steps - t0);
    __syncthreads();
    for (int idx = tid; idx < nt * EMB; idx += 256) {
      const int e = idx % EMB, t = idx / EMB;
      float tv = te ? __ldg(te + (size_t)(te_per_block ? b : t0 + t) * EMB + e) : 0.f;      // no time embedding: ResNet1D
      if (cls_emb) tv += __ldg(cls_emb + (size_t)(b / z_div) * EMB + e);      // class_conditioned_resnet.py:96-98
      float a = 0.f;
      for (int r = 0; r < R; ++r) { const float zz = tv + s_ie[r * EMB + e]; a += zz / (1.0f + __expf(-zz)); }
      s_u[t][e] = a;
    }
    __syncthreads();
    for (int q = tid; q < total; q += 256) {
      int j = 0, c = q;
      while (c >= fj.ch[j]) { c -= fj.ch[j]; ++j; }
      const int ch = fj.ch[j];
      float ws[EMB], wh[EMB];
#pragma unroll
      for (int e = 0; e < EMB; ++e) {
        ws[e] = __ldg(W + fj.o_mlpw[j] + (size_t)c * EMB + e);
        wh[e] = __ldg(W + fj.o_mlpw[j] + (size_t)(ch + c) * EMB + e);
      }
      const float bs = (float)R * __ldg(W + fj.o_mlpb[j] + c) + (float)R, bh = (float)R * __ldg(W + fj.o_mlpb[j] + ch + c);
      const float gamma = __ldg(W + fj.o_gamma[j] + c), beta = __ldg(W + fj.o_beta[j] + c);
      float* dst = film + ((size_t)b * n_steps + t0) * stride + fj.o_film[j] + c;
      for (int t = 0; t < nt; ++t) {
        float S = bs, H = bh;
#pragma unroll
        for (int e = 0; e < EMB; ++e) { S = fmaf(ws[e], s_u[t][e], S); H = fmaf(wh[e], s_u[t][e], H); }
        dst[(size_t)t * stride] = gamma * S;
        dst[(size_t)t * stride + ch] = fmaf(beta, S, H);
      }
    }
  }
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
  p.n_jobs = build_jobs(*cfg, p.jobs, &total, &p.film_stride);
  return GLDM_OK;
}

static long long* g_tc_prof = nullptr;

template <int L, int NSETS>
static int launch_tc_l(TcParams& p, cudaStream_t s) {
  static SmemOptIn attr;
  using T = Tr<L, NSETS>;
  const int smem = T::SM_TOTAL + 1024;
  static_assert(T::SM_TOTAL + 1024 <= 232448, "shared memory budget");
  if (int rc = opt_in_smem(attr, resnet_tc_kernel<L, NSETS>, smem, "resnet_tc_kernel")) return rc;
  resnet_tc_kernel<L, NSETS><<<ceil_div(p.n, T::NS * NSETS), stc::NTHREADS, smem, s>>>(p);
  return check_launch(L == 4 ? "resnet_tc_kernel<4>" : "resnet_tc_kernel<16>");
}

// sample sets per CTA: 0 = automatic.  Two sets (32 samples per CTA, the UMMA phase of one set overlaps the epilogue
// of the other) pay off once the batch no longer fits one wave of single-set CTAs.
static int g_tc_sets = 0;

// row-major kernel (sampler_rows.cuh): 1 = on where the configuration allows it, 0 = channel-major kernel, -1 (default)
// = by batch size: the row-major kernel has the better throughput per SM (32 samples per CTA), the channel-major one
// spreads a small batch over twice as many SMs (16 samples per CTA) and finishes it sooner
static int g_tc_rows = -1;

// fpc latent denoiser (L = 4, dim 4, emb 16) or the L = 16 family (ppc latent denoiser, grasp decoder trunk: dim 16, emb 64)
static bool rows_supported(const GldmResNetCfg& c) {
  const bool l4 = c.L == 4 && c.emb_dim == 16 && c.time_cond && c.ch[0] == 4;
  const bool l16 = c.L == 16 && c.emb_dim == 64 && c.ch[0] == 16;
  return (l4 || l16) && c.n_stages == 4 && c.groups == 4 && c.ch[1] == 32 && c.ch[2] == 64 && c.ch[3] == 128 && c.ch[4] == 256;
}

static const bool g_rows_persistent_decoder = !(getenv("GLDM_DECODER_PERSISTENT") && atoi(getenv("GLDM_DECODER_PERSISTENT")) == 0);
template <int L>
static int launch_rows_l(TcParams& p, cudaStream_t s) {
  static SmemOptIn attr;
  constexpr int EMB = (L == 4) ? 16 : 64;
  const int smem = rows::SM_TOTAL + 1024;
  if (int rc = opt_in_smem(attr, rows::resnet_rows_kernel<L>, smem, "resnet_rows_kernel")) return rc;
  // FiLM table: one row per (object, step) in the sampler, per sample in a single evaluation with its own time, per object
  // in the decoder (no time embedding); stream-ordered scratch
  FilmJobs fj = {};
  for (int j = 0; j < p.n_jobs; ++j)
    if (p.jobs[j].o_film >= 0) {
      GLDM_REQUIRE(fj.n < 12, "resnet_rows: too many FiLM layers");
      fj.ch[fj.n] = p.jobs[j].ch; fj.o_mlpw[fj.n] = p.jobs[j].o_mlpw; fj.o_mlpb[fj.n] = p.jobs[j].o_mlpb;
      fj.o_gamma[fj.n] = p.jobs[j].o_gamma; fj.o_beta[fj.n] = p.jobs[j].o_beta; fj.o_film[fj.n] = p.jobs[j].o_film;
      ++fj.n;
    }
  const int blocks = p.mode == 1 ? p.n : ceil_div(p.n, p.gpo), steps = p.mode == 0 ? p.n_steps : 1;
  float* film = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&film), sizeof(float) * (size_t)blocks * steps * p.film_stride, s) != cudaSuccess) {
    set_error("resnet_rows: cudaMallocAsync of the FiLM table (%lld bytes) failed",
              (long long)(sizeof(float) * (size_t)blocks * steps * p.film_stride));
    return GLDM_ECUDA;
  }
  film_table_kernel<EMB><<<blocks, 256, 0, s>>>(p.W, p.lay, fj, p.cfg.cond_ch, p.cfg.cond_dim, p.z_cond, p.mode == 1 ? p.gpo : 1,
                                                p.cfg.time_cond ? p.te : nullptr, p.mode == 1 ? 1 : 0, steps, p.film_stride,
                                                p.cls_emb, film);
  int rc = check_launch("film_table_kernel");
  if (rc == GLDM_OK) {
    p.film = film;
    // decoder: persistent CTAs, every one the same number of sample groups (+- 1); sampler / single evaluation: one group
    int grid = ceil_div(p.n, rows::Geo<L>::NS);
    if (p.mode == 2 && g_rows_persistent_decoder) grid = ceil_div(grid, ceil_div(grid, kNumSMs));
    rows::resnet_rows_kernel<L><<<grid, rows::NTHREADS, smem, s>>>(p);
    rc = check_launch("resnet_rows_kernel");
  }
  cudaFreeAsync(film, s);
  return rc;
}

static int launch_rows(TcParams& p, cudaStream_t s) {
  return p.cfg.L == 4 ? launch_rows_l<4>(p, s) : launch_rows_l<16>(p, s);
}

// L = 16 networks (grasp decoder trunk, ppc latent denoiser) on the row-major kernel: 8 samples per CTA instead of the
// channel-major kernel's 4, at any batch size.  GLDM_TC_ROWS16=0 selects the channel-major kernel.
static bool rows16_enabled() {
  static int on = -1;
  if (on < 0) { const char* ev = getenv("GLDM_TC_ROWS16"); on = ev ? atoi(ev) : 1; }
  return on != 0;
}

static int launch_tc(TcParams& p, cudaStream_t s) {
  p.prof = g_tc_prof;
  if (p.mode == 0 && p.sched_kind == GLDM_SCHED_EDM) {
    // evaluation programs (elucidated samplers) are implemented by the row-major kernel
    if (!(rows_supported(p.cfg) && p.cfg.L == 4)) {
      set_error("sampler_tc: the elucidated samplers run on the row-major kernel (fpc latent denoiser); use precision fp32 for this model");
      return GLDM_ENOSUP;
    }
    return launch_rows(p, s);
  }
  if (p.cfg.L != 4) {
    if (g_tc_rows != 0 && rows16_enabled() && rows_supported(p.cfg)) return launch_rows(p, s);
    return launch_tc_l<16, 1>(p, s);
  }
  const bool want_rows = g_tc_rows == 1 || (g_tc_rows < 0 && p.n > 16 * kNumSMs);
  if (want_rows && rows_supported(p.cfg)) return launch_rows(p, s);
  const bool two = g_tc_sets == 2 || (g_tc_sets == 0 && p.n > 16 * kNumSMs);
  return two ? launch_tc_l<4, 2>(p, s) : launch_tc_l<4, 1>(p, s);
}

}  // namespace gldm

using namespace gldm;

extern "C" int gldm_sampler_tc_set_profile(long long* dev_buf) {
  g_tc_prof = dev_buf;
  return GLDM_OK;
}

extern "C" int gldm_sampler_tc_set_sets(int sets) {
  if (sets < 0 || sets > 2) {
    set_error("sampler_tc_set_sets: 0 (automatic), 1 or 2");
    return GLDM_EINVAL;
  }
  g_tc_sets = sets;
  return GLDM_OK;
}

extern "C" int gldm_sampler_tc_set_rows(int on) {
  g_tc_rows = on < 0 ? -1 : on ? 1 : 0;
  return GLDM_OK;
}

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
  const int emb = cfg->emb_dim;
  const uint32_t ftile = film_tile_bytes(emb);
  auto pack_main = [&](const TcJob& j, int src_off, int cout, int cin, int standardize) {
    uint8_t* m = dst + j.a_off + (film_first(j, emb) ? (size_t)j.film_tiles * ftile : 0);
    pack_image_kernel<<<ceil_div(j.mtiles * 128, 8), 256, 0, s>>>(raw + src_off, m, cout, 0, cin, j.taps, j.mtiles, j.kpt,
                                                                  j.a_swb, standardize, j.mtiles >= 2);
    ++launches;
  };
  auto pack_film = [&](const TcJob& j, int mlp_off, int ch) {
    uint8_t* f = dst + j.a_off + (film_first(j, emb) ? 0 : main_bytes(j));
    const int ct = (ch + 127) / 128;
    for (int half = 0; half < 2; ++half)
      for (int t = 0; t < ct; ++t) {
        const int row0 = half * ch + t * 128;
        pack_image_kernel<<<ceil_div(128, 8), 256, 0, s>>>(raw + mlp_off, f + (size_t)(half * ct + t) * ftile,
                                                           min(128, ch - t * 128), row0, emb, 1, 1, pad16(emb),
                                                           swb_for(pad16(emb)), 0, 0);
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

static int run_time_embed(const TcParams& p, const int* ts_dev, int count, float* te, cudaStream_t s,
                          const float* tsf_dev = nullptr) {
  time_embed_kernel<<<ceil_div(count, 8), 256, 0, s>>>(p.W, p.lay, p.cfg.fourier_half, p.cfg.emb_dim, ts_dev, tsf_dev, count, te);
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
  GLDM_REQUIRE(cfg->time_cond, "sampler_run_tc: the sampler needs a time-conditioned denoiser configuration");
  GLDM_REQUIRE(n <= 0 || (x_T && z_obj && x_out && timesteps_host && coef_host), "sampler_run_tc: null pointer");
  GLDM_REQUIRE(n >= 0 && grasps_per_obj > 0 && n_steps > 0, "sampler_run_tc: bad sizes");
  GLDM_REQUIRE(sched_kind == GLDM_SCHED_DDPM || sched_kind == GLDM_SCHED_DDIM, "sampler_run_tc: bad scheduler");
  if (n == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  void* scratch = nullptr;
  const size_t cb = sizeof(float) * 8 * (size_t)n_steps, tb = sizeof(int) * (size_t)n_steps,
               eb = sizeof(float) * p.cfg.emb_dim * (size_t)n_steps;
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

/* device-table variant: nothing is copied or allocated per call (pageable H2D copies synchronise the stream) */
extern "C" int gldm_time_embed_table(const GldmResNetCfg* cfg, const float* raw, const int* timesteps_dev, int count,
                                     float* te_dev, void* stream) {
  int rc = check_tc_cfg(cfg);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->time_cond && raw && timesteps_dev && te_dev && count > 0, "time_embed_table: bad arguments");
  TcParams p = {};
  p.cfg = *cfg;
  make_layout(*cfg, p.lay);
  p.W = raw;
  return run_time_embed(p, timesteps_dev, count, te_dev, (cudaStream_t)stream);
}

extern "C" int gldm_sampler_run_tc_dev(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x_T,
                                       const float* z_obj, int n, int grasps_per_obj, int n_steps, const float* coef_dev,
                                       const float* te_dev, int sched_kind, int clip_sample, const float* noise,
                                       unsigned long long seed, float* x_out, float* x_all, void* stream) {
  TcParams p = {};
  int rc = fill_tc(p, cfg, raw, pack);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->time_cond, "sampler_run_tc_dev: the sampler needs a time-conditioned denoiser configuration");
  GLDM_REQUIRE(n <= 0 || (x_T && z_obj && x_out && coef_dev && te_dev), "sampler_run_tc_dev: null pointer");
  GLDM_REQUIRE(n >= 0 && grasps_per_obj > 0 && n_steps > 0, "sampler_run_tc_dev: bad sizes");
  GLDM_REQUIRE(sched_kind == GLDM_SCHED_DDPM || sched_kind == GLDM_SCHED_DDIM, "sampler_run_tc_dev: bad scheduler");
  if (n == 0) return GLDM_OK;
  p.mode = 0; p.n = n; p.gpo = grasps_per_obj; p.x_in = x_T; p.z_cond = z_obj; p.te = te_dev;
  p.n_steps = n_steps; p.coef = coef_dev; p.sched_kind = sched_kind; p.clip = clip_sample;
  p.noise = noise; p.seed = seed; p.x_out = x_out; p.x_all = x_all;
  return launch_tc(p, (cudaStream_t)stream);
}

extern "C" int gldm_time_embed_table_f(const GldmResNetCfg* cfg, const float* raw, const float* times_dev, int count,
                                       float* te_dev, void* stream) {
  int rc = check_tc_cfg(cfg);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->time_cond && raw && times_dev && te_dev && count > 0, "time_embed_table_f: bad arguments");
  TcParams p = {};
  p.cfg = *cfg;
  make_layout(*cfg, p.lay);
  p.W = raw;
  return run_time_embed(p, nullptr, count, te_dev, (cudaStream_t)stream, times_dev);
}

extern "C" int gldm_sampler_run_ex_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const GldmSamplerArgs* a,
                                      void* stream) {
  TcParams p = {};
  int rc = fill_tc(p, cfg, raw, pack);
  if (rc) return rc;
  GLDM_REQUIRE(a, "sampler_run_ex_tc: null arguments");
  GLDM_REQUIRE(cfg->time_cond, "sampler_run_ex_tc: the sampler needs a time-conditioned denoiser configuration");
  GLDM_REQUIRE(a->n >= 0 && a->grasps_per_obj > 0 && a->n_steps > 0, "sampler_run_ex_tc: bad sizes");
  if (a->n == 0) return GLDM_OK;
  GLDM_REQUIRE(a->x_init && a->z_obj && a->x_out && a->coef && a->te, "sampler_run_ex_tc: null pointer");
  GLDM_REQUIRE(a->sched_kind == GLDM_SCHED_DDPM || a->sched_kind == GLDM_SCHED_DDIM || a->sched_kind == GLDM_SCHED_EDM,
               "sampler_run_ex_tc: bad scheduler");
  p.mode = 0; p.n = a->n; p.gpo = a->grasps_per_obj; p.x_in = a->x_init; p.z_cond = a->z_obj; p.te = a->te;
  p.n_steps = a->n_steps; p.coef = a->coef; p.sched_kind = a->sched_kind; p.clip = a->clip_sample;
  p.noise = a->noise; p.seed = a->seed; p.cls_emb = a->cls_emb; p.x_out = a->x_out; p.x_all = a->x_all;
  return launch_tc(p, (cudaStream_t)stream);
}

extern "C" int gldm_denoiser_forward_ex_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x,
                                           const int* t, const float* tf, const float* z_cond, const float* cls_emb, int n,
                                           float* eps, void* stream) {
  TcParams p = {};
  int rc = fill_tc(p, cfg, raw, pack);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->time_cond, "denoiser_forward_ex_tc: needs a time-conditioned denoiser configuration");
  GLDM_REQUIRE(n >= 0, "denoiser_forward_ex_tc: bad n");
  if (n == 0) return GLDM_OK;
  GLDM_REQUIRE(x && (t || tf) && z_cond && eps, "denoiser_forward_ex_tc: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  float* d_te = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&d_te), sizeof(float) * p.cfg.emb_dim * (size_t)n, s) != cudaSuccess) {
    set_error("denoiser_forward_ex_tc: cudaMallocAsync failed");
    return GLDM_ECUDA;
  }
  p.mode = 1; p.n = n; p.gpo = 1; p.x_in = x; p.z_cond = z_cond; p.te = d_te; p.n_steps = 1; p.x_out = eps; p.cls_emb = cls_emb;
  rc = run_time_embed(p, t, n, d_te, s, tf);
  if (rc == GLDM_OK) rc = launch_tc(p, s);
  cudaFreeAsync(d_te, s);
  return rc;
}

extern "C" int gldm_denoiser_forward_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x,
                                        const int* t, const float* z_cond, int n, float* eps, void* stream) {
  TcParams p = {};
  int rc = fill_tc(p, cfg, raw, pack);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->time_cond, "denoiser_forward_tc: needs a time-conditioned denoiser configuration");
  GLDM_REQUIRE(n <= 0 || (x && t && z_cond && eps), "denoiser_forward_tc: null pointer");
  GLDM_REQUIRE(n >= 0, "denoiser_forward_tc: bad n");
  if (n == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* d_te = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&d_te), sizeof(float) * p.cfg.emb_dim * (size_t)n, s) != cudaSuccess) {
    set_error("denoiser_forward_tc: cudaMallocAsync failed");
    return GLDM_ECUDA;
  }
  p.mode = 1; p.n = n; p.gpo = 1; p.x_in = x; p.z_cond = z_cond; p.te = d_te; p.n_steps = 1; p.x_out = eps;
  rc = run_time_embed(p, t, n, d_te, s);
  if (rc == GLDM_OK) rc = launch_tc(p, s);
  cudaFreeAsync(d_te, s);
  return rc;
}

extern "C" int gldm_denoiser_forward_tc_ftime(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x,
                                              const float* t, const float* z_cond, int n, float* eps, void* stream) {
  TcParams p = {};
  int rc = fill_tc(p, cfg, raw, pack);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->time_cond, "denoiser_forward_tc_ftime: needs a time-conditioned denoiser configuration");
  GLDM_REQUIRE(n <= 0 || (x && t && z_cond && eps), "denoiser_forward_tc_ftime: null pointer");
  GLDM_REQUIRE(n >= 0, "denoiser_forward_tc_ftime: bad n");
  if (n == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* d_te = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&d_te), sizeof(float) * p.cfg.emb_dim * (size_t)n, s) != cudaSuccess) {
    set_error("denoiser_forward_tc_ftime: cudaMallocAsync failed");
    return GLDM_ECUDA;
  }
  p.mode = 1; p.n = n; p.gpo = 1; p.x_in = x; p.z_cond = z_cond; p.te = d_te; p.n_steps = 1; p.x_out = eps;
  rc = run_time_embed(p, nullptr, n, d_te, s, t);
  if (rc == GLDM_OK) rc = launch_tc(p, s);
  cudaFreeAsync(d_te, s);
  return rc;
}

extern "C" int gldm_decoder_forward_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* head,
                                       int D, const float* z_h, const float* z_obj, int n, int grasps_per_obj,
                                       float* tmrp, float* logit, void* stream) {
  TcParams p = {};
  int rc = fill_tc(p, cfg, raw, pack);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->L == 16 && !cfg->time_cond, "decoder_forward_tc: needs the decoder trunk configuration (L = 16)");
  GLDM_REQUIRE(n <= 0 || (head && z_h && z_obj && tmrp && logit), "decoder_forward_tc: null pointer");
  GLDM_REQUIRE(n >= 0 && grasps_per_obj > 0 && D > 0 && D <= 64, "decoder_forward_tc: bad sizes");
  if (n == 0) return GLDM_OK;
  p.mode = 2; p.n = n; p.gpo = grasps_per_obj; p.x_in = z_h; p.z_cond = z_obj; p.n_steps = 1;
  p.head = head; p.D = D; p.tmrp = tmrp; p.logit = logit;
  return launch_tc(p, (cudaStream_t)stream);
}
