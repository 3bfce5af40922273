// Row-major instantiation of the persistent T-step sampler (included by sampler_tc.cu; L = 4 latent denoiser).
//
// resnet_tc_kernel puts the WEIGHTS on the M side of the UMMA (thread = output channel).  For this network that is the
// wrong way round: three of its four stages are 4, 32 and 64 channels wide, so most 128-lane tiles are padding, and
// every GroupNorm / LayerNorm / attention reduction runs ACROSS lanes (20 % of the executed instructions were
// select / shuffle / add triples).  Here the operand roles are swapped:
//
//   A operand (M = 128 rows)  = activations, row = position * 32 + sample (32 samples per CTA), written by the epilogue
//                               as bf16 K-major SWIZZLE_128B slabs of 64 channels with 32-row zero halos: a k = 3 tap is
//                               a 32-row shift of the descriptor start address
//   B operand (N = layer width, 16 ... 256) = the same packed weight images, streamed from L2 through the ring
//   D (TMEM)                  = [row][channel]: a thread owns one row, so channel reductions are in-thread, the next
//                               operand is written with 16-byte stores, and a narrow layer costs N columns, not a tile
//
// 16 epilogue warps: warp-group g (4 warps = the 4 positions, lane = sample) owns GroupNorm group g of the layer
// (channels [g * ch/4, (g+1) * ch/4)), or attention head g.  What still crosses threads is small: GroupNorm sums over
// the 4 positions of a sample (4 warps of a warp-group), LayerNorm / final-conv sums over the 4 warp-groups of a row,
// and the k / v rows of the other 3 positions in the linear attention (through the idle operand buffer).
//
// FiLM is hoisted out of the sample loop: the conditioning embedding depends only on (object, step), so a prologue
// kernel (film_table_kernel, sampler_tc.cu) writes one [A | B] vector pair per FiLM layer and (object, step) - GroupNorm
// affine and FiLM folded into one multiply-add - and the epilogue reads it through L1 (prefetched while the UMMAs run).
// No FiLM UMMAs, no per-step operand rebuild, no FiLM columns in TMEM.
//
// Layers of 32 / 64 / 128 channels (8 / 16 / 32 channels per thread) keep their accumulator row in REGISTERS between
// the statistics and the apply pass: TMEM is read once per element (TMEM read bandwidth, 64 B/cycle/SM, is the wall of
// this epilogue).  The 256-wide final block (64 channels per thread) and the 4-channel first stage use the two-pass form.
//
// TMEM columns: [0,256) accumulator (q | k | v = [0,384) for the attention job), [384,512) residual stream: fp32 for
// widths <= 128, packed bf16 pairs for the 256-wide final block.
#pragma once

namespace rows {
template <int L> struct Geo { static constexpr int NS = 128 / L; };   // samples per CTA: 32 (L = 4) or 8 (L = 16)
constexpr int NEPI = 512;                       // 16 epilogue warps
constexpr int NTHREADS = 640;                   // + producer, issuer, two idle register donors (warps are allocated in fours)
constexpr int SLAB = 160 * 128;                 // slab stride: 64 channels x (32 halo + 128) rows; the upper 32-row halo of a slab
                                                // IS the lower halo of the next one (both stay zero), one extra halo after the last
constexpr int CHUNK = stc::CHUNK, STAGES = 3;      // 96 KB in flight: the 256-wide jobs were bound by ring depth x L2 latency
constexpr int MAXRJ = 40, MAXOPS = 400, MAXCHUNKS = 256;
constexpr uint32_t T_ACC = 0, T_RES = 384;
// shared memory map (from a 1024-aligned base)
constexpr int SM_A = 0;
constexpr int SM_RING = SM_A + 4 * SLAB + 32 * 128;
constexpr int PAR_FLOATS = 5120;                          // every per-channel parameter of the network, resident for all steps
constexpr int SM_PAR = SM_RING + STAGES * CHUNK;
constexpr int SM_PTAB = SM_PAR + PAR_FLOATS * 4;          // [MAXRJ][5] int16: offsets of a job's bias, gamma, beta, g1, g2 (-1: none)
constexpr int SM_JD = SM_PTAB + MAXRJ * 5 * 2 + 16;       // [MAXRJ] uint4 job descriptors read by the epilogue warps
constexpr int SM_XG = SM_JD + MAXRJ * 16;                 // GroupNorm exchange [4 groups][4 positions][32] float2
constexpr int SM_XL = SM_XG + 4 * 4 * 32 * 8;             // row exchange [128 rows][4 warp-groups] float2
constexpr int SM_X = SM_XL + 128 * 4 * 8;                 // state x [32][4], then y, z (evaluation programs), x_in, scaled input
constexpr int SM_BAR = SM_X + 5 * 32 * 4 * 4;
constexpr int SM_CHUNKS = SM_BAR + 256;
constexpr int SM_OPS = SM_CHUNKS + MAXCHUNKS * 8;
constexpr int SM_OPBEG = SM_OPS + MAXOPS * 16;
constexpr int SM_TOTAL = SM_OPBEG + (MAXRJ + 2) * 2 + 16;
static_assert(SM_TOTAL + 1024 <= 232448, "shared memory budget");

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int NV>
__device__ __forceinline__ void ld_film(const float* f, float (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; i += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(f + i));
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
}

__device__ __forceinline__ void bar_wg(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 5, 512;" ::: "memory"); }
// job-start barrier: the 16 epilogue warps PARK here (a hardware barrier takes no issue slots) until the waiter warp, the
// only one that polls the accumulator mbarrier, arrives.  512 threads spinning on mbarrier.try_wait took 22 % of all
// issued instructions and starved the UMMA issuer warp of issue slots.
__device__ __forceinline__ void bar_job_sync() { asm volatile("bar.sync 6, 544;" ::: "memory"); }
__device__ __forceinline__ void bar_job_arrive() { asm volatile("bar.arrive 6, 544;" ::: "memory"); }

// per-thread epilogue state, kept small: the epilogue warps run at the register limit
struct Ep {
  uint32_t tmem;        // TMEM base + this warp's lane quarter
  int g, q, lane;       // warp-group, TMEM lane quarter (warp & 3), lane: row = q * 32 + lane
  int pos, s, row;      // position and sample of the row: row = pos * NS + s (L = 4: pos = q, s = lane)
  uint8_t* smem;
  uint32_t po[3];       // this job's channel parameters (float offsets into SM_PAR): bias | gamma << 16, beta | g1 << 16, g2
  __device__ __forceinline__ float2* xg() const { return reinterpret_cast<float2*>(smem + SM_XG); }
  __device__ __forceinline__ float2* xl() const { return reinterpret_cast<float2*>(smem + SM_XL); }
};

// GroupNorm statistics of a sample's group: this thread's partial (sum, sum of squares) over its channels, summed over all
// L positions of the sample.  A warp holds 32 / NS positions of a sample (lanes s, s + NS, ...): butterfly over those,
// then the four lane quarters through shared memory.
template <int L>
__device__ __forceinline__ float2 gn_total(const Ep& e, float2 part) {
  if (L == 16) {
    part.x += __shfl_xor_sync(0xffffffffu, part.x, 8);  part.y += __shfl_xor_sync(0xffffffffu, part.y, 8);
    part.x += __shfl_xor_sync(0xffffffffu, part.x, 16); part.y += __shfl_xor_sync(0xffffffffu, part.y, 16);
  }
  e.xg()[(e.g * 4 + e.q) * 32 + e.lane] = part;
  bar_wg(e.g);
  float2 t = e.xg()[(e.g * 4) * 32 + e.lane];
#pragma unroll
  for (int pp = 1; pp < 4; ++pp) { const float2 u = e.xg()[(e.g * 4 + pp) * 32 + e.lane]; t.x += u.x; t.y += u.y; }
  return t;
}

template <int NV>
__device__ __forceinline__ void ld_cols(const Ep& e, uint32_t col, float (&v)[NV]) {
  uint32_t r[NV];
  if (NV == 8) tmem_ld8(e.tmem + col, *reinterpret_cast<uint32_t(*)[8]>(&r));
  else tmem_ld4(e.tmem + col, *reinterpret_cast<uint32_t(*)[4]>(&r));
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __uint_as_float(r[i]);
}
// asynchronous forms: issue the loads of a chunk back to back, one tcgen05.wait::ld, then convert
template <int NV>
__device__ __forceinline__ void ld_issue(const Ep& e, uint32_t col, uint32_t (&r)[NV]) {
  if (NV == 8) tmem_ld8(e.tmem + col, *reinterpret_cast<uint32_t(*)[8]>(&r));
  else tmem_ld4(e.tmem + col, *reinterpret_cast<uint32_t(*)[4]>(&r));
}
template <int NV>
__device__ __forceinline__ void ld_use(const uint32_t (&r)[NV], float (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    uint32_t x = r[i];
    asm volatile("" : "+r"(x));     // ordered after the tcgen05.wait::ld (volatile asm keeps program order)
    v[i] = __uint_as_float(x);
  }
}
template <int NV>
__device__ __forceinline__ void st_cols(const Ep& e, uint32_t col, const float (&v)[NV]) {
  uint32_t r[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) r[i] = __float_as_uint(v[i]);
  if (NV == 8) tmem_st8(e.tmem + col, *reinterpret_cast<const uint32_t(*)[8]>(&r));
  else tmem_st4(e.tmem + col, *reinterpret_cast<const uint32_t(*)[4]>(&r));
}
template <int NV>
__device__ __forceinline__ void ld_par(const Ep& e, int k, int c, float (&v)[NV]) {
  const uint32_t off = (k & 1) ? e.po[k >> 1] >> 16 : e.po[k >> 1] & 0xffffu;
  const float* p = reinterpret_cast<const float*>(e.smem + SM_PAR) + off + c;
#pragma unroll
  for (int i = 0; i < NV; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
}
// residual chunk (channels c .. c+NV-1 of the job): fp32 columns, or packed bf16 pairs for the 256-wide block
template <int NV>
__device__ __forceinline__ void ld_res(const Ep& e, bool packed, int chan, float (&v)[NV]) {
  if (!packed) { ld_cols<NV>(e, T_RES + chan, v); return; }
  uint32_t r[4];
  tmem_ld4(e.tmem + T_RES + (chan >> 1), r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&r[i]);
    v[2 * i] = __low2float(h);
    v[2 * i + 1] = __high2float(h);
  }
}
template <int NV>
__device__ __forceinline__ void st_res(const Ep& e, bool packed, int chan, const float (&v)[NV]) {
  if (!packed) { st_cols<NV>(e, T_RES + chan, v); return; }
  uint32_t r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = (2 * i + 1 < NV) ? pack_bf16(v[2 * i], v[2 * i + 1]) : 0u;
  tmem_st4(e.tmem + T_RES + (chan >> 1), r);
}
// next operand: channels chan .. chan+NV-1 of this thread's row
template <int NV>
__device__ __forceinline__ void st_operand(const Ep& e, int chan, const float (&v)[NV]) {
  const int R = 32 + e.row;
  uint8_t* dst = e.smem + SM_A + (chan >> 6) * SLAB + R * 128 + ((((chan & 63) >> 3) ^ (R & 7)) << 4) + (chan & 7) * 2;
  if (NV == 8) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                                pack_bf16(v[6], v[7]));
  } else {
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
  }
}

// 4-channel first stage: warp-group 0 alone, a thread holds the 4 channels of its row = 4 GroupNorm groups of one channel
// (statistics over the 4 positions of the sample); LayerNorm over the 4 channels is in-thread.  Static register indexing.
__device__ __forceinline__ void narrow_epilogue(const Ep& e, int flags, const float* film) {
  if (e.g != 0) return;
  float x[4], a[4], b[4];
  ld_cols<4>(e, T_ACC, x);
  ld_par<4>(e, 0, 0, b);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] += b[i];
  if (flags & E_GN) {
#pragma unroll
    for (int i = 0; i < 4; ++i) e.xg()[(i * 4 + e.pos) * 32 + e.s] = make_float2(x[i], x[i] * x[i]);
    bar_wg(0);
    if (flags & E_FILM) {
      ld_film<4>(film, a);
      ld_film<4>(film + 4, b);
    } else {
      ld_par<4>(e, 1, 0, a);
      ld_par<4>(e, 2, 0, b);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float ts = 0.f, tq = 0.f;
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) { const float2 t = e.xg()[(i * 4 + pp) * 32 + e.s]; ts += t.x; tq += t.y; }
      const float m = ts * 0.25f, r = rsqrtf(fmaxf(tq * 0.25f - m * m, 0.f) + 1e-5f);
      x[i] = fmaf(x[i] - m, r * a[i], b[i]);
    }
  }
  if (flags & E_SILU) {
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      const float2 t = silu_fast2(make_float2(x[i], x[i + 1]));
      x[i] = t.x; x[i + 1] = t.y;
    }
  }
  if (flags & E_LN) {
    const float m = 0.25f * ((x[0] + x[1]) + (x[2] + x[3]));
    const float q = 0.25f * (fmaf(x[0], x[0], x[1] * x[1]) + fmaf(x[2], x[2], x[3] * x[3]));
    const float r = rsqrtf(fmaxf(q - m * m, 0.f) + 1e-5f);
    ld_par<4>(e, 3, 0, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (x[i] - m) * r * a[i];
  }
  if (flags & E_ADDRES) {
    ld_cols<4>(e, T_RES, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] += b[i];
  }
  if (flags & E_STORERES) {
    st_cols<4>(e, T_RES, x);
    tmem_st_wait();
  }
  if (flags & E_LNNEXT) {       // PreNorm of the attention: operand = LayerNorm(result) * g2
    const float m = 0.25f * ((x[0] + x[1]) + (x[2] + x[3]));
    const float q = 0.25f * (fmaf(x[0], x[0], x[1] * x[1]) + fmaf(x[2], x[2], x[3] * x[3]));
    const float r = rsqrtf(fmaxf(q - m * m, 0.f) + 1e-5f);
    ld_par<4>(e, 4, 0, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (x[i] - m) * r * a[i];
  }
  st_operand<4>(e, 0, x);
}

// 16-channel first stage of the L = 16 networks (grasp decoder, ppc latent denoiser): 4 channels per thread = GroupNorm
// group g; same recipe flags as reg_epilogue.
template <int L>
__device__ __forceinline__ void quad_epilogue(const Ep& e, int flags, const float* film) {
  constexpr int CH = 16, CW = 4;
  const int c = e.g * CW;
  float x[4], a[4], b[4];
  ld_cols<4>(e, T_ACC + c, x);
  ld_par<4>(e, 0, c, b);
  float sm = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { x[i] += b[i]; sm += x[i]; sq = fmaf(x[i], x[i], sq); }
  float mean = 0.f, rstd = 1.f;
  if (flags & (E_GN | E_LN)) {
    float ts, tq, inv;
    if (flags & E_GN) {
      const float2 t = gn_total<L>(e, make_float2(sm, sq));
      ts = t.x; tq = t.y; inv = 1.0f / (float)(CW * L);
    } else {
      e.xl()[e.row * 4 + e.g] = make_float2(sm, sq);
      bar_all();
      ts = 0.f; tq = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) { const float2 t = e.xl()[e.row * 4 + w]; ts += t.x; tq += t.y; }
      inv = 1.0f / (float)CH;
    }
    mean = ts * inv;
    rstd = rsqrtf(fmaxf(tq * inv - mean * mean, 0.f) + 1e-5f);
  }
  if (flags & E_GN) {
    if (flags & E_FILM) {
      ld_film<4>(film + c, a);
      ld_film<4>(film + CH + c, b);
    } else {
      ld_par<4>(e, 1, c, a);
      ld_par<4>(e, 2, c, b);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = fmaf(x[i] - mean, rstd * a[i], b[i]);
  }
  if (flags & E_SILU) {
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      const float2 t = silu_fast2(make_float2(x[i], x[i + 1]));
      x[i] = t.x; x[i + 1] = t.y;
    }
  }
  if (flags & E_LN) {
    ld_par<4>(e, 3, c, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (x[i] - mean) * rstd * a[i];
  }
  if (flags & E_ADDRES) {
    ld_cols<4>(e, T_RES + c, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] += b[i];
  }
  if (flags & E_STORERES) {
    st_cols<4>(e, T_RES + c, x);
    tmem_st_wait();
  }
  if (flags & E_LNNEXT) {       // PreNorm of the attention: operand = LayerNorm(result) * g2 over the 16 channels of the row
    e.xl()[e.row * 4 + e.g] = make_float2((x[0] + x[1]) + (x[2] + x[3]), fmaf(x[0], x[0], x[1] * x[1]) + fmaf(x[2], x[2], x[3] * x[3]));
    bar_all();
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { const float2 t = e.xl()[e.row * 4 + w]; ts += t.x; tq += t.y; }
    const float m = ts * (1.0f / CH), r = rsqrtf(fmaxf(tq * (1.0f / CH) - m * m, 0.f) + 1e-5f);
    ld_par<4>(e, 4, c, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (x[i] - m) * r * a[i];
  }
  st_operand<4>(e, c, x);
}

// packed fp32 pairs: FADD2 / FMUL2 / FFMA2 do two values per issue slot, and the epilogue is issue bound
struct F8 { float2 p[4]; };
__device__ __forceinline__ F8 ld_par8(const Ep& e, int k, int c) {
  const uint32_t off = (k & 1) ? e.po[k >> 1] >> 16 : e.po[k >> 1] & 0xffffu;
  const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.smem + SM_PAR) + off + c);
  const float4 t0 = p[0], t1 = p[1];
  F8 r;
  r.p[0] = make_float2(t0.x, t0.y); r.p[1] = make_float2(t0.z, t0.w);
  r.p[2] = make_float2(t1.x, t1.y); r.p[3] = make_float2(t1.z, t1.w);
  return r;
}
__device__ __forceinline__ F8 ld_film8(const float* f) {
  const float4 t0 = __ldg(reinterpret_cast<const float4*>(f)), t1 = __ldg(reinterpret_cast<const float4*>(f) + 1);
  F8 r;
  r.p[0] = make_float2(t0.x, t0.y); r.p[1] = make_float2(t0.z, t0.w);
  r.p[2] = make_float2(t1.x, t1.y); r.p[3] = make_float2(t1.z, t1.w);
  return r;
}
// 8 accumulator words (after the tcgen05.wait::ld) as four float pairs
__device__ __forceinline__ void use8(const uint32_t* r, float2 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t a = r[2 * i], b = r[2 * i + 1];
    asm volatile("" : "+r"(a), "+r"(b));     // ordered after the wait
    v[i] = make_float2(__uint_as_float(a), __uint_as_float(b));
  }
}
__device__ __forceinline__ void st_operand8(const Ep& e, int chan, const float2 (&v)[4]) {
  const int R = 32 + e.row;
  uint8_t* dst = e.smem + SM_A + (chan >> 6) * SLAB + R * 128 + ((((chan & 63) >> 3) ^ (R & 7)) << 4);
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(v[0].x, v[0].y), pack_bf16(v[1].x, v[1].y), pack_bf16(v[2].x, v[2].y),
                                              pack_bf16(v[3].x, v[3].y));
}
// v = (v - mean) * (rstd * a) + b for 8 channels
__device__ __forceinline__ void norm8(float2 (&v)[4], const F8& a, const F8& b, float2 nmean, float2 rstd) {
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __ffma2_rn(__fadd2_rn(v[i], nmean), __fmul2_rn(a.p[i], rstd), b.p[i]);
}
__device__ __forceinline__ void silu8(float2 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = silu_fast2(v[i]);
}

// register-resident layer epilogue: ch = 32 / 64 / 128, this thread's cw = ch / 4 channels ([g * cw, (g+1) * cw) = one
// GroupNorm group) stay in x[] from the single TMEM read to the operand store; chunks of 8 channels, k < cw / 8.
template <int L>
__device__ __forceinline__ void reg_epilogue(const Ep& e, int flags, int ch, const float* film, long long* rec) {
  const int cw = ch >> 2, nk = cw >> 3, c_lo = e.g * cw;
  float2 x[4][4];
  float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
  {
    uint32_t r[32];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nk) tmem_ld8(e.tmem + T_ACC + c_lo + 8 * k, *reinterpret_cast<uint32_t(*)[8]>(&r[8 * k]));
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nk) {
        const F8 b = ld_par8(e, 0, c_lo + 8 * k);
        use8(&r[8 * k], x[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          x[k][i] = __fadd2_rn(x[k][i], b.p[i]);
          s2 = __fadd2_rn(s2, x[k][i]);
          q2 = __ffma2_rn(x[k][i], x[k][i], q2);
        }
      }
  }
  if (rec) rec[3] = clock64();
  float mean = 0.f, rstd = 1.f;           // GroupNorm of this thread's group, or channel LayerNorm of its row
  if (flags & (E_GN | E_LN)) {
    float ts = 0.f, tq = 0.f, inv;
    if (flags & E_GN) {                   // sums over the L positions of a sample
      const float2 t = gn_total<L>(e, make_float2(s2.x + s2.y, q2.x + q2.y));
      ts = t.x; tq = t.y;
      inv = 1.0f / (float)(cw * L);
    } else {                              // sums over the 4 warp-groups of a row
      e.xl()[e.row * 4 + e.g] = make_float2(s2.x + s2.y, q2.x + q2.y);
      bar_all();
#pragma unroll
      for (int w = 0; w < 4; ++w) { const float2 t = e.xl()[e.row * 4 + w]; ts += t.x; tq += t.y; }
      inv = 1.0f / (float)ch;
    }
    mean = ts * inv;
    rstd = rsqrtf(fmaxf(tq * inv - mean * mean, 0.f) + 1e-5f);
  }
  if (rec) rec[4] = clock64();
  const float2 nmean = make_float2(-mean, -mean), rstd2 = make_float2(rstd, rstd);
  float2 l2 = make_float2(0.f, 0.f), lq2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < nk) {
      const int c = c_lo + 8 * k;
      float2(&v)[4] = x[k];
      uint32_t rr[8];
      if (flags & E_ADDRES) tmem_ld8(e.tmem + T_RES + c, rr);        // residual chunk in flight under the math below
      if (flags & E_GN) {
        if (flags & E_FILM) norm8(v, ld_film8(film + c), ld_film8(film + ch + c), nmean, rstd2);
        else norm8(v, ld_par8(e, 1, c), ld_par8(e, 2, c), nmean, rstd2);
      }
      if (flags & E_SILU) silu8(v);
      if (flags & E_LN) {
        const F8 g1 = ld_par8(e, 3, c);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __fmul2_rn(__fadd2_rn(v[i], nmean), __fmul2_rn(g1.p[i], rstd2));
      }
      if (flags & E_ADDRES) {
        tmem_ld_wait();
        float2 t[4];
        use8(rr, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __fadd2_rn(v[i], t[i]);
      }
      if (flags & E_STORERES) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { w[2 * i] = __float_as_uint(v[i].x); w[2 * i + 1] = __float_as_uint(v[i].y); }
        tmem_st8(e.tmem + T_RES + c, w);
      }
      if (flags & E_LNNEXT) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { l2 = __fadd2_rn(l2, v[i]); lq2 = __ffma2_rn(v[i], v[i], lq2); }
      } else {
        st_operand8(e, c, v);
      }
    }
  }
  if (flags & E_STORERES) tmem_st_wait();
  if (rec) rec[5] = clock64();
  if (flags & E_LNNEXT) {
    // PreNorm of the attention: operand = LayerNorm(result) * g2, the result is still in x[]
    e.xl()[e.row * 4 + e.g] = make_float2(l2.x + l2.y, lq2.x + lq2.y);
    bar_all();
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { const float2 t = e.xl()[e.row * 4 + w]; ts += t.x; tq += t.y; }
    const float inv = 1.0f / (float)ch;
    const float m = ts * inv, r = rsqrtf(fmaxf(tq * inv - m * m, 0.f) + 1e-5f);
    const float2 nm = make_float2(-m, -m), r2 = make_float2(r, r);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nk) {
        const F8 g2 = ld_par8(e, 4, c_lo + 8 * k);
        float2 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fmul2_rn(__fadd2_rn(x[k][i], nm), __fmul2_rn(g2.p[i], r2));
        st_operand8(e, c_lo + 8 * k, o);
      }
  }
}

// 256-wide layers (final block, last stage conv): 64 channels per thread do not fit the register file, so the
// accumulator is read twice (statistics, apply) - 32 / 16 columns per tcgen05.wait::ld.  Residual stream: packed bf16 pairs.
// Returns the final-conv partial dot product (E_FINAL) of this thread's channels.
template <int L>
__device__ __forceinline__ float wide_epilogue(const Ep& e, int flags, const float* film) {
  constexpr int CH = 256, CW = 64;
  const int c_lo = e.g * CW;
  float mean = 0.f, rstd = 1.f;
  if (flags & E_GN) {
    float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      uint32_t r[32];
#pragma unroll
      for (int k = 0; k < 4; ++k) tmem_ld8(e.tmem + T_ACC + c_lo + 32 * h + 8 * k, *reinterpret_cast<uint32_t(*)[8]>(&r[8 * k]));
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const F8 b = ld_par8(e, 0, c_lo + 32 * h + 8 * k);
        float2 v[4];
        use8(&r[8 * k], v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = __fadd2_rn(v[i], b.p[i]);
          s2 = __fadd2_rn(s2, v[i]);
          q2 = __ffma2_rn(v[i], v[i], q2);
        }
      }
    }
    const float2 t = gn_total<L>(e, make_float2(s2.x + s2.y, q2.x + q2.y));
    const float inv = 1.0f / (float)(CW * L);
    mean = t.x * inv;
    rstd = rsqrtf(fmaxf(t.y * inv - mean * mean, 0.f) + 1e-5f);
  }
  const float2 nmean = make_float2(-mean, -mean), rstd2 = make_float2(rstd, rstd);
  float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int c = c_lo; c < c_lo + CW; c += 16) {
    uint32_t r[16], rr[8];
    tmem_ld8(e.tmem + T_ACC + c, *reinterpret_cast<uint32_t(*)[8]>(&r[0]));
    tmem_ld8(e.tmem + T_ACC + c + 8, *reinterpret_cast<uint32_t(*)[8]>(&r[8]));
    if (flags & E_ADDRES) tmem_ld8(e.tmem + T_RES + (c >> 1), rr);
    tmem_ld_wait();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int cc = c + 8 * hh;
      float2 v[4];
      use8(&r[8 * hh], v);
      {
        const F8 b = ld_par8(e, 0, cc);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __fadd2_rn(v[i], b.p[i]);
      }
      if (flags & E_GN) {
        if (flags & E_FILM) norm8(v, ld_film8(film + cc), ld_film8(film + CH + cc), nmean, rstd2);
        else norm8(v, ld_par8(e, 1, cc), ld_par8(e, 2, cc), nmean, rstd2);
      }
      if (flags & E_SILU) silu8(v);
      if (flags & E_ADDRES) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t t = rr[4 * hh + i];
          asm volatile("" : "+r"(t));
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&t);
          v[i] = __fadd2_rn(v[i], make_float2(__low2float(h2), __high2float(h2)));
        }
      }
      if (flags & E_STORERES) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = pack_bf16(v[i].x, v[i].y);
        tmem_st4(e.tmem + T_RES + (cc >> 1), w);
      }
      if (flags & E_FINAL) {
        const F8 w = ld_par8(e, 4, cc);
#pragma unroll
        for (int i = 0; i < 4; ++i) dot2 = __ffma2_rn(w.p[i], v[i], dot2);
      } else {
        st_operand8(e, cc, v);
      }
    }
  }
  if (flags & E_STORERES) tmem_st_wait();
  return dot2.x + dot2.y;
}

// exchange rows of the attention: bf16, 32 values = four 16-byte chunks per row; region 0 / 1 use complementary halves
// of the 8 swizzled 16-byte slots of a 128-byte operand row (slot = chunk ^ (sample & 7), region 1 flips bit 2), so a
// warp's 512-byte access always spreads over all 32 banks
__device__ __forceinline__ uint4* xslot(uint8_t* base, int region, int pos, int s, int j) {
  return reinterpret_cast<uint4*>(base + (32 + pos * 32 + s) * 128 + (((j ^ (s & 7)) ^ (region << 2)) << 4));
}
__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    f[2 * i] = __low2float(h);
    f[2 * i + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// linear attention core (resnets.py:211-235) on the q | k | v accumulator; warp-group = head, thread = (position n, sample).
//   out[e][n] = sum_n' A[n][n'] v[e][n'],  A[n][n'] = sum_d q^[d][n] k^[d][n'],
//   q^ = softmax over d of q (* 32^-0.5),  k^ = softmax over the 4 positions of k.
// The four threads of a (sample, head) split the work by quarters of the 32-wide head dimension instead of each doing
// everything for its own position: thread t takes d in [8t, 8t+8) for all positions (partial A, 16 values) and later
// e in [8t, 8t+8) for all positions.  Rows travel through the idle operand slab of the head as bf16 (q^, k, v: every
// thread of a sample sees the same rounded values, so the soft-max over positions stays consistent); the partial A's
// as fp32.
// packed-pair helpers of the attention epilogues: the epilogue is issue bound, so its arithmetic runs on float2 (FFMA2 /
// FADD2 / FMUL2), the exponentials as ex2(x * log2e - m * log2e) with the scale folded into one FFMA2 per pair
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 ex2_pair(float2 x) { return make_float2(ex2f(x.x), ex2f(x.y)); }
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
__device__ __forceinline__ void unpack8p(const uint4 u, float2 (&f)[4]) {
  f[0] = bf2_to_f2(u.x); f[1] = bf2_to_f2(u.y); f[2] = bf2_to_f2(u.z); f[3] = bf2_to_f2(u.w);
}
__device__ __forceinline__ uint4 pack8p(const float2* f) {
  return make_uint4(pack_bf16(f[0].x, f[0].y), pack_bf16(f[1].x, f[1].y), pack_bf16(f[2].x, f[2].y), pack_bf16(f[3].x, f[3].y));
}
constexpr float kLog2e = 1.4426950408889634f;
// q^ = softmax over the 32 head channels of this thread's row, times 32^-0.5: 16 accumulator pairs in, 16 pairs out
__device__ __forceinline__ void softmax32_scaled(float2 (&q)[16]) {
  float m = fmaxf(q[0].x, q[0].y);
#pragma unroll
  for (int i = 1; i < 16; ++i) m = fmaxf(fmaxf(m, q[i].x), q[i].y);
  const float2 l2 = make_float2(kLog2e, kLog2e), nm = make_float2(-m * kLog2e, -m * kLog2e);
  float2 sum = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) { q[i] = ex2_pair(__ffma2_rn(q[i], l2, nm)); sum = __fadd2_rn(sum, q[i]); }
  const float sc = __fdividef(0.17677669529663687f, sum.x + sum.y);
  const float2 sc2 = make_float2(sc, sc);
#pragma unroll
  for (int i = 0; i < 16; ++i) q[i] = __fmul2_rn(q[i], sc2);
}

__device__ __forceinline__ void attention_epilogue(const Ep& e, long long* rec) {
  uint8_t* X = e.smem + SM_A + e.g * SLAB;        // rows 32..159 of this head's slab (the halo rows are not touched)
  const int h = e.g, t = e.pos, s = e.s;
  {
    uint32_t rq[4][8], rk[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ld_issue<8>(e, T_ACC + h * 32 + i * 8, rq[i]);
      ld_issue<8>(e, T_ACC + 128 + h * 32 + i * 8, rk[i]);
    }
    tmem_ld_wait();
    float2 qs[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 kk[4];
      use8(rk[i], kk);
      *xslot(X, 1, t, s, i) = pack8p(kk);               // raw k of this position
      use8(rq[i], *reinterpret_cast<float2(*)[4]>(&qs[4 * i]));
    }
    softmax32_scaled(qs);
#pragma unroll
    for (int j = 0; j < 4; ++j) *xslot(X, 0, t, s, j) = pack8p(&qs[4 * j]);      // q^ of this position
  }
  if (rec) rec[3] = clock64();
  bar_wg(h);
  float Ap[16];                                       // partial A[n][n'] over this thread's quarter of d
  {
    float2 q[4][4], k[4][4];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      unpack8p(*xslot(X, 0, pp, s, t), q[pp]);
      unpack8p(*xslot(X, 1, pp, s, t), k[pp]);
    }
    const float2 l2 = make_float2(kLog2e, kLog2e);
#pragma unroll
    for (int i = 0; i < 4; ++i) {                     // soft-max over the 4 positions, two channels at a time
      const float mx = fmaxf(fmaxf(k[0][i].x, k[1][i].x), fmaxf(k[2][i].x, k[3][i].x));
      const float my = fmaxf(fmaxf(k[0][i].y, k[1][i].y), fmaxf(k[2][i].y, k[3][i].y));
      const float2 nm = make_float2(-mx * kLog2e, -my * kLog2e);
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) k[pp][i] = ex2_pair(__ffma2_rn(k[pp][i], l2, nm));
      const float2 z = __fadd2_rn(__fadd2_rn(k[0][i], k[1][i]), __fadd2_rn(k[2][i], k[3][i]));
      const float2 zi = make_float2(__fdividef(1.0f, z.x), __fdividef(1.0f, z.y));
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) k[pp][i] = __fmul2_rn(k[pp][i], zi);
    }
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int n1 = 0; n1 < 4; ++n1) {
        float2 a = __fmul2_rn(q[n][0], k[n1][0]);
#pragma unroll
        for (int i = 1; i < 4; ++i) a = __ffma2_rn(q[n][i], k[n1][i], a);
        Ap[n * 4 + n1] = a.x + a.y;
      }
  }
  if (rec) rec[4] = clock64();
  bar_wg(h);                                          // every q^ / k row has been read
  {
    uint32_t rv[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_issue<8>(e, T_ACC + 256 + h * 32 + i * 8, rv[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(xslot(X, 0, t, s, j)) = make_float4(Ap[4 * j], Ap[4 * j + 1], Ap[4 * j + 2], Ap[4 * j + 3]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 vv[4];
      use8(rv[i], vv);
      *xslot(X, 1, t, s, i) = pack8p(vv);             // v of this position
    }
  }
  bar_wg(h);
  if (rec) rec[5] = clock64();
  float2 o[4][4];                                     // out[position n][e in quarter t], channel pairs
  {
    float A[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 x0 = *reinterpret_cast<const float4*>(xslot(X, 0, 0, s, j)), x1 = *reinterpret_cast<const float4*>(xslot(X, 0, 1, s, j));
      const float4 x2 = *reinterpret_cast<const float4*>(xslot(X, 0, 2, s, j)), x3 = *reinterpret_cast<const float4*>(xslot(X, 0, 3, s, j));
      const float2 lo = __fadd2_rn(__fadd2_rn(make_float2(x0.x, x0.y), make_float2(x1.x, x1.y)),
                                   __fadd2_rn(make_float2(x2.x, x2.y), make_float2(x3.x, x3.y)));
      const float2 hi = __fadd2_rn(__fadd2_rn(make_float2(x0.z, x0.w), make_float2(x1.z, x1.w)),
                                   __fadd2_rn(make_float2(x2.z, x2.w), make_float2(x3.z, x3.w)));
      A[4 * j] = lo.x; A[4 * j + 1] = lo.y; A[4 * j + 2] = hi.x; A[4 * j + 3] = hi.y;
    }
    float2 v[4][4];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) unpack8p(*xslot(X, 1, pp, s, t), v[pp]);
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 a = __fmul2_rn(v[0][i], make_float2(A[n * 4], A[n * 4]));
#pragma unroll
        for (int n1 = 1; n1 < 4; ++n1) a = __ffma2_rn(v[n1][i], make_float2(A[n * 4 + n1], A[n * 4 + n1]), a);
        o[n][i] = a;
      }
  }
  bar_all();                                          // the exchange rows become operand rows again
  // channels 32h + 8t .. + 8 of the four rows (n, s)
  {
    const int chan = h * 32 + t * 8;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int R = 32 + n * 32 + s;
      *reinterpret_cast<uint4*>(e.smem + SM_A + (chan >> 6) * SLAB + R * 128 + ((((chan & 63) >> 3) ^ (R & 7)) << 4)) = pack8p(o[n]);
    }
  }
}

// ---- linear attention for L = 16 (8 samples per CTA, row = position * 8 + sample) -----------------------------------
// Per (sample, head): out[n][e] = sum_n' S[n][n'] v[n'][e],  S = Q^ K^T,  Q^[n][d] = softmax_d(q[n][d]) * 32^-0.5,
// K^[n'][d] = softmax over the 16 positions n' of k[n'][d]   (resnets.py:211-235).  16 x 16 x 32 per pair is too much to
// pass around thread by thread (every thread would read every k / v row), so the rows go to shared memory once as bf16
// [sample][position][32] matrices and the two small products run on mma.sync (m16n8k16, fp32 accumulate): warp q of
// head h takes samples 2q and 2q + 1.  Matrices live in rows 32..159 of the head's idle operand slab: K^ (8 KB), then
// Q^ and later V (8 KB).  16-byte chunk c of row (s, n) sits at chunk c ^ (n >> 1) ^ s: conflict-free for the row
// writers (lanes = 4 positions x 8 samples) and for ldmatrix (8 consecutive positions of one sample).
__device__ __forceinline__ uint8_t* arow(uint8_t* base, int s, int n, int c) {
  return base + (s * 16 + n) * 64 + (((c ^ (n >> 1) ^ s) & 3) << 4);
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void attention16_epilogue(const Ep& e) {
  uint8_t* Kb = e.smem + SM_A + e.g * SLAB + 32 * 128;      // K^ [8][16][32] bf16
  uint8_t* Qb = Kb + 8192;                                   // Q^, later V
  const int h = e.g, n = e.pos, s = e.s;
  {  // raw k row of this (position, sample)
    uint32_t rk[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_issue<8>(e, T_ACC + 128 + h * 32 + i * 8, rk[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t[8];
      ld_use<8>(rk[i], t);
      *reinterpret_cast<uint4*>(arow(Kb, s, n, i)) = pack8(t);
    }
  }
  {  // Q^ row: softmax over the 32 head channels, in-thread
    float2 qs[16];
    uint32_t rq[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_issue<8>(e, T_ACC + h * 32 + i * 8, rq[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) use8(rq[i], *reinterpret_cast<float2(*)[4]>(&qs[4 * i]));
    softmax32_scaled(qs);
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(arow(Qb, s, n, i)) = pack8p(&qs[4 * i]);
  }
  bar_wg(h);
  {  // K^: soft-max over the 16 positions, in place.  Thread (of the 128 of the head) = (sample, channel pair).
    const int tl = e.q * 32 + e.lane, ks = tl >> 4, j = tl & 15;
    float2 v[16];
    float2 mx = make_float2(-INFINITY, -INFINITY);
#pragma unroll
    for (int nn = 0; nn < 16; ++nn) {
      const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(arow(Kb, ks, nn, j >> 2) + (j & 3) * 4);
      v[nn] = make_float2(__low2float(hh), __high2float(hh));
      mx.x = fmaxf(mx.x, v[nn].x); mx.y = fmaxf(mx.y, v[nn].y);
    }
    float2 z = make_float2(0.f, 0.f);
#pragma unroll
    for (int nn = 0; nn < 16; ++nn) {
      v[nn].x = __expf(v[nn].x - mx.x); v[nn].y = __expf(v[nn].y - mx.y);
      z.x += v[nn].x; z.y += v[nn].y;
    }
    z.x = __fdividef(1.0f, z.x); z.y = __fdividef(1.0f, z.y);
#pragma unroll
    for (int nn = 0; nn < 16; ++nn)
      *reinterpret_cast<uint32_t*>(arow(Kb, ks, nn, j >> 2) + (j & 3) * 4) = pack_bf16(v[nn].x * z.x, v[nn].y * z.y);
  }
  bar_wg(h);
  // S = Q^ K^T for the two samples of this warp (fp32 accumulators), as bf16 A fragments of the second product
  const int l = e.lane;
  uint32_t pa[2][4];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int ss = 2 * e.q + u;
    float sacc[2][4] = {};
    uint32_t a0[4], a1[4], b[4];
    ldsm_x4(a0, arow(Qb, ss, l & 15, l >> 4));            // d 0..15: (rows 0-7, lo), (rows 8-15, lo), (rows 0-7, hi), (rows 8-15, hi)
    ldsm_x4(a1, arow(Qb, ss, l & 15, 2 + (l >> 4)));      // d 16..31
#pragma unroll
    for (int t = 0; t < 2; ++t) {                         // n' tile: positions 8t .. 8t+7, the four 8-channel chunks
      ldsm_x4(b, arow(Kb, ss, 8 * t + (l & 7), l >> 3));
      mma_bf16_16816(sacc[t], a0, b[0], b[1]);
      mma_bf16_16816(sacc[t], a1, b[2], b[3]);
    }
    pa[u][0] = pack_bf16(sacc[0][0], sacc[0][1]); pa[u][1] = pack_bf16(sacc[0][2], sacc[0][3]);
    pa[u][2] = pack_bf16(sacc[1][0], sacc[1][1]); pa[u][3] = pack_bf16(sacc[1][2], sacc[1][3]);
  }
  bar_wg(h);                                              // every Q^ fragment has been read: V takes its place
  {
    uint32_t rv[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_issue<8>(e, T_ACC + 256 + h * 32 + i * 8, rv[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t[8];
      ld_use<8>(rv[i], t);
      *reinterpret_cast<uint4*>(arow(Qb, s, n, i)) = pack8(t);
    }
  }
  bar_wg(h);
  float o[2][4][4];                                       // out[n][e]: [sample][e tile][fragment]
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int ss = 2 * e.q + u;
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2) {                      // e tiles 2 t2, 2 t2 + 1: V^T fragments (k = n', n = e)
      uint32_t b[4];
      ldsm_x4_t(b, arow(Qb, ss, l & 15, 2 * t2 + (l >> 4)));
#pragma unroll
      for (int i = 0; i < 4; ++i) { o[u][2 * t2][i] = 0.f; o[u][2 * t2 + 1][i] = 0.f; }
      mma_bf16_16816(o[u][2 * t2], pa[u], b[0], b[1]);
      mma_bf16_16816(o[u][2 * t2 + 1], pa[u], b[2], b[3]);
    }
  }
  bar_all();                                              // the exchange rows become operand rows again
  {
    const int gid = l >> 2, tig = l & 3;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ss = 2 * e.q + u;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int chan = h * 32 + 8 * t + 2 * tig;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int R = 32 + (gid + 8 * hf) * 8 + ss;
          *reinterpret_cast<uint32_t*>(e.smem + SM_A + (chan >> 6) * SLAB + R * 128 + ((((chan & 63) >> 3) ^ (R & 7)) << 4) +
                                       (chan & 7) * 2) = pack_bf16(o[u][t][2 * hf], o[u][t][2 * hf + 1]);
        }
      }
    }
  }
}

template <int L>
__global__ void __launch_bounds__(NTHREADS, 1) resnet_rows_kernel(const __grid_constant__ TcParams p) {
  constexpr int NS = Geo<L>::NS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* full = bars;          // [STAGES]
  uint64_t* empty = bars + 4;     // [STAGES]
  uint64_t* b_ready = bars + 8;
  uint64_t* acc_ready = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  uint2* chunk_tab = reinterpret_cast<uint2*>(smem + SM_CHUNKS);
  uint4* ops = reinterpret_cast<uint4*>(smem + SM_OPS);
  uint16_t* op_begin = reinterpret_cast<uint16_t*>(smem + SM_OPBEG);
  float* s_par = reinterpret_cast<float*>(smem + SM_PAR);
  int16_t* s_ptab = reinterpret_cast<int16_t*>(smem + SM_PTAB);
  float* s_x = reinterpret_cast<float*>(smem + SM_X);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  int cta_s0 = blockIdx.x * NS;
  const ResNetLayout& lay = p.lay;
  const float* W = p.W;
  // sampler: the denoising steps of this CTA's samples.  Decoder (mode 2): the CTA is persistent over sample groups
  // blockIdx.x, blockIdx.x + gridDim.x, ... - one "step" per group, so the table set-up below (40 k cycles, a fifth of a
  // single evaluation) is paid once per CTA instead of once per 8 grasps
  const int n_groups = (p.n + NS - 1) / NS;
  const int n_steps = (p.mode == 0) ? p.n_steps : (p.mode == 2) ? (n_groups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 1;
  const int n_rj = p.n_jobs;

  // ---- one-time setup
  if (p.prof && blockIdx.x == 0 && tid == 0) p.prof[0] = clock64();
  for (int i = tid; i < SM_RING / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(b_ready, NEPI / 32);
    mbar_init(acc_ready, 1);
    fence_barrier_init();
  }
  if (wid == 0) tmem_alloc<512>(tmem_slot);
  auto load_input = [&]() {          // threads 0 .. NS * L - 1 (= warp-group 0): the state / network input of this group
    const int s = tid / L, l = tid % L;
    float v = 0.f;
    if (cta_s0 + s < p.n) {
      if (p.mode != 2) {   // denoiser: the latent itself
        v = __ldg(p.x_in + (size_t)(cta_s0 + s) * L + l);
      } else {             // decoder in_layer: Linear(D -> L)   (grasp_vae.py:419)
        v = __ldg(p.head + L * p.D + l);
        for (int d = 0; d < p.D; ++d) v = fmaf(__ldg(p.head + l * p.D + d), __ldg(p.x_in + (size_t)(cta_s0 + s) * p.D + d), v);
      }
    }
    s_x[s * L + l] = v;
    s_x[128 + s * L + l] = 0.f;
    s_x[256 + s * L + l] = 0.f;
    if (p.mode == 0 && p.x_all && cta_s0 + s < p.n) p.x_all[(size_t)(cta_s0 + s) * L + l] = v;
  };
  if (tid < NS * L) load_input();
  // ---- weight-chunk and UMMA op tables of one step (only the convolution / projection blocks of a job's image are
  // streamed: the FiLM projection tiles behind them belong to the channel-major kernel)
  // op.x / op.y: low words of the A (activation) / B (weight) descriptors; op.z: high word of the B descriptor;
  // op.w: [0,9) TMEM column, [9,12) K steps, 12 accumulate, 13 first use of a ring chunk, 15 first op of the job,
  //       [16,19) ring stage, 19 ring padding, [20,25) N / 16
  // Thread 0 lays out the prefix sums only (parameter offsets, first op / first chunk of every job); the tables themselves
  // are written by one thread per job.  (A fully serial build cost 80 k cycles: nothing next to a 100-step sampler launch,
  // 40 % on top of the single evaluation of the decoder.)
  uint32_t* chunk_begin = reinterpret_cast<uint32_t*>(smem + SM_XL);       // [n_rj + 1], scratch until the step loop starts
  if (tid == 0) {
    uint32_t nops = 0, ncp = 0;
    int npar = 0;
    for (int j = 0; j < n_rj; ++j) {
      const TcJob& job = p.jobs[j];
      // FiLM layers take their GroupNorm affine from the FiLM table (folded), so gamma / beta are not staged for them
      const bool film = (job.flags & E_FILM) != 0;
      const int src[5] = {job.o_bias, film ? -1 : job.o_gamma, film ? -1 : job.o_beta, job.o_g, job.o_g2};
      for (int k = 0; k < 5; ++k) {
        const bool use = src[k] >= 0 && !(job.flags & E_ATTN) && npar + job.ch <= PAR_FLOATS;
        s_ptab[j * 5 + k] = (int16_t)(use ? npar : -1);
        if (use) npar += (job.ch + 3) & ~3;
      }
      const uint32_t a_swb = job.a_swb, nkb = a_swb == 128 ? (uint32_t)job.kpt >> 6 : 1u;
      const uint32_t mb = job.mtiles * job.taps * nkb * (a_swb << 7);            // bytes of the main blocks
      op_begin[j] = (uint16_t)nops;
      chunk_begin[j] = ncp;
      nops += job.taps * nkb * ((job.flags & E_ATTN) ? 3u : 1u);
      ncp += (mb + CHUNK - 1) / CHUNK;
    }
    chunk_begin[n_rj] = ncp;
    while (ncp % STAGES) {
      ops[nops++] = make_uint4(0, 0, 0, (1u << 13) | (1u << 19) | ((ncp % STAGES) << 16));
      chunk_tab[ncp++] = make_uint2(p.jobs[0].a_off, 16u);
    }
    op_begin[n_rj] = (uint16_t)nops;
    op_begin[MAXRJ] = (uint16_t)ncp;
  }
  __syncthreads();
  if (tid < n_rj) {
    const int j = tid;
    const TcJob& job = p.jobs[j];
    const uint32_t ring_a = smem_u32(smem + SM_RING), a_base = smem_u32(smem + SM_A);
    {
      // x: flags | ch << 16; y: FiLM offset | g2 << 16; z: bias | gamma << 16; w: beta | g1 << 16 (float offsets in SM_PAR)
      uint32_t o[5];
      for (int k = 0; k < 5; ++k) o[k] = (uint32_t)max((int)s_ptab[j * 5 + k], 0);
      reinterpret_cast<uint4*>(smem + SM_JD)[j] = make_uint4((uint32_t)job.flags | ((uint32_t)job.ch << 16),
                                                            (uint32_t)max(job.o_film, 0) | (o[4] << 16), o[0] | (o[1] << 16),
                                                            o[2] | (o[3] << 16));
    }
    const uint32_t a_swb = job.a_swb, blk = a_swb << 7, nkb = a_swb == 128 ? (uint32_t)job.kpt >> 6 : 1u;
    const uint32_t mb = job.mtiles * job.taps * nkb * blk;            // bytes of the main blocks
    const uint32_t w_hi = ((8u * a_swb) >> 4) | (1u << 14) |
                          ((a_swb == 128 ? (uint32_t)SW_128 : a_swb == 64 ? (uint32_t)SW_64 : (uint32_t)SW_32) << 29);
    // the image stores a job's blocks in descending size: FiLM tiles (emb 64: 16 KB) come first when they are larger
    // than the main blocks (see film_first in sampler_tc.cu)
    const uint32_t f_bytes = 128u * swb_for(pad16(L == 4 ? 16 : 64));
    const uint32_t main_off = (job.film_tiles && f_bytes > blk) ? job.film_tiles * f_bytes : 0u;
    const uint32_t chunk_base = chunk_begin[j];
    for (uint32_t off = 0, c = chunk_base; off < mb; off += CHUNK, ++c)
      chunk_tab[c] = make_uint2(job.a_off + main_off + off, min((uint32_t)CHUNK, mb - off));
    uint32_t nops = op_begin[j];
    const uint32_t n_main = (job.flags & E_ATTN) ? 128u : (uint32_t)((job.ch + 15) & ~15);
    uint32_t last_chunk = 0xffffffffu;
    bool first = true;
    auto emit = [&](uint32_t a_addr, uint32_t off, uint32_t col, uint32_t ks, uint32_t acc, uint32_t nn) {
      const uint32_t ci = chunk_base + off / CHUNK, stage = ci % STAGES;
      const uint32_t b_addr = ring_a + stage * CHUNK + (off % CHUNK);
      const uint32_t w = col | (ks << 9) | (acc << 12) | ((ci != last_chunk ? 1u : 0u) << 13) | ((first ? 1u : 0u) << 15) |
                         (stage << 16) | ((nn >> 4) << 20);
      ops[nops++] = make_uint4(0x10000u | (a_addr >> 4), 0x10000u | (b_addr >> 4), w_hi, w);
      last_chunk = ci;
      first = false;
    };
    for (uint32_t tap = 0; tap < job.taps; ++tap)
      for (uint32_t kb = 0; kb < nkb; ++kb) {
        const uint32_t tsel = job.taps == 3 ? tap : 1u;
        const uint32_t a_addr = a_base + kb * SLAB + (32 + ((int)tsel - 1) * NS) * 128;      // a tap = a shift by one position = NS rows
        const uint32_t boff = (tap * nkb + kb) * job.mtiles * blk;
        const uint32_t acc = (tap | kb) != 0 ? 1u : 0u;
        if (job.flags & E_ATTN) {
          for (uint32_t t = 0; t < 3; ++t) emit(a_addr, boff + t * blk, T_ACC + t * 128, a_swb >> 5, acc, 128u);
        } else {
          emit(a_addr, boff, T_ACC, a_swb >> 5, acc, job.mtiles == 2 ? 256u : n_main);
        }
      }
  }
  // every per-channel parameter of the network, once (step-invariant): one (job, parameter) pair per warp at a time
  for (int jk = wid; jk < n_rj * 5; jk += NTHREADS / 32) {
    const int j = jk / 5, k = jk - 5 * j;
    const TcJob& job = p.jobs[j];
    const int o = s_ptab[jk];
    if (o < 0) continue;
    const int src = k == 0 ? job.o_bias : k == 1 ? job.o_gamma : k == 2 ? job.o_beta : k == 3 ? job.o_g : job.o_g2;
    for (int i = lane; i < job.ch; i += 32) s_par[o + i] = __ldg(W + src + i);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.prof && blockIdx.x == 0 && tid == 0) p.prof[1] = clock64();
  const uint32_t cps = op_begin[MAXRJ];
  const int wid_u = __shfl_sync(0xffffffffu, wid, 0);
  if (wid_u >= 16) {
    // setmaxnreg only redistributes the CTA's own allocation (640 x 96): the 4 service warps give up 64 registers each,
    // exactly what the 16 epilogue warps gain (16 each)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (wid_u == 16) {
      // =========================== weight ring producer ===========================
      uint32_t used = 0, par = 0;
#pragma unroll 1
      for (int step = 0; step < n_steps; ++step)
#pragma unroll 1
        for (uint32_t ci = 0; ci < cps; ++ci) {
          const uint32_t s = ci % STAGES;
          if ((used >> s) & 1u)
            while (!mbar_try_wait(&empty[s], ((par >> s) & 1u) ^ 1u)) __nanosleep(64);     // off the critical path: back off
          used |= 1u << s;
          par ^= 1u << s;
          const uint2 c = chunk_tab[ci];
          bulk_g2s_elect(smem + SM_RING + s * CHUNK, p.pack + c.x, c.y, &full[s]);
        }
    } else if (wid_u == 17) {
      // =========================== UMMA issuer ===========================
      const uint32_t a_hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SW_128 << 29);
      uint32_t full_par = 0, jobn = 0;
#pragma unroll 1
      for (int step = 0; step < n_steps; ++step)
#pragma unroll 1
        for (int j = 0; j < n_rj; ++j, ++jobn) {
          const uint32_t o0 = op_begin[j], o_end = op_begin[j + 1];      // the last job also walks the ring padding
          mbar_wait(b_ready, jobn & 1);
          tc_fence_after();
          uint32_t prev_stage = 0;
#pragma unroll 1
          for (uint32_t i = o0; i < o_end; ++i) {
            uint4 op = ops[i];
            op.x = __shfl_sync(0xffffffffu, op.x, 0); op.y = __shfl_sync(0xffffffffu, op.y, 0);
            op.z = __shfl_sync(0xffffffffu, op.z, 0); op.w = __shfl_sync(0xffffffffu, op.w, 0);
            const uint32_t stage = (op.w >> 16) & 7u;
            if (op.w & (1u << 13)) {
              if (!(op.w & (1u << 15))) umma_commit_elect(&empty[prev_stage]);
              mbar_wait(&full[stage], (full_par >> stage) & 1u);
              full_par ^= 1u << stage;
              tc_fence_after();
              prev_stage = stage;
            }
            if (op.w & (1u << 19)) continue;
            const uint64_t ad = ((uint64_t)a_hi << 32) | op.x;
            const uint64_t bd = ((uint64_t)op.z << 32) | op.y;
            const uint32_t d = tmem_base + (op.w & 0x1FFu), acc = (op.w >> 12) & 1u, ks = (op.w >> 9) & 7u;
            const uint32_t idesc = idesc_bf16(128, (int)(((op.w >> 20) & 31u) << 4));
            if (ks == 4) umma_bf16_block_elect<4>(d, ad, bd, idesc, acc);
            else if (ks == 2) umma_bf16_block_elect<2>(d, ad, bd, idesc, acc);
            else umma_bf16_block_elect<1>(d, ad, bd, idesc, acc);
          }
          umma_commit_elect(&empty[prev_stage]);
          umma_commit_elect(acc_ready);
        }
    } else if (wid_u == 18) {
      // =========================== accumulator waiter ===========================
      uint32_t jobn = 0;
#pragma unroll 1
      for (int step = 0; step < n_steps; ++step)
#pragma unroll 1
        for (int j = 0; j < n_rj; ++j, ++jobn) {
          mbar_wait(acc_ready, jobn & 1);
          tc_fence_after();
          tc_fence_before();
          bar_job_arrive();
        }
    }
  } else {
    // =========================== epilogue warps ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    Ep e;
    e.g = wid >> 2; e.q = wid & 3; e.lane = lane; e.row = e.q * 32 + lane;
    e.pos = e.row / NS; e.s = e.row % NS;
    e.tmem = tmem_base + ((uint32_t)(e.q * 32) << 16);
    e.smem = smem;
    const bool dbg = p.prof != nullptr && blockIdx.x == 0;
    auto handoff = [&]() {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ready);
    };
    // FiLM table row of this thread's sample: (object, step) in the sampler, the sample itself in a single evaluation,
    // the object in the decoder
    const int smp = min(cta_s0 + e.s, p.n - 1);
    const size_t film_row0 = ((p.mode == 0) ? (size_t)(smp / p.gpo) * n_steps : (p.mode == 2) ? (size_t)(smp / p.gpo) : (size_t)smp) *
                             p.film_stride;
#pragma unroll 1
    for (int step = 0; step < n_steps; ++step) {
      const float* film_step = p.film + (film_row0 + (size_t)step * p.film_stride);
      if (p.mode == 2) {
        if (step > 0) {            // next sample group of this persistent decoder CTA (warp-group 0 = threads 0 .. 127)
          cta_s0 = ((int)blockIdx.x + step * (int)gridDim.x) * NS;
          bar_all();               // the heads of the previous group have read s_x
          if (e.g == 0) {
            load_input();
            bar_wg(0);             // (sample, position) readers below are other threads than the writers
          }
        }
        film_step = p.film + (size_t)(min(cta_s0 + e.s, p.n - 1) / p.gpo) * p.film_stride;
      }
      // ---- network input: the state, or (evaluation programs) c_in * x_in with the stochastic churn added
      const bool edm = p.mode == 0 && p.sched_kind == GLDM_SCHED_EDM;
      float* s_y = s_x + 128;
      float* s_z = s_x + 256;
      float* s_xin = s_x + 384;
      float* s_sc = s_x + 512;
      if (e.g == 0) {
        float v = s_x[e.s * L + e.pos];
        if (edm) {
          const float* cf = p.coef + (size_t)step * kEvalRow;
          const int slot = (int)__ldg(cf + 8);
          float z = 0.f;
          if (slot >= 0 && cta_s0 + e.s < p.n)
            z = p.noise ? __ldg(p.noise + ((size_t)slot * p.n + cta_s0 + e.s) * L + e.pos)
                        : philox_normal(p.seed, (unsigned)(cta_s0 + e.s), (unsigned)slot, (unsigned)e.pos);
          float c4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) c4[i] = __ldg(cf + i);
          v = eval_input(c4, v, z);
          s_xin[e.s * L + e.pos] = v;
          v = __fmul_rn(c4[0], v);
        }
        s_sc[e.s * L + e.pos] = v;
        bar_wg(0);
      }
      // ---- init_conv: Conv1d(1 -> dim, k7, p3) on the input -> residual stream and operand.  dim = 4 (L = 4): warp-group
      //      0 alone; dim = 16 (L = 16): four channels per warp-group
      if (L == 16) bar_all();          // every warp-group reads s_sc
      if (L == 16 || e.g == 0) {
        const int c0 = (L == 16) ? 4 * e.g : 0;
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float a = __ldg(W + lay.init_b + c0 + c);
#pragma unroll
          for (int t = 0; t < 7; ++t) {
            const int ll = e.pos + t - 3;
            if (ll >= 0 && ll < L) a = fmaf(__ldg(W + lay.init_w + (c0 + c) * 7 + t), s_sc[e.s * L + ll], a);
          }
          v[c] = a;
        }
        st_res<4>(e, false, c0, v);
        tmem_st_wait();
        st_operand<4>(e, c0, v);
      }
      handoff();
#pragma unroll 1
      for (int j = 0; j < n_rj; ++j) {
        const uint4 jd = reinterpret_cast<const uint4*>(smem + SM_JD)[j];
        const int flags = (int)(jd.x & 0xffffu), ch = (int)(jd.x >> 16);
        const float* film = film_step + (jd.y & 0xffffu);
        // ---- this job's channel parameters (resident in shared memory); pull its FiLM vectors into L1 while its UMMAs run
        if (!(flags & E_ATTN)) {
          e.po[0] = jd.z; e.po[1] = jd.w; e.po[2] = jd.y >> 16;
          if (flags & E_FILM) {
            const int cw = ch >= 32 ? ch >> 2 : 4, c_lo = ch >= 32 ? e.g * cw : ch == 16 ? 4 * e.g : 0;
            prefetch_l1(film + c_lo);
            prefetch_l1(film + c_lo + cw - 1);
            prefetch_l1(film + ch + c_lo);
            prefetch_l1(film + ch + c_lo + cw - 1);
          }
        }
        const bool rec = dbg && tid == 0 && step == (n_steps > 1 ? 1 : 0);
        if (rec) p.prof[64 + 8 * j] = clock64();
        bar_job_sync();
        tc_fence_after();
        if (rec) p.prof[64 + 8 * j + 1] = clock64();
        if (rec) p.prof[64 + 8 * j + 2] = clock64();
        if (dbg && step == 0 && p.prof[8] == (long long)j + 1) {
          // development aid: raw TMEM image [128 rows][512 columns] of CTA 0 when job j's accumulator is ready
          float* dump = reinterpret_cast<float*>(p.prof + 16);
          for (int c = e.g * 128; c < e.g * 128 + 128; c += 8) {
            float t[8];
            ld_cols<8>(e, c, t);
#pragma unroll
            for (int q = 0; q < 8; ++q) dump[(size_t)e.row * 512 + c + q] = t[q];
          }
        }
        if (flags & E_ATTN) {
          if (L == 4) attention_epilogue(e, rec ? p.prof + 64 + 8 * j : nullptr);
          else attention16_epilogue(e);
        } else if (ch >= 32 && ch <= 128) {
          reg_epilogue<L>(e, flags, ch, film, rec ? p.prof + 64 + 8 * j : nullptr);
        } else if (ch < 32) {
          if (L == 4) narrow_epilogue(e, flags, film);
          else quad_epilogue<L>(e, flags, film);
        } else {
          const float dot = wide_epilogue<L>(e, flags, film);
          if (flags & E_FINAL) {
            // ======== final_conv (1x1 -> 1 channel) + scheduler update of x[sample][position]
            e.xl()[e.row * 4 + e.g] = make_float2(dot, 0.f);
            bar_all();
            if (e.g == 0) {
              float eps = __ldg(W + lay.fc_b);
#pragma unroll
              for (int w = 0; w < 4; ++w) eps += e.xl()[e.row * 4 + w].x;
              const int l = e.pos, s = e.s;
              const bool ok = cta_s0 + s < p.n;
              if (edm) {
                float c16[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) c16[i] = __ldg(p.coef + (size_t)step * kEvalRow + i);
                float x = s_x[s * L + l], y = s_y[s * L + l], z = s_z[s * L + l];
                eval_update(c16, s_xin[s * L + l], eps, p.clip, x, y, z);
                s_x[s * L + l] = x; s_y[s * L + l] = y; s_z[s * L + l] = z;
                const int slot = (int)c16[9];
                if (p.x_all && ok && slot >= 0) p.x_all[((size_t)slot * p.n + cta_s0 + s) * L + l] = x;
              } else if (p.mode == 0) {
                const float* cf = p.coef + (size_t)step * 8;
                const float x = s_x[s * L + l];
                float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(__ldg(cf + 0), eps)), __ldg(cf + 1));
                if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
                float prev;
                if (p.sched_kind == GLDM_SCHED_DDPM) {
                  prev = __fadd_rn(__fmul_rn(__ldg(cf + 2), x0), __fmul_rn(__ldg(cf + 3), x));
                  const float sg = __ldg(cf + 4);
                  if (sg > 0.f && ok) {
                    const float z = p.noise ? __ldg(p.noise + ((size_t)step * p.n + cta_s0 + s) * L + l)
                                            : philox_normal(p.seed, (unsigned)(cta_s0 + s), (unsigned)step, (unsigned)l);
                    prev = __fadd_rn(prev, __fmul_rn(sg, z));
                  }
                } else {
                  prev = __fadd_rn(__fmul_rn(__ldg(cf + 2), x0), __fmul_rn(__ldg(cf + 3), eps));
                }
                s_x[s * L + l] = prev;
                if (p.x_all && ok) p.x_all[((size_t)(step + 1) * p.n + cta_s0 + s) * L + l] = prev;
              } else {
                s_x[s * L + l] = eps;
              }
            }
            bar_all();
          }
        }
        if (rec) p.prof[64 + 8 * j + 6] = clock64();
        if (j + 1 < n_rj) handoff();
      }
      if (p.mode == 2 && tid < NS * 7) {
        // decoder heads: tmrp = Linear(L -> 6), class_logits = Linear(L -> 1)   (grasp_vae.py:428-430); s_x was written by
        // warp-group 0 behind the bar_all of the final job
        const float* hw = p.head + L * p.D + L;   // tmrp_w [6][L], tmrp_b [6], cls_w [L], cls_b [1]
        const int s = tid / 7, o = tid - s * 7;
        if (cta_s0 + s < p.n) {
          const float* x = s_x + s * L;
          if (o < 6) {
            float a = __ldg(hw + 6 * L + o);
            for (int l = 0; l < L; ++l) a = fmaf(__ldg(hw + o * L + l), x[l], a);
            p.tmrp[(size_t)(cta_s0 + s) * 6 + o] = a;
          } else {
            float a = __ldg(hw + 6 * L + 6 + L);
            for (int l = 0; l < L; ++l) a = fmaf(__ldg(hw + 6 * L + 6 + l), x[l], a);
            p.logit[cta_s0 + s] = a;
          }
        }
      }
    }
    if (p.mode != 2 && e.g == 0 && cta_s0 + e.s < p.n) p.x_out[(size_t)(cta_s0 + e.s) * L + e.pos] = s_x[e.s * L + e.pos];
  }
  tc_fence_before();
  __syncthreads();
  if (p.prof && blockIdx.x == 0 && tid == 0) p.prof[2] = clock64();
  if (wid == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace rows
