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
// TMEM columns: [0,256) accumulator (q | k | v = [0,384) for the attention job), FiLM scale [128,256) and shift
// [256,384) (FiLM layers are <= 128 wide; the 256-wide first conv of the final block runs as two 128-column halves),
// [384,512) residual stream: fp32 for widths <= 128, packed bf16 pairs for the 256-wide final block.
#pragma once

namespace rows {
constexpr int NS = 32;                          // samples per CTA
constexpr int NEPI = 512;                       // 16 epilogue warps
constexpr int NTHREADS = 640;                   // + producer, issuer, two idle register donors
constexpr int SLAB = 192 * 128;                 // 64 channels x (32 halo + 128 + 32 halo) rows
constexpr int CHUNK = stc::CHUNK, STAGES = 2;
constexpr int MAXRJ = 40, MAXOPS = 400, MAXCHUNKS = 256;
constexpr uint32_t T_ACC = 0, T_FS = 128, T_FH = 256, T_RES = 384;
// shared memory map (from a 1024-aligned base)
constexpr int SM_A = 0;
constexpr int SM_U = SM_A + 4 * SLAB;                     // FiLM operand u: 128 rows x 128 B (K = 16 used)
constexpr int SM_RING = SM_U + 128 * 128;
constexpr int SM_PAR = SM_RING + STAGES * CHUNK;          // per-job channel parameters [7][256] floats
constexpr int SM_XG = SM_PAR + 7 * 256 * 4;               // GroupNorm exchange [4 groups][4 positions][32] float2
constexpr int SM_XL = SM_XG + 4 * 4 * 32 * 8;             // row exchange [128 rows][4 warp-groups] float2
constexpr int SM_INEMB = SM_XL + 128 * 4 * 8;             // [32][3][16] floats
constexpr int SM_X = SM_INEMB + 32 * 3 * 16 * 4;          // state [32][4]
constexpr int SM_BAR = SM_X + 32 * 4 * 4;
constexpr int SM_CHUNKS = SM_BAR + 256;
constexpr int SM_RJ = SM_CHUNKS + MAXCHUNKS * 8;          // row-job table
constexpr int SM_OPS = SM_RJ + MAXRJ * 8;
constexpr int SM_OPBEG = SM_OPS + MAXOPS * 16;
constexpr int SM_TOTAL = SM_OPBEG + (MAXRJ + 2) * 2 + 16;
static_assert(SM_TOTAL + 1024 <= 232448, "shared memory budget");

// kind 0: a whole layer.  The 256-wide FiLM layer (first conv of the final block) does not fit TMEM next to its FiLM
// vectors (256 + 2 * 256 columns) and its epilogue overwrites the operand it is computed from, so it runs as five
// row-jobs: kind 1 = the convolution (N = 256) + GroupNorm, normalised values kept in the accumulator columns;
// kind 2 / 3 (h = 0, 1) = FiLM scale / shift of channel half h into columns [256, 384), applied in place; kind 3 ends
// with SiLU and writes the operand.
struct RJob { int16_t job; int8_t kind, h; int16_t ch, wgs; };
constexpr int E_INPLACE = 0x4000;      // internal: GroupNorm result back into the accumulator columns, nothing else

__device__ __forceinline__ void bar_wg(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 5, 512;" ::: "memory"); }

struct Ep {
  uint32_t tmem;        // TMEM base + this warp's lane quarter
  int g, pos, s, row;   // warp-group, position, sample (lane), row = pos * 32 + s
  uint8_t* smem;
  const float* par;     // staged channel parameters [7][256]: bias, gamma, beta, film scale bias (+1), film shift bias, g1, g2
  float2* xg;
  float2* xl;
};

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
  const float* p = e.par + k * 256 + c;
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

// generic layer epilogue.  NV = 8: warp-group g owns channels [g*cw, (g+1)*cw), cw = ch / wgs >= 8, one GroupNorm group;
// NV = 4: the 4-channel first stage, warp-group 0 owns all 4 channels = 4 GroupNorm groups of one channel.
// Returns the final-conv partial dot product (E_FINAL) of this thread's channels.
template <int NV>
__device__ __forceinline__ float generic_epilogue(const Ep& e, int flags, int ch, int ch_total, int ch_off, int wgs) {
  const bool active = e.g < wgs;
  const int cw = (NV == 8) ? ch / wgs : 4;
  const int c_lo = (NV == 8) ? e.g * cw : 0, c_hi = c_lo + cw;
  const bool packed = ch_total > 128;
  float mean[NV == 4 ? 4 : 1], rstd[NV == 4 ? 4 : 1];
  float lmean = 0.f, lrstd = 1.f;
  // ---- pass 1: statistics of (acc + bias)
  if (flags & (E_GN | E_LN)) {
    float s[NV == 4 ? 4 : 1], q[NV == 4 ? 4 : 1];
#pragma unroll
    for (int i = 0; i < (NV == 4 ? 4 : 1); ++i) { s[i] = 0.f; q[i] = 0.f; }
    if (active) {
      for (int c = c_lo; c < c_hi; c += NV) {
        float x[NV], b[NV];
        ld_cols<NV>(e, T_ACC + c, x);
        ld_par<NV>(e, 0, c, b);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const float t = x[j] + b[j];
          const int gi = (NV == 4 && (flags & E_GN)) ? j : 0;
          s[gi] += t;
          q[gi] = fmaf(t, t, q[gi]);
        }
      }
    }
    if (flags & E_GN) {
      // sums over the 4 positions of a sample: the 4 warps of this warp-group
      if (active) {
#pragma unroll
        for (int i = 0; i < (NV == 4 ? 4 : 1); ++i) e.xg[((NV == 4 ? i : e.g) * 4 + e.pos) * 32 + e.s] = make_float2(s[i], q[i]);
      }
      bar_wg(e.g);
      if (active) {
        const float inv = 1.0f / (float)((NV == 4 ? 1 : cw) * 4);
#pragma unroll
        for (int i = 0; i < (NV == 4 ? 4 : 1); ++i) {
          float ts = 0.f, tq = 0.f;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float2 t = e.xg[((NV == 4 ? i : e.g) * 4 + p) * 32 + e.s];
            ts += t.x; tq += t.y;
          }
          const float m = ts * inv;
          mean[i] = m;
          rstd[i] = rsqrtf(fmaxf(tq * inv - m * m, 0.f) + 1e-5f);
        }
      }
    } else {
      // channel LayerNorm of the row: sums over the warp-groups
      e.xl[e.row * 4 + e.g] = active ? make_float2(s[0], q[0]) : make_float2(0.f, 0.f);
      bar_all();
      float ts = 0.f, tq = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) { const float2 t = e.xl[e.row * 4 + w]; ts += t.x; tq += t.y; }
      const float inv = 1.0f / (float)ch_total;
      lmean = ts * inv;
      lrstd = rsqrtf(fmaxf(tq * inv - lmean * lmean, 0.f) + 1e-5f);
      bar_all();            // xl is reused by E_LNNEXT / E_FINAL below
    }
  }
  // ---- pass 2
  float ls = 0.f, lq = 0.f, dot = 0.f;
  if (active) {
    for (int c = c_lo; c < c_hi; c += NV) {
      float v[NV], t0[NV], t1[NV];
      ld_cols<NV>(e, T_ACC + c, v);
      ld_par<NV>(e, 0, c, t0);
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] += t0[j];
      if (flags & E_GN) {
        ld_par<NV>(e, 1, c, t0);
        ld_par<NV>(e, 2, c, t1);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int gi = NV == 4 ? j : 0;
          const float a = rstd[gi] * t0[j];
          v[j] = fmaf(v[j] - mean[gi], a, t1[j]);
        }
      }
      if (flags & E_INPLACE) {
        st_cols<NV>(e, T_ACC + c, v);
        continue;
      }
      if (flags & E_FILM) {
        float fs[NV], fh[NV];
        ld_cols<NV>(e, T_FS + c, fs);
        ld_cols<NV>(e, T_FH + c, fh);
        ld_par<NV>(e, 3, c, t0);
        ld_par<NV>(e, 4, c, t1);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = fmaf(v[j], fs[j] + t0[j], fh[j] + t1[j]);
      }
      if (flags & E_SILU) {
#pragma unroll
        for (int j = 0; j < NV; j += 2) {
          const float2 r = silu_fast2(make_float2(v[j], v[j + 1]));
          v[j] = r.x; v[j + 1] = r.y;
        }
      }
      if (flags & E_LN) {
        ld_par<NV>(e, 5, c, t0);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = (v[j] - lmean) * lrstd * t0[j];
      }
      if (flags & E_ADDRES) {
        ld_res<NV>(e, packed, ch_off + c, t0);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] += t0[j];
      }
      if (flags & E_STORERES) st_res<NV>(e, packed, ch_off + c, v);
      if (flags & E_LNNEXT) {
#pragma unroll
        for (int j = 0; j < NV; ++j) { ls += v[j]; lq = fmaf(v[j], v[j], lq); }
      } else if (flags & E_FINAL) {
        ld_par<NV>(e, 6, c, t0);
#pragma unroll
        for (int j = 0; j < NV; ++j) dot = fmaf(t0[j], v[j], dot);
      } else {
        st_operand<NV>(e, ch_off + c, v);
      }
    }
    if (flags & (E_STORERES | E_INPLACE)) tmem_st_wait();
  }
  // ---- pass 3: PreNorm of the attention: operand = LayerNorm(result) * g2, result re-read from the residual stream
  if (flags & E_LNNEXT) {
    e.xl[e.row * 4 + e.g] = active ? make_float2(ls, lq) : make_float2(0.f, 0.f);
    bar_all();
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { const float2 t = e.xl[e.row * 4 + w]; ts += t.x; tq += t.y; }
    const float inv = 1.0f / (float)ch_total;
    const float m = ts * inv, r = rsqrtf(fmaxf(tq * inv - m * m, 0.f) + 1e-5f);
    if (active) {
      for (int c = c_lo; c < c_hi; c += NV) {
        float v[NV], g2[NV];
        ld_res<NV>(e, packed, ch_off + c, v);
        ld_par<NV>(e, 6, c, g2);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = (v[j] - m) * r * g2[j];
        st_operand<NV>(e, ch_off + c, v);
      }
    }
  }
  return dot;
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
__device__ __forceinline__ void attention_epilogue(const Ep& e) {
  uint8_t* X = e.smem + SM_A + e.g * SLAB;        // rows 32..159 of this head's slab (the halo rows are not touched)
  const int h = e.g, t = e.pos, s = e.s;
  {
    float qs[32], kv[32];
    uint32_t rq[4][8], rk[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ld_issue<8>(e, T_ACC + h * 32 + i * 8, rq[i]);
      ld_issue<8>(e, T_ACC + 128 + h * 32 + i * 8, rk[i]);
    }
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float tmp[8];
      ld_use<8>(rq[i], tmp);
#pragma unroll
      for (int j = 0; j < 8; ++j) qs[i * 8 + j] = tmp[j];
      ld_use<8>(rk[i], tmp);
#pragma unroll
      for (int j = 0; j < 8; ++j) kv[i * 8 + j] = tmp[j];
    }
    float m = qs[0];
#pragma unroll
    for (int d = 1; d < 32; ++d) m = fmaxf(m, qs[d]);
    float sum = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) { qs[d] = __expf(qs[d] - m); sum += qs[d]; }
    const float sc = __fdividef(0.17677669529663687f, sum);
#pragma unroll
    for (int d = 0; d < 32; ++d) qs[d] *= sc;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *xslot(X, 0, t, s, j) = pack8(&qs[8 * j]);      // q^ of this position
      *xslot(X, 1, t, s, j) = pack8(&kv[8 * j]);      // raw k of this position
    }
  }
  bar_wg(h);
  float Ap[16];                                       // partial A[n][n'] over this thread's quarter of d
  {
    float q[4][8], k[4][8];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      unpack8(*xslot(X, 0, pp, s, t), q[pp]);
      unpack8(*xslot(X, 1, pp, s, t), k[pp]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float m = fmaxf(fmaxf(k[0][i], k[1][i]), fmaxf(k[2][i], k[3][i]));
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) k[pp][i] = __expf(k[pp][i] - m);
      const float zi = __fdividef(1.0f, (k[0][i] + k[1][i]) + (k[2][i] + k[3][i]));
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) k[pp][i] *= zi;
    }
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int n1 = 0; n1 < 4; ++n1) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a = fmaf(q[n][i], k[n1][i], a);
        Ap[n * 4 + n1] = a;
      }
  }
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
      float tmp[8];
      ld_use<8>(rv[i], tmp);
      *xslot(X, 1, t, s, i) = pack8(tmp);             // v of this position
    }
  }
  bar_wg(h);
  float o[4][8];                                      // out[e in quarter t][position n]
  {
    float A[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 acc = *reinterpret_cast<const float4*>(xslot(X, 0, 0, s, j));
#pragma unroll
      for (int pp = 1; pp < 4; ++pp) {
        const float4 x = *reinterpret_cast<const float4*>(xslot(X, 0, pp, s, j));
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
      A[4 * j] = acc.x; A[4 * j + 1] = acc.y; A[4 * j + 2] = acc.z; A[4 * j + 3] = acc.w;
    }
    float v[4][8];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) unpack8(*xslot(X, 1, pp, s, t), v[pp]);
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = A[n * 4] * v[0][i];
#pragma unroll
        for (int n1 = 1; n1 < 4; ++n1) a = fmaf(A[n * 4 + n1], v[n1][i], a);
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
      *reinterpret_cast<uint4*>(e.smem + SM_A + (chan >> 6) * SLAB + R * 128 + ((((chan & 63) >> 3) ^ (R & 7)) << 4)) = pack8(o[n]);
    }
  }
}

// FiLM rounds of the split 256-wide layer: channel half h (128 channels, 32 per warp-group); FiLM vector in [256, 384)
__device__ __forceinline__ void film_round(const Ep& e, int kind, int h) {
  for (int c = e.g * 32; c < e.g * 32 + 32; c += 8) {
    float v[8], f[8], b[8];
    ld_cols<8>(e, T_ACC + 128 * h + c, v);
    ld_cols<8>(e, T_FH + c, f);
    ld_par<8>(e, kind == 2 ? 3 : 4, c, b);
    if (kind == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= f[j] + b[j];
      st_cols<8>(e, T_ACC + 128 * h + c, v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float2 r = silu_fast2(make_float2(v[j] + f[j] + b[j], v[j + 1] + f[j + 1] + b[j + 1]));
        v[j] = r.x; v[j + 1] = r.y;
      }
      st_operand<8>(e, 128 * h + c, v);
    }
  }
  if (kind == 2) tmem_st_wait();
}

__global__ void __launch_bounds__(NTHREADS, 1) resnet_rows_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* full = bars;          // [STAGES]
  uint64_t* empty = bars + 4;     // [STAGES]
  uint64_t* b_ready = bars + 8;
  uint64_t* acc_ready = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  RJob* rj = reinterpret_cast<RJob*>(smem + SM_RJ);
  uint2* chunk_tab = reinterpret_cast<uint2*>(smem + SM_CHUNKS);
  uint4* ops = reinterpret_cast<uint4*>(smem + SM_OPS);
  uint16_t* op_begin = reinterpret_cast<uint16_t*>(smem + SM_OPBEG);
  float* s_par = reinterpret_cast<float*>(smem + SM_PAR);
  float* s_inemb = reinterpret_cast<float*>(smem + SM_INEMB);
  float* s_x = reinterpret_cast<float*>(smem + SM_X);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int cta_s0 = blockIdx.x * NS;
  const GldmResNetCfg& cfg = p.cfg;
  const ResNetLayout& lay = p.lay;
  const float* W = p.W;
  const int R = cfg.cond_ch;
  constexpr int EMB = 16, L = 4;
  const int n_steps = (p.mode == 0) ? p.n_steps : 1;

  // ---- one-time setup
  for (int i = tid; i < SM_RING / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(b_ready, NEPI / 32);
    mbar_init(acc_ready, 1);
    fence_barrier_init();
  }
  if (wid == 0) tmem_alloc<512>(tmem_slot);
  for (int idx = tid; idx < NS * R * EMB; idx += NTHREADS) {
    const int e = idx % EMB, r = (idx / EMB) % R, s = idx / (EMB * R);
    float a = 0.f;
    if (cta_s0 + s < p.n) {
      const int obj = (cta_s0 + s) / p.gpo;
      const float* z = p.z_cond + ((size_t)obj * R + r) * cfg.cond_dim;
      const float* w = W + lay.in_w + (size_t)e * cfg.cond_dim;
      a = __ldg(W + lay.in_b + e);
      for (int j = 0; j < cfg.cond_dim; ++j) a = fmaf(__ldg(w + j), __ldg(z + j), a);
      a = a / (1.0f + expf(-a));
    }
    s_inemb[(s * 3 + r) * EMB + e] = a;
  }
  if (tid < NS * L) {
    const int s = tid / L, l = tid % L;
    float v = 0.f;
    if (cta_s0 + s < p.n) v = __ldg(p.x_in + (size_t)(cta_s0 + s) * L + l);
    s_x[s * L + l] = v;
    if (p.mode == 0 && p.x_all && cta_s0 + s < p.n) p.x_all[(size_t)(cta_s0 + s) * L + l] = v;
  }
  // ---- row-job, weight-chunk and UMMA op tables of one step
  // op.x / op.y: low words of the A (activation) / B (weight) descriptors; op.z: high word of the B descriptor;
  // op.w: [0,9) TMEM column, [9,12) K steps, 12 accumulate, 13 first use of a ring chunk, 15 first op of the job,
  //       [16,19) ring stage, 19 ring padding, [20,25) N / 16
  if (tid == 0) {
    const uint32_t ring_a = smem_u32(smem + SM_RING), a_base = smem_u32(smem + SM_A), u_base = smem_u32(smem + SM_U);
    const uint32_t f_swb = swb_for(pad16(EMB)), f_bytes = 128u * f_swb;
    const uint32_t f_hi = ((8u * f_swb) >> 4) | (1u << 14) | ((f_swb == 128 ? (uint32_t)SW_128 : (uint32_t)SW_32) << 29);
    uint32_t nops = 0, chunk_base = 0, ncp = 0, nrj = 0;
    for (int j = 0; j < p.n_jobs; ++j) {
      const TcJob& job = p.jobs[j];
      const bool split = job.mtiles == 2 && job.film_tiles;
      const int n_sub = split ? 5 : 1;
      const uint32_t a_swb = job.a_swb, blk = a_swb << 7, nkb = a_swb == 128 ? (uint32_t)job.kpt >> 6 : 1u;
      const uint32_t mb = job.mtiles * job.taps * nkb * blk;            // bytes of the main blocks
      const uint32_t w_hi = ((8u * a_swb) >> 4) | (1u << 14) |
                            ((a_swb == 128 ? (uint32_t)SW_128 : a_swb == 64 ? (uint32_t)SW_64 : (uint32_t)SW_32) << 29);
      for (int sub = 0; sub < n_sub; ++sub) {
        RJob r;
        r.job = (int16_t)j;
        r.kind = (int8_t)(!split ? 0 : sub == 0 ? 1 : (sub & 1) ? 2 : 3);
        r.h = (int8_t)(sub >= 3 ? 1 : 0);
        r.ch = (int16_t)(r.kind >= 2 ? 128 : job.ch);
        r.wgs = (int16_t)((job.flags & E_ATTN) ? 4 : job.ch == 4 ? 1 : 4);
        rj[nrj] = r;
        // the part of the image this row-job streams
        const uint32_t s_off = r.kind >= 2 ? mb : 0u, s_bytes = r.kind == 1 ? mb : r.kind >= 2 ? job.bytes - mb : job.bytes;
        for (uint32_t off = 0; off < s_bytes; off += CHUNK)
          chunk_tab[ncp++] = make_uint2(job.a_off + s_off + off, min((uint32_t)CHUNK, s_bytes - off));
        op_begin[nrj] = (uint16_t)nops;
        const uint32_t n_main = (job.flags & E_ATTN) ? 128u : (uint32_t)((job.ch + 15) & ~15);
        uint32_t last_chunk = 0xffffffffu;
        bool first = true;
        auto emit = [&](uint32_t a_addr, uint32_t off, uint32_t hi, uint32_t col, uint32_t ks, uint32_t acc, uint32_t nn) {
          const uint32_t ci = chunk_base + off / CHUNK, stage = ci % STAGES;
          const uint32_t b_addr = ring_a + stage * CHUNK + (off % CHUNK);
          const uint32_t w = col | (ks << 9) | (acc << 12) | ((ci != last_chunk ? 1u : 0u) << 13) | ((first ? 1u : 0u) << 15) |
                             (stage << 16) | ((nn >> 4) << 20);
          ops[nops++] = make_uint4(0x10000u | (a_addr >> 4), 0x10000u | (b_addr >> 4), hi, w);
          last_chunk = ci;
          first = false;
        };
        if (r.kind <= 1) {
          for (uint32_t tap = 0; tap < job.taps; ++tap)
            for (uint32_t kb = 0; kb < nkb; ++kb) {
              const uint32_t tsel = job.taps == 3 ? tap : 1u;
              const uint32_t a_addr = a_base + kb * SLAB + tsel * 32 * 128;
              const uint32_t boff = (tap * nkb + kb) * job.mtiles * blk;
              const uint32_t acc = (tap | kb) != 0 ? 1u : 0u;
              if (job.flags & E_ATTN) {
                for (uint32_t t = 0; t < 3; ++t) emit(a_addr, boff + t * blk, w_hi, T_ACC + t * 128, a_swb >> 5, acc, 128u);
              } else {
                emit(a_addr, boff, w_hi, T_ACC, a_swb >> 5, acc, job.mtiles == 2 ? 256u : n_main);
              }
            }
          if (job.film_tiles && r.kind == 0) {
            emit(u_base, mb, f_hi, T_FS, f_swb >> 5, 0u, n_main);
            emit(u_base, mb + f_bytes, f_hi, T_FH, f_swb >> 5, 0u, n_main);
          }
        } else {
          // FiLM tiles of the split layer: [scale 0, scale 1, shift 0, shift 1]; this row-job streams only that part
          const uint32_t tile = (r.kind == 2 ? 0u : 2u) + (uint32_t)r.h;
          emit(u_base, tile * f_bytes, f_hi, T_FH, f_swb >> 5, 0u, 128u);
        }
        chunk_base += (s_bytes + CHUNK - 1) / CHUNK;
        ++nrj;
      }
    }
    while (ncp % STAGES) {
      ops[nops++] = make_uint4(0, 0, 0, (1u << 13) | (1u << 19) | ((ncp % STAGES) << 16));
      chunk_tab[ncp++] = make_uint2(p.jobs[0].a_off, 16u);
    }
    op_begin[nrj] = (uint16_t)nops;
    op_begin[MAXRJ] = (uint16_t)ncp;
    op_begin[MAXRJ + 1] = (uint16_t)nrj;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cps = op_begin[MAXRJ];
  const int n_rj = op_begin[MAXRJ + 1];
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
          if ((used >> s) & 1u) mbar_wait(&empty[s], ((par >> s) & 1u) ^ 1u);
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
    }
  } else {
    // =========================== epilogue warps ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    Ep e;
    e.g = wid >> 2; e.pos = wid & 3; e.s = lane; e.row = e.pos * 32 + lane;
    e.tmem = tmem_base + ((uint32_t)(e.pos * 32) << 16);
    e.smem = smem;
    e.par = s_par;
    e.xg = reinterpret_cast<float2*>(smem + SM_XG);
    e.xl = reinterpret_cast<float2*>(smem + SM_XL);
    uint32_t jobn = 0;
    auto handoff = [&]() {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ready);
    };
#pragma unroll 1
    for (int step = 0; step < n_steps; ++step) {
      // ---- FiLM operand u[s][e] = sum_r silu(time_emb[e] + in_emb[s][r][e]), replicated over the 4 positions
      {
        const int s = tid / EMB, ee = tid % EMB;
        const int ti = (p.mode == 0) ? step : min(cta_s0 + s, p.n - 1);
        const float te = __ldg(p.te + (size_t)ti * EMB + ee);
        float a = 0.f;
        for (int r = 0; r < R; ++r) { const float z = te + s_inemb[(s * 3 + r) * EMB + ee]; a += z / (1.0f + __expf(-z)); }
        const __nv_bfloat16 hv = __float2bfloat16(a);
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const int row = pp * 32 + s;
          *reinterpret_cast<__nv_bfloat16*>(smem + SM_U + row * 128 + (((ee >> 3) ^ (row & 7)) << 4) + (ee & 7) * 2) = hv;
        }
      }
      // ---- init_conv: Conv1d(1 -> 4, k7, p3) on the state -> residual stream and operand (warp-group 0: 4 channels)
      if (e.g == 0) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float a = __ldg(W + lay.init_b + c);
#pragma unroll
          for (int t = 0; t < 7; ++t) {
            const int ll = e.pos + t - 3;
            if (ll >= 0 && ll < 4) a = fmaf(__ldg(W + lay.init_w + c * 7 + t), s_x[e.s * 4 + ll], a);
          }
          v[c] = a;
        }
        st_res<4>(e, false, 0, v);
        tmem_st_wait();
        st_operand<4>(e, 0, v);
      }
      handoff();
#pragma unroll 1
      for (int j = 0; j < n_rj; ++j, ++jobn) {
        const RJob r = rj[j];
        const TcJob& job = p.jobs[r.job];
        const int flags = r.kind == 1 ? (E_GN | E_INPLACE) : job.flags;
        const int ch = r.ch, ch_total = job.ch, ch_off = r.kind >= 2 ? 128 * r.h : 0;
        // ---- stage the channel parameters of this job while its UMMAs run
        if (!(flags & E_ATTN)) {
          for (int i = tid; i < ch; i += NEPI) {
            const int cc = ch_off + i;
            s_par[0 * 256 + i] = job.o_bias >= 0 ? __ldg(W + job.o_bias + cc) : 0.f;
            s_par[1 * 256 + i] = job.o_gamma >= 0 ? __ldg(W + job.o_gamma + cc) : 0.f;
            s_par[2 * 256 + i] = job.o_beta >= 0 ? __ldg(W + job.o_beta + cc) : 0.f;
            s_par[3 * 256 + i] = job.o_mlpb >= 0 ? (float)R * __ldg(W + job.o_mlpb + cc) + (float)R : 1.f;   // sum_r (scale_r + 1)
            s_par[4 * 256 + i] = job.o_mlpb >= 0 ? (float)R * __ldg(W + job.o_mlpb + ch_total + cc) : 0.f;
            s_par[5 * 256 + i] = job.o_g >= 0 ? __ldg(W + job.o_g + cc) : 0.f;
            s_par[6 * 256 + i] = job.o_g2 >= 0 ? __ldg(W + job.o_g2 + cc) : 0.f;
          }
        }
        const bool rec = p.prof && blockIdx.x == 0 && tid == 0 && step == 1;
        if (rec) p.prof[64 + 8 * j] = clock64();
        mbar_wait(acc_ready, jobn & 1);
        tc_fence_after();
        if (rec) p.prof[64 + 8 * j + 1] = clock64();
        bar_all();
        if (rec) p.prof[64 + 8 * j + 2] = clock64();
        if (p.prof && blockIdx.x == 0 && p.prof[8] == (long long)j + 1 && step == 0) {
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
          attention_epilogue(e);
        } else if (r.kind >= 2) {
          film_round(e, r.kind, r.h);
        } else {
          const float dot = (ch_total == 4) ? generic_epilogue<4>(e, flags, ch, ch_total, ch_off, r.wgs)
                                            : generic_epilogue<8>(e, flags, ch, ch_total, ch_off, r.wgs);
          if (flags & E_FINAL) {
            // ======== final_conv (1x1 -> 1 channel) + scheduler update of x[sample][position]
            e.xl[e.row * 4 + e.g] = make_float2(dot, 0.f);
            bar_all();
            if (e.g == 0) {
              float eps = __ldg(W + lay.fc_b);
#pragma unroll
              for (int w = 0; w < 4; ++w) eps += e.xl[e.row * 4 + w].x;
              const int l = e.pos, s = e.s;
              const bool ok = cta_s0 + s < p.n;
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
                    const float z = p.noise ? __ldg(p.noise + ((size_t)step * p.n + cta_s0 + s) * L + l)
                                            : philox_normal(p.seed, (unsigned)(cta_s0 + s), (unsigned)step, (unsigned)l);
                    prev = __fadd_rn(prev, __fmul_rn(sg, z));
                  }
                } else {
                  prev = __fadd_rn(__fmul_rn(__ldg(cf + 2), x0), __fmul_rn(__ldg(cf + 3), eps));
                }
                s_x[s * 4 + l] = prev;
                if (p.x_all && ok) p.x_all[((size_t)(step + 1) * p.n + cta_s0 + s) * L + l] = prev;
              } else {
                s_x[s * 4 + l] = eps;
              }
            }
            bar_all();
          }
        }
        if (rec) p.prof[64 + 8 * j + 6] = clock64();
        if (j + 1 < n_rj) handoff();
      }
    }
    if (e.g == 0 && cta_s0 + e.s < p.n) p.x_out[(size_t)(cta_s0 + e.s) * L + e.pos] = s_x[e.s * 4 + e.pos];
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace rows
