// Raw parameter-blob layout of the ResNet1D family (shared by the fp32 and the tensor-core kernels) and the
// counter-based RNG of the fused samplers.
#pragma once
#include "common.cuh"

namespace gldm {

// ---------------------------------------------------------------------------------------------
// parameter layout (offsets in floats into the raw / prepared blobs; every tensor padded to 4 floats)
// ---------------------------------------------------------------------------------------------
struct RbOff { int mlp_w, mlp_b, p1_w, p1_b, n1_w, n1_b, p2_w, p2_b, n2_w, n2_b; };
struct StageOff { RbOff rb[2]; int ln_g, qkv_w, out_w, out_b, out_g, down_w, down_b; };
struct ResNetLayout {
  int init_w, init_b, tm_freq, tm_w1, tm_b1, tm_w2, tm_b2, in_w, in_b;
  StageOff st[5];
  RbOff fin;
  int fc_w, fc_b, total;
};

inline int pad4(int x) { return (x + 3) & ~3; }

inline void make_rb(const GldmResNetCfg& c, int ch, int& o, RbOff& r) {
  r.mlp_w = o; o += pad4(2 * ch * c.emb_dim);
  r.mlp_b = o; o += pad4(2 * ch);
  r.p1_w = o; o += pad4(ch * ch * 3);
  r.p1_b = o; o += pad4(ch);
  r.n1_w = o; o += pad4(ch);
  r.n1_b = o; o += pad4(ch);
  r.p2_w = o; o += pad4(ch * ch * 3);
  r.p2_b = o; o += pad4(ch);
  r.n2_w = o; o += pad4(ch);
  r.n2_b = o; o += pad4(ch);
}

inline void make_layout(const GldmResNetCfg& c, ResNetLayout& l) {
  int o = 0;
  const int hd = c.heads * c.dim_head;
  l.init_w = o; o += pad4(c.ch[0] * 7);
  l.init_b = o; o += pad4(c.ch[0]);
  l.tm_freq = l.tm_w1 = l.tm_b1 = l.tm_w2 = l.tm_b2 = -1;
  if (c.time_cond) {
    l.tm_freq = o; o += pad4(c.fourier_half);
    l.tm_w1 = o; o += pad4(c.emb_dim * (2 * c.fourier_half + 1));
    l.tm_b1 = o; o += pad4(c.emb_dim);
    l.tm_w2 = o; o += pad4(c.emb_dim * c.emb_dim);
    l.tm_b2 = o; o += pad4(c.emb_dim);
  }
  l.in_w = o; o += pad4(c.emb_dim * c.cond_dim);
  l.in_b = o; o += pad4(c.emb_dim);
  for (int i = 0; i < c.n_stages; ++i) {
    const int ch = c.ch[i], cn = c.ch[i + 1];
    make_rb(c, ch, o, l.st[i].rb[0]);
    make_rb(c, ch, o, l.st[i].rb[1]);
    l.st[i].ln_g = o; o += pad4(ch);
    l.st[i].qkv_w = o; o += pad4(3 * hd * ch);
    l.st[i].out_w = o; o += pad4(ch * hd);
    l.st[i].out_b = o; o += pad4(ch);
    l.st[i].out_g = o; o += pad4(ch);
    l.st[i].down_w = o; o += pad4(cn * ch * 3);
    l.st[i].down_b = o; o += pad4(cn);
  }
  make_rb(c, c.ch[c.n_stages], o, l.fin);
  l.fc_w = o; o += pad4(c.ch[c.n_stages]);
  l.fc_b = o; o += pad4(1);
  l.total = o;
}

inline int check_cfg(const GldmResNetCfg* c) {
  GLDM_REQUIRE(c, "resnet: null cfg");
  GLDM_REQUIRE(c->L == 4 || c->L == 16, "resnet: sequence length L=%d not supported (4 or 16)", c->L);
  GLDM_REQUIRE(c->n_stages >= 1 && c->n_stages <= 5, "resnet: n_stages=%d", c->n_stages);
  GLDM_REQUIRE(c->heads * c->dim_head == 128 && c->dim_head == 32, "resnet: attention must be 4 heads x 32");
  GLDM_REQUIRE(c->emb_dim >= 4 && c->emb_dim <= 64, "resnet: emb_dim=%d (<=64)", c->emb_dim);
  GLDM_REQUIRE(c->cond_ch >= 1 && c->cond_ch <= 4, "resnet: cond_ch=%d (<=4)", c->cond_ch);
  GLDM_REQUIRE(c->cond_dim >= 1 && c->cond_dim <= 1024, "resnet: cond_dim=%d", c->cond_dim);
  GLDM_REQUIRE(!c->time_cond || (c->fourier_half >= 1 && c->fourier_half <= 16), "resnet: fourier_half");
  GLDM_REQUIRE(c->groups >= 1 && c->groups <= 8, "resnet: groups=%d (1..8)", c->groups);
  for (int i = 0; i <= c->n_stages; ++i) {
    const int ch = c->ch[i];
    GLDM_REQUIRE(ch == 4 || ch == 8 || ch == 16 || ch == 32 || ch == 64 || ch == 128 || ch == 256,
                 "resnet: channel width %d not supported (4,8,16,32,64,128,256)", ch);
    GLDM_REQUIRE(ch % c->groups == 0, "resnet: channels %d not divisible by groups %d", ch, c->groups);
  }
  return GLDM_OK;
}

// Philox4x32-10 (Salmon et al. 2011) + Box-Muller: standard normal keyed by (seed, sample, step, l)
__device__ __forceinline__ void philox_round(uint4& c, uint2& k) {
  const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
  const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
  c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
  k.x += 0x9E3779B9u;
  k.y += 0xBB67AE85u;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned sample, unsigned step, unsigned l) {
  uint4 c = make_uint4(sample, step, l >> 1, 0x5eedu);
  uint2 k = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(c, k);
  const float u1 = ((float)(c.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(c.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincosf(6.283185307179586f * u2, &sn, &cs);
  return (l & 1) ? rad * sn : rad * cs;
}

// ---------------------------------------------------------------------------------------------
// evaluation programs (GLDM_SCHED_EDM): the elucidated samplers as a list of network evaluations, one 16-float row each
//   [0] c_in  [1] c_skip  [2] c_out  [3] kind  [4..7] k0..k3  [8] noise slot (-1: none)  [9] x_all slot (-1: none)
// state per (sample, position): x (current), y, z.  D = c_skip * x_in + c_out * net(c_in * x_in)   (Karras et al. eq. 7,
// R/grasp_ldm/models/diffusion/elucidated_diffusion.py:126-150), clamped to [-1, 1] when the caller asks for it.
//   kind 0  stochastic Heun, first evaluation (:214-237):  x_in = x + k0 * (k3 * noise)   [k0 = sqrt(s_hat^2 - s^2), k3 = S_noise]
//           d = (x_in - D) / k1 [s_hat];  x <- x_in + k2 * d [k2 = s_next - s_hat];  y <- x_in;  z <- d
//   kind 1  second-order correction (:239-256):  x_in = x;  d' = (x_in - D) / k1 [s_next];  x <- y + k2 * (z + d') [k2 = (s_next - s_hat) / 2]
//   kind 2  DPM-Solver++(2M) (:282-313):  x_in = x;  dd = k0 * D + k1 * y [1 - gamma, gamma];  x <- k2 * x - k3 * dd;  y <- D
// ---------------------------------------------------------------------------------------------
constexpr int kEvalRow = 16;
__device__ __forceinline__ float eval_input(const float* cf, float x, float noise) {
  return ((int)cf[3] == 0) ? __fadd_rn(x, __fmul_rn(cf[4], __fmul_rn(cf[7], noise))) : x;
}
__device__ __forceinline__ void eval_update(const float* cf, float xin, float net, int clip, float& x, float& y, float& z) {
  float D = __fadd_rn(__fmul_rn(cf[1], xin), __fmul_rn(cf[2], net));
  if (clip) D = fminf(fmaxf(D, -1.0f), 1.0f);
  const int kind = (int)cf[3];
  if (kind == 0) {
    const float d = __fdiv_rn(__fsub_rn(xin, D), cf[5]);
    x = __fadd_rn(xin, __fmul_rn(cf[6], d));
    y = xin;
    z = d;
  } else if (kind == 1) {
    const float d2 = __fdiv_rn(__fsub_rn(xin, D), cf[5]);
    x = __fadd_rn(y, __fmul_rn(cf[6], __fadd_rn(z, d2)));
  } else {
    const float dd = __fadd_rn(__fmul_rn(cf[4], D), __fmul_rn(cf[5], y));
    x = __fsub_rn(__fmul_rn(cf[6], xin), __fmul_rn(cf[7], dd));
    y = D;
  }
}

}  // namespace gldm
