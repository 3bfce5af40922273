// ResNet1D family in strict fp32 (SIMT FFMA): the whole network - and for the sampler the whole
// T-step reverse-diffusion loop - runs inside ONE persistent kernel launch.
//
//   TimeConditionedResNet1D.forward   R/models/modules/resnets.py:558-616   (denoiser, eps prediction)
//   ResNet1D.forward                  resnets.py:373-424                    (decoder trunk)
//   GaussianDiffusion1D.sample        R/models/diffusion/gaussian_diffusion.py:232-277
//   DDPM/DDIM scheduler step          diffusers (restated in oracle/schedulers.py)
//   ConditionalGraspPoseDecoder       R/models/grasp_vae.py:401-436
//
// One CTA owns a tile of S = 32/L samples for the entire run (L = sequence length: 4 for the fpc
// denoiser, 16 for the decoder / ppc denoiser).  Activations never leave shared memory: three buffers
// X (block input / residual), H (intermediate), T (GEMM output, 384 rows for qkv) hold [channel][row]
// with row m = l*S + s and a zero halo of PAD floats on both sides, so a k=3 convolution is three
// accumulating GEMMs whose A operand is the same buffer shifted by -S / 0 / +S floats (no im2col).
// Weights are streamed L2 -> shared memory with cp.async in double-buffered [K-chunk][N] panels; they were
// weight-standardised and transposed to [K][N] once by gldm_resnet_prepare.  GroupNorm, FiLM, SiLU,
// LayerNorm, the linear attention, the scheduler update and the noise draw are fused in between.
#include <math.h>

#include "common.cuh"
#include "resnet_layout.cuh"

namespace gldm {

// ---------------------------------------------------------------------------------------------
// weight preparation kernels
// ---------------------------------------------------------------------------------------------
// dst[k][n] = src[n][k]   (src [N][K] row-major)
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int K) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  int n = i / K, k = i - n * K;
  dst[(size_t)k * N + n] = src[i];
}
// weight standardisation per output channel (resnets.py:85-91, eps 1e-5), written transposed [K][N]
__global__ void ws_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* w = src + (size_t)n * K;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += w[k];
  const float mean = warp_sum(s) / (float)K;
  float q = 0.f;
  for (int k = lane; k < K; k += 32) { float d = w[k] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) / (float)K + 1e-5f);
  for (int k = lane; k < K; k += 32) dst[(size_t)k * N + n] = (w[k] - mean) * rstd;
}
__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

constexpr int kWStage = 4096;   // floats per weight-panel stage (16 KB), two stages

struct ResNetParams {
  GldmResNetCfg cfg;
  ResNetLayout lay;
  const float* W;          // prepared blob
  int mode;                // 0 sampler, 1 single denoiser evaluation, 2 decoder
  int n;                   // samples
  int gpo;                 // grasps per object: conditioning row of sample i is i / gpo
  const float* x_in;       // [n][L] (x_T / x) or z_h [n][D] for the decoder
  const float* z_cond;     // [n_obj][cond_ch][cond_dim]
  const int* t_sample;     // mode 1: per-sample timestep
  const float* tf_sample;  // mode 1, continuous time (elucidated sampler: c_noise(sigma)); used when non-NULL
  int n_steps;
  const int* timesteps;    // device [n_steps]
  const float* times_f;    // device [n_steps] continuous time (evaluation programs); used when non-NULL
  const float* cls_emb;    // [n_obj][emb] added to the time embedding (class-conditioned denoiser) or NULL
  const float* coef;       // device [n_steps][8] (DDPM / DDIM) or [n_steps][16] (GLDM_SCHED_EDM)
  int sched_kind, clip;
  const float* noise;      // [n_steps][n][L] or null
  unsigned long long seed;
  float* x_out;            // [n][L]
  float* x_all;            // [n_steps+1][n][L] or null
  const float* head;       // decoder head weights
  int D;
  float* tmrp;
  float* logit;
};

// sum over the 32 lanes of v[lane] for each index: lane i returns sum_lanes v[i] (31 shuffles)
__device__ __forceinline__ float reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float send = (lane & 16) ? v[i] : v[i + 16];
    const float keep = (lane & 16) ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = (lane & 8) ? v[i] : v[i + 8];
    const float keep = (lane & 8) ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = (lane & 4) ? v[i] : v[i + 4];
    const float keep = (lane & 4) ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (lane & 2) ? v[i] : v[i + 2];
    const float keep = (lane & 2) ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const float send = (lane & 1) ? v[0] : v[1];
    const float keep = (lane & 1) ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

template <int L>
struct Tile {
  static constexpr int S = 32 / L;                 // samples per CTA
  static constexpr int PAD = (S < 4 ? 4 : S);      // zero halo (floats) on both sides of a channel row
  static constexpr int CS = 32 + 2 * PAD;          // channel stride
};

// ---------------------------------------------------------------------------------------------
// out[n][PAD+m] = bias[n] + sum_{ci,dk} W[(ci*TAPS+dk)][n] * in[ci][PAD + m + (dk-TAPS/2)*S]
// 256 threads: rg = tid&7 owns rows 4rg..4rg+3, cg = tid>>3 owns columns cg*TN .. +TN
// ---------------------------------------------------------------------------------------------
template <int L, int TN, int TAPS>
__device__ __noinline__ void conv_gemm_tn(const float* __restrict__ Wg, const float* __restrict__ bias, int Cin, int N,
                             const float* xin, float* xout, float* wst) {
  using T = Tile<L>;
  const int tid = threadIdx.x, rg = tid & 7, cg = tid >> 3;
  const int n0 = cg * TN;
  const bool active = n0 < N;
  const int K = Cin * TAPS;
  int kc = (kWStage / N);
  kc -= kc % TAPS;
  if (kc > K) kc = K;
  const int nchunks = (K + kc - 1) / kc;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto issue = [&](int chunk) {
    const int k0 = chunk * kc;
    const int rows = min(kc, K - k0);
    const int nvec = (rows * N) >> 2;
    const float* src = Wg + (size_t)k0 * N;
    float* dst = wst + (chunk & 1) * kWStage;
    for (int i = tid; i < nvec; i += 256) cp_async16(dst + i * 4, src + i * 4);
    cp_async_commit();
  };
  __syncthreads();   // producers of xin are done; previous users of wst are done
  issue(0);
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    cp_async_wait_all();
    __syncthreads();
    if (chunk + 1 < nchunks) issue(chunk + 1);
    if (active) {
      const float* ws = wst + (chunk & 1) * kWStage + n0;
      const int k0 = chunk * kc;
      const int rows = min(kc, K - k0);
      const int ci0 = k0 / TAPS;
      const float* abase = xin + (size_t)ci0 * T::CS + T::PAD + 4 * rg;
      for (int cc = 0; cc < rows / TAPS; ++cc) {
#pragma unroll
        for (int dk = 0; dk < TAPS; ++dk) {
          const float* ap = abase + cc * T::CS + (dk - TAPS / 2) * T::S;
          float a[4];
          if (T::S % 4 == 0 || TAPS == 1) {
            const float4 t = *reinterpret_cast<const float4*>(ap);
            a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
          } else {
            const float2 t0 = *reinterpret_cast<const float2*>(ap);
            const float2 t1 = *reinterpret_cast<const float2*>(ap + 2);
            a[0] = t0.x; a[1] = t0.y; a[2] = t1.x; a[3] = t1.y;
          }
          const float* wp = ws + (cc * TAPS + dk) * N;
          float w[TN];
          if (TN % 4 == 0) {
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
              const float4 t = *reinterpret_cast<const float4*>(wp + j);
              w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
            }
          } else if (TN % 2 == 0) {
#pragma unroll
            for (int j = 0; j < TN; j += 2) {
              const float2 t = *reinterpret_cast<const float2*>(wp + j);
              w[j] = t.x; w[j + 1] = t.y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < TN; ++j) w[j] = wp[j];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const float bz = bias ? __ldg(bias + n0 + j) : 0.f;
      *reinterpret_cast<float4*>(xout + (size_t)(n0 + j) * T::CS + T::PAD + 4 * rg) =
          make_float4(acc[0][j] + bz, acc[1][j] + bz, acc[2][j] + bz, acc[3][j] + bz);
    }
  }
  __syncthreads();
}

template <int L, int TAPS>
__device__ void conv_gemm(const float* Wg, const float* bias, int Cin, int N, const float* xin, float* xout,
                          float* wst) {
  if (N <= 32) conv_gemm_tn<L, 1, TAPS>(Wg, bias, Cin, N, xin, xout, wst);
  else if (N == 64) conv_gemm_tn<L, 2, TAPS>(Wg, bias, Cin, N, xin, xout, wst);
  else if (N == 128) conv_gemm_tn<L, 4, TAPS>(Wg, bias, Cin, N, xin, xout, wst);
  else if (N == 256) conv_gemm_tn<L, 8, TAPS>(Wg, bias, Cin, N, xin, xout, wst);
  else conv_gemm_tn<L, 12, TAPS>(Wg, bias, Cin, N, xin, xout, wst);   // 384 (qkv)
}

// GroupNorm over (channels of the group x L positions) per sample, then optional FiLM, SiLU, residual.
// One warp per group, lane = row m; statistics are reduced over the lanes that share the sample.
template <int L>
__device__ __noinline__ void groupnorm_film_silu(const float* src, float* dst, int c, int groups, const float* gamma,
                                    const float* beta, const float* film /* [S][2c] or null */,
                                    const float* resid) {
  using T = Tile<L>;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int s = lane % T::S;
  const int cpg = c / groups;
  for (int g = wid; g < groups; g += 8) {
    const float* p = src + (size_t)g * cpg * T::CS + T::PAD + lane;
    float sum = 0.f;
    for (int ch = 0; ch < cpg; ++ch) sum += p[ch * T::CS];
#pragma unroll
    for (int o = T::S; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)(cpg * L);
    float sq = 0.f;
    for (int ch = 0; ch < cpg; ++ch) { const float d = p[ch * T::CS] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
    for (int o = T::S; o < 32; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)(cpg * L) + 1e-5f);
    for (int ch = 0; ch < cpg; ++ch) {
      const int cc = g * cpg + ch;
      float v = (p[ch * T::CS] - mean) * rstd * __ldg(gamma + cc) + __ldg(beta + cc);
      if (film) v = fmaf(v, film[s * 2 * c + cc], film[s * 2 * c + c + cc]);
      v = silu_acc(v);
      const size_t o = (size_t)cc * T::CS + T::PAD + lane;
      if (resid) v += resid[o];
      dst[o] = v;
    }
  }
}

// LayerNorm over channels for every row (resnets.py:104-113): dst = (x-mean)*rstd*g (+ resid)
template <int L>
__device__ __noinline__ void chan_layernorm(const float* src, float* dst, int c, const float* g, const float* resid) {
  using T = Tile<L>;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* p = src + T::PAD + lane;
  float sum = 0.f;
  for (int ch = 0; ch < c; ++ch) sum += p[ch * T::CS];
  const float mean = sum / (float)c;
  float sq = 0.f;
  for (int ch = 0; ch < c; ++ch) { const float d = p[ch * T::CS] - mean; sq = fmaf(d, d, sq); }
  const float rstd = rsqrtf(sq / (float)c + 1e-5f);
  for (int ch = wid; ch < c; ch += 8) {
    const size_t o = (size_t)ch * T::CS + T::PAD + lane;
    float v = (src[o] - mean) * rstd * __ldg(g + ch);
    if (resid) v += resid[o];
    dst[o] = v;
  }
}

// linear attention core (resnets.py:221-235) for one tile: qkv rows in T (q:0..127, k:128..255, v:256..383),
// result -> H rows 0..127.  One warp per (sample, head), lane = d / e.
template <int L>
__device__ __noinline__ void linear_attention_core(const float* bufT, float* bufH, float* scratch) {
  using T = Tile<L>;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* sA = scratch + wid * (L * L);
  constexpr float scale = 0.17677669529663687f;   // 32^-0.5
  for (int unit = wid; unit < T::S * 4; unit += 8) {
    const int s = unit % T::S, h = unit / T::S;
    const float* qp = bufT + (size_t)(h * 32 + lane) * T::CS + T::PAD + s;
    const float* kp = qp + 128 * T::CS;
    const float* vp = qp + 256 * T::CS;
    float q[L], k[L], v[L];
    float kmax = -INFINITY;
#pragma unroll
    for (int n = 0; n < L; ++n) {
      q[n] = qp[n * T::S];
      k[n] = kp[n * T::S];
      v[n] = vp[n * T::S];
      kmax = fmaxf(kmax, k[n]);
    }
    float ksum = 0.f;
#pragma unroll
    for (int n = 0; n < L; ++n) { k[n] = expf(k[n] - kmax); ksum += k[n]; }
#pragma unroll
    for (int n = 0; n < L; ++n) {
      k[n] = k[n] / ksum;                                   // softmax over positions (dim=-1)
      const float mx = warp_max(q[n]);
      const float e = expf(q[n] - mx);
      q[n] = (e / warp_sum(e)) * scale;                     // softmax over d (dim=-2), then * scale
    }
    // A[n'][n] = sum_d k[d][n'] q[d][n]  (lanes = d), reduce-scattered 32 pairs at a time
    __syncwarp();
#pragma unroll
    for (int blk = 0; blk < (L * L + 31) / 32; ++blk) {
      float pr[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int pidx = blk * 32 + i;
        pr[i] = (pidx < L * L) ? k[pidx / L] * q[pidx % L] : 0.f;
      }
      const float tot = reduce_scatter32(pr, lane);
      if (blk * 32 + lane < L * L) sA[blk * 32 + lane] = tot;
    }
    __syncwarp();
    // out[e][n] = sum_n' v[e][n'] A[n'][n]   (lane = e)
    float* op = bufH + (size_t)(h * 32 + lane) * T::CS + T::PAD + s;
#pragma unroll
    for (int n = 0; n < L; ++n) {
      float o = 0.f;
#pragma unroll
      for (int m = 0; m < L; ++m) o = fmaf(v[m], sA[m * L + n], o);
      op[n * T::S] = o;
    }
    __syncwarp();
  }
}

template <int L>
__device__ __noinline__ void resnet_block(const ResNetParams& p, const RbOff& o, int c, float* X, float* H, float* Tb,
                             float* film, float* wst, const float* s_u) {
  using T = Tile<L>;
  const float* W = p.W;
  const int emb = p.cfg.emb_dim, R = p.cfg.cond_ch;
  // FiLM vectors: mult = W_s u + R b_s + R ; add = W_h u + R b_h   (Block.forward :172-175 summed over r)
  for (int idx = threadIdx.x; idx < T::S * 2 * c; idx += 256) {
    const int s = idx / (2 * c), j = idx - s * 2 * c;
    float a = (float)R * __ldg(W + o.mlp_b + j) + (j < c ? (float)R : 0.f);
    const float* wt = W + o.mlp_w + j;            // [emb][2c]
    for (int e = 0; e < emb; ++e) a = fmaf(__ldg(wt + (size_t)e * 2 * c), s_u[s * 64 + e], a);
    film[idx] = a;
  }
  conv_gemm<L, 3>(W + o.p1_w, W + o.p1_b, c, c, X, Tb, wst);
  groupnorm_film_silu<L>(Tb, H, c, p.cfg.groups, W + o.n1_w, W + o.n1_b, film, nullptr);
  conv_gemm<L, 3>(W + o.p2_w, W + o.p2_b, c, c, H, Tb, wst);
  groupnorm_film_silu<L>(Tb, X, c, p.cfg.groups, W + o.n2_w, W + o.n2_b, nullptr, X);
}

template <int L>
__global__ void __launch_bounds__(256, 1) resnet_kernel(const ResNetParams p) {
  using T = Tile<L>;
  constexpr int S = T::S, PAD = T::PAD, CS = T::CS;
  extern __shared__ __align__(16) float smem[];
  float* X = smem;
  float* H = X + 256 * CS;
  float* Tb = H + 256 * CS;
  float* film = Tb + 256 * CS;          // rows 256..383 of T double as FiLM storage
  float* wst = Tb + 384 * CS;
  float* s_u = wst + 2 * kWStage;       // [S][64]
  float* s_in = s_u + S * 64;           // [S][4][64]
  float* s_te = s_in + S * 4 * 64;      // [S][64]
  float* s_h1 = s_te + S * 64;          // [S][64]
  float* s_four = s_h1 + S * 64;        // [S][36]
  float* s_x = s_four + S * 36;         // [32]   state, row m = l*S + s
  float* s_eps = s_x + 32;              // [32]
  float* s_y = s_eps + 32;              // [32]   evaluation programs: second / third state vector, network input
  float* s_z = s_y + 32;
  float* s_xin = s_z + 32;

  const GldmResNetCfg& cfg = p.cfg;
  const ResNetLayout& lay = p.lay;
  const float* W = p.W;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int s0 = blockIdx.x * S;
  const int emb = cfg.emb_dim, R = cfg.cond_ch;

  for (int i = tid; i < (256 + 256 + 384) * CS; i += 256) smem[i] = 0.f;
  // ---- input state
  if (tid < 32) {
    const int l = tid / S, s = tid % S;
    float v = 0.f;
    if (s0 + s < p.n) {
      if (p.mode == 2) {   // decoder in_layer: Linear(D -> L)   (grasp_vae.py:419)
        v = __ldg(p.head + L * p.D + l);
        for (int d = 0; d < p.D; ++d) v = fmaf(__ldg(p.head + l * p.D + d), __ldg(p.x_in + (size_t)(s0 + s) * p.D + d), v);
      } else {
        v = __ldg(p.x_in + (size_t)(s0 + s) * L + l);
      }
    }
    s_x[tid] = v;
    s_y[tid] = 0.f;
    s_z[tid] = 0.f;
    if (p.mode == 0 && p.x_all && s0 + s < p.n) p.x_all[(size_t)(s0 + s) * L + l] = v;
  }
  // ---- input (conditioning) embedding: SiLU(Linear(z_cond))  (resnets.py:531-533,596) - once per launch
  for (int idx = tid; idx < S * R * emb; idx += 256) {
    const int e = idx % emb, r = (idx / emb) % R, s = idx / (emb * R);
    float a = 0.f;
    if (s0 + s < p.n) {
      const int obj = (s0 + s) / p.gpo;
      const float* z = p.z_cond + ((size_t)obj * R + r) * cfg.cond_dim;
      const float* w = W + lay.in_w + (size_t)e * cfg.cond_dim;
      a = __ldg(W + lay.in_b + e);
      for (int j = 0; j < cfg.cond_dim; ++j) a = fmaf(__ldg(w + j), __ldg(z + j), a);
      a = silu_acc(a);
    }
    s_in[(s * 4 + r) * 64 + e] = a;
  }
  __syncthreads();

  const int n_steps = (p.mode == 0) ? p.n_steps : 1;
  for (int step = 0; step < n_steps; ++step) {
    // ---- time embedding (resnets.py:44-56, 517-522) per sample
    if (cfg.time_cond) {
      const int fh = cfg.fourier_half, fd = 2 * fh + 1;
      for (int idx = tid; idx < S * fd; idx += 256) {
        const int s = idx / fd, j = idx - s * fd;
        float tf = 0.f;
        if (p.mode == 0) tf = p.times_f ? p.times_f[step] : (float)p.timesteps[step];
        else if (s0 + s < p.n) tf = p.tf_sample ? p.tf_sample[s0 + s] : (float)p.t_sample[s0 + s];
        float v = tf;
        if (j > 0) {
          const int i = (j - 1) % fh;
          const float f = __fmul_rn(__fmul_rn(__fmul_rn(tf, __ldg(W + lay.tm_freq + i)), 2.0f), 3.14159274101257324f);
          v = (j - 1 < fh) ? sinf(f) : cosf(f);
        }
        s_four[s * 36 + j] = v;
      }
      __syncthreads();
      for (int idx = tid; idx < S * emb; idx += 256) {
        const int s = idx / emb, e = idx - s * emb;
        float a = __ldg(W + lay.tm_b1 + e);
        const float* w = W + lay.tm_w1 + e * fd;
        for (int j = 0; j < fd; ++j) a = fmaf(__ldg(w + j), s_four[s * 36 + j], a);
        s_h1[s * 64 + e] = gelu_erf(a);
      }
      __syncthreads();
      for (int idx = tid; idx < S * emb; idx += 256) {
        const int s = idx / emb, e = idx - s * emb;
        float a = __ldg(W + lay.tm_b2 + e);
        const float* w = W + lay.tm_w2 + e * emb;
        for (int j = 0; j < emb; ++j) a = fmaf(__ldg(w + j), s_h1[s * 64 + j], a);
        if (p.cls_emb) a += __ldg(p.cls_emb + (size_t)(min(s0 + s, p.n - 1) / p.gpo) * emb + e);    // class_conditioned_resnet.py:96-98
        s_te[s * 64 + e] = a;
      }
      __syncthreads();
    }
    // ---- network input: the state itself, or (evaluation programs) c_in * x_in with the stochastic churn added
    const bool edm = p.mode == 0 && p.sched_kind == GLDM_SCHED_EDM;
    if (tid < 32) {
      float v = s_x[tid];
      if (edm) {
        const float* cf = p.coef + (size_t)step * kEvalRow;
        const int l = tid / S, s = tid % S, slot = (int)cf[8];
        float z = 0.f;
        if (slot >= 0 && s0 + s < p.n)
          z = p.noise ? __ldg(p.noise + ((size_t)slot * p.n + s0 + s) * L + l)
                      : philox_normal(p.seed, (unsigned)(s0 + s), (unsigned)slot, (unsigned)l);
        v = eval_input(cf, v, z);
        s_xin[tid] = v;
        v = __fmul_rn(cf[0], v);
      }
      s_eps[tid] = v;                       // init_conv reads the (scaled) input from here
    }
    __syncthreads();
    // ---- u[s][e] = sum_r silu(latent_emb[s][r][e])   (ResnetBlock.mlp's SiLU, hoisted; :183-197)
    for (int idx = tid; idx < S * emb; idx += 256) {
      const int s = idx / emb, e = idx - s * emb;
      const float te = cfg.time_cond ? s_te[s * 64 + e] : 0.f;
      float a = 0.f;
      for (int r = 0; r < R; ++r) a += silu_acc(te + s_in[(s * 4 + r) * 64 + e]);
      s_u[s * 64 + e] = a;
    }
    // ---- init_conv: Conv1d(1 -> ch0, k7, p3)
    {
      const int c0 = cfg.ch[0];
      for (int idx = tid; idx < c0 * 32; idx += 256) {
        const int co = idx >> 5, m = idx & 31, l = m / S, s = m % S;
        float a = __ldg(W + lay.init_b + co);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const int ll = l + j - 3;
          if (ll >= 0 && ll < L) a = fmaf(__ldg(W + lay.init_w + co * 7 + j), s_eps[ll * S + s], a);
        }
        X[(size_t)co * CS + PAD + m] = a;
      }
    }
    __syncthreads();
    float* cx = X;
    float* chh = H;
    for (int st = 0; st < cfg.n_stages; ++st) {
      const int c = cfg.ch[st], cn = cfg.ch[st + 1];
      const StageOff& so = lay.st[st];
      resnet_block<L>(p, so.rb[0], c, cx, chh, Tb, film, wst, s_u);
      __syncthreads();
      resnet_block<L>(p, so.rb[1], c, cx, chh, Tb, film, wst, s_u);
      __syncthreads();
      // Residual(PreNorm(LinearAttention))
      chan_layernorm<L>(cx, chh, c, W + so.ln_g, nullptr);
      conv_gemm<L, 1>(W + so.qkv_w, nullptr, c, 384, chh, Tb, wst);
      linear_attention_core<L>(Tb, chh, wst);
      conv_gemm<L, 1>(W + so.out_w, W + so.out_b, 128, c, chh, Tb, wst);
      chan_layernorm<L>(Tb, cx, c, W + so.out_g, cx);
      // Conv1d(c -> cn, k3, p1); rows [c, 256) of the destination are stale but never read
      conv_gemm<L, 3>(W + so.down_w, W + so.down_b, c, cn, cx, chh, wst);
      float* t = cx; cx = chh; chh = t;
    }
    const int cl = cfg.ch[cfg.n_stages];
    resnet_block<L>(p, lay.fin, cl, cx, chh, Tb, film, wst, s_u);
    __syncthreads();
    // ---- final_conv (1x1 -> 1 channel) and the step epilogue
    if (wid == 0) {
      float a = __ldg(W + lay.fc_b);
      const float* xp = cx + PAD + lane;
      for (int ch = 0; ch < cl; ++ch) a = fmaf(__ldg(W + lay.fc_w + ch), xp[(size_t)ch * CS], a);
      const int l = lane / S, s = lane % S;
      const bool valid = s0 + s < p.n;
      if (edm) {
        const float* cf = p.coef + (size_t)step * kEvalRow;
        float x = s_x[lane], y = s_y[lane], z = s_z[lane];
        eval_update(cf, s_xin[lane], a, p.clip, x, y, z);
        s_x[lane] = x; s_y[lane] = y; s_z[lane] = z;
        const int slot = (int)cf[9];
        if (p.x_all && valid && slot >= 0) p.x_all[((size_t)slot * p.n + s0 + s) * L + l] = x;
      } else if (p.mode == 0) {
        const float* cf = p.coef + (size_t)step * 8;
        const float x = s_x[lane];
        float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(cf[0], a)), cf[1]);
        if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
        float prev;
        if (p.sched_kind == GLDM_SCHED_DDPM) {
          prev = __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], x));
          if (cf[4] > 0.f && valid) {
            const float z = p.noise ? __ldg(p.noise + ((size_t)step * p.n + s0 + s) * L + l)
                                    : philox_normal(p.seed, (unsigned)(s0 + s), (unsigned)step, (unsigned)l);
            prev = __fadd_rn(prev, __fmul_rn(cf[4], z));
          }
        } else {
          prev = __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], a));
        }
        s_x[lane] = prev;
        if (p.x_all && valid) p.x_all[((size_t)(step + 1) * p.n + s0 + s) * L + l] = prev;
      } else {
        s_eps[lane] = a;
      }
    }
    __syncthreads();
    // re-zero the stale rows that a later, narrower GEMM would otherwise expose as conv input halos:
    // halos are never written, so nothing to do here.
  }
  // ---- outputs
  if (p.mode == 0 || p.mode == 1) {
    if (tid < 32) {
      const int l = tid / S, s = tid % S;
      if (s0 + s < p.n) p.x_out[(size_t)(s0 + s) * L + l] = (p.mode == 0) ? s_x[tid] : s_eps[tid];
    }
  } else {
    // decoder heads: tmrp = Linear(L -> 6), class_logits = Linear(L -> 1)   (grasp_vae.py:428-430)
    const float* hw = p.head + L * p.D + L;   // tmrp_w [6][L], tmrp_b [6], cls_w [L], cls_b [1]
    for (int idx = tid; idx < S * 7; idx += 256) {
      const int s = idx / 7, j = idx - s * 7;
      if (s0 + s >= p.n) continue;
      float a;
      if (j < 6) {
        a = __ldg(hw + 6 * L + j);
        for (int l = 0; l < L; ++l) a = fmaf(__ldg(hw + j * L + l), s_eps[l * S + s], a);
        p.tmrp[(size_t)(s0 + s) * 6 + j] = a;
      } else {
        a = __ldg(hw + 6 * L + 6 + L);
        for (int l = 0; l < L; ++l) a = fmaf(__ldg(hw + 6 * L + 6 + l), s_eps[l * S + s], a);
        p.logit[s0 + s] = a;
      }
    }
  }
}

template <int L>
static size_t resnet_smem_bytes() {
  using T = Tile<L>;
  return sizeof(float) * ((256 + 256 + 384) * T::CS + 2 * kWStage + T::S * (64 + 256 + 64 + 64 + 36) + 160);
}

static int launch_resnet(const ResNetParams& p, cudaStream_t s) {
  if (p.n == 0) return GLDM_OK;
  if (p.cfg.L == 4) {
    static SmemOptIn attr;
    if (int rc = opt_in_smem(attr, resnet_kernel<4>, (int)resnet_smem_bytes<4>(), "resnet_kernel<4>")) return rc;
    resnet_kernel<4><<<ceil_div(p.n, Tile<4>::S), 256, resnet_smem_bytes<4>(), s>>>(p);
  } else {
    static SmemOptIn attr;
    if (int rc = opt_in_smem(attr, resnet_kernel<16>, (int)resnet_smem_bytes<16>(), "resnet_kernel<16>")) return rc;
    resnet_kernel<16><<<ceil_div(p.n, Tile<16>::S), 256, resnet_smem_bytes<16>(), s>>>(p);
  }
  return check_launch("resnet_kernel");
}

}  // namespace gldm

using namespace gldm;

extern "C" long long gldm_resnet_raw_floats(const GldmResNetCfg* cfg) {
  if (check_cfg(cfg) != GLDM_OK) return -1;
  ResNetLayout l;
  make_layout(*cfg, l);
  return l.total;
}
extern "C" long long gldm_resnet_prepared_floats(const GldmResNetCfg* cfg) { return gldm_resnet_raw_floats(cfg); }

extern "C" int gldm_resnet_prepare(const GldmResNetCfg* cfg, const float* raw, float* prep, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  GLDM_REQUIRE(raw && prep, "resnet_prepare: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  ResNetLayout l;
  make_layout(*cfg, l);
  int launches = 0;
  // everything starts as a straight copy; matrices the kernels want as [K][N] are then rewritten
  copy_kernel<<<ceil_div(l.total, 256), 256, 0, s>>>(raw, prep, l.total);
  ++launches;
  auto tr = [&](int off, int N, int K) {
    transpose_kernel<<<ceil_div(N * K, 256), 256, 0, s>>>(raw + off, prep + off, N, K);
    ++launches;
  };
  auto ws = [&](int off, int N, int K) {
    ws_transpose_kernel<<<ceil_div(N, 8), 256, 0, s>>>(raw + off, prep + off, N, K);
    ++launches;
  };
  auto rb = [&](const RbOff& o, int ch) {
    tr(o.mlp_w, 2 * ch, cfg->emb_dim);
    ws(o.p1_w, ch, ch * 3);
    ws(o.p2_w, ch, ch * 3);
  };
  const int hd = cfg->heads * cfg->dim_head;
  for (int i = 0; i < cfg->n_stages; ++i) {
    const int ch = cfg->ch[i], cn = cfg->ch[i + 1];
    rb(l.st[i].rb[0], ch);
    rb(l.st[i].rb[1], ch);
    tr(l.st[i].qkv_w, 3 * hd, ch);
    tr(l.st[i].out_w, ch, hd);
    tr(l.st[i].down_w, cn, ch * 3);
  }
  rb(l.fin, cfg->ch[cfg->n_stages]);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("resnet_prepare: %s", cudaGetErrorString(e));
    return GLDM_ECUDA;
  }
  count_launch(launches);
  return GLDM_OK;
}

static int fill_common(ResNetParams& p, const GldmResNetCfg* cfg, const float* prepared) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  GLDM_REQUIRE(prepared, "resnet: null prepared weights");
  p.cfg = *cfg;
  make_layout(*cfg, p.lay);
  p.W = prepared;
  return GLDM_OK;
}

extern "C" int gldm_sampler_run_f32(const GldmResNetCfg* cfg, const float* prepared, const float* x_T,
                                    const float* z_obj, int n, int grasps_per_obj, int n_steps,
                                    const int* timesteps_host, const float* coef_host, int sched_kind,
                                    int clip_sample, const float* noise, unsigned long long seed, float* x_out,
                                    float* x_all, void* stream) {
  ResNetParams p = {};
  int rc = fill_common(p, cfg, prepared);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->time_cond, "sampler_run: the denoiser must be time conditioned");
  GLDM_REQUIRE(x_T && z_obj && x_out && timesteps_host && coef_host, "sampler_run: null pointer");
  GLDM_REQUIRE(n >= 0 && grasps_per_obj > 0 && n_steps > 0, "sampler_run: bad sizes");
  GLDM_REQUIRE(sched_kind == GLDM_SCHED_DDPM || sched_kind == GLDM_SCHED_DDIM, "sampler_run: bad scheduler");
  cudaStream_t s = (cudaStream_t)stream;
  // timestep / coefficient tables travel with the launch (stream-ordered scratch)
  void* scratch = nullptr;
  const size_t tb = sizeof(int) * n_steps, cb = sizeof(float) * 8 * (size_t)n_steps;
  if (cudaMallocAsync(&scratch, tb + cb + 64, s) != cudaSuccess) {
    set_error("sampler_run: cudaMallocAsync failed");
    return GLDM_ECUDA;
  }
  float* d_coef = reinterpret_cast<float*>(scratch);
  int* d_ts = reinterpret_cast<int*>(reinterpret_cast<char*>(scratch) + cb);
  cudaMemcpyAsync(d_coef, coef_host, cb, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_ts, timesteps_host, tb, cudaMemcpyHostToDevice, s);
  p.mode = 0; p.n = n; p.gpo = grasps_per_obj; p.x_in = x_T; p.z_cond = z_obj;
  p.n_steps = n_steps; p.timesteps = d_ts; p.coef = d_coef; p.sched_kind = sched_kind; p.clip = clip_sample;
  p.noise = noise; p.seed = seed; p.x_out = x_out; p.x_all = x_all;
  rc = launch_resnet(p, s);
  cudaFreeAsync(scratch, s);
  return rc;
}

extern "C" int gldm_sampler_run_ex_f32(const GldmResNetCfg* cfg, const float* prepared, const GldmSamplerArgs* a, void* stream) {
  ResNetParams p = {};
  int rc = fill_common(p, cfg, prepared);
  if (rc) return rc;
  GLDM_REQUIRE(a, "sampler_run_ex: null arguments");
  GLDM_REQUIRE(cfg->time_cond, "sampler_run_ex: the denoiser must be time conditioned");
  GLDM_REQUIRE(a->n >= 0 && a->grasps_per_obj > 0 && a->n_steps > 0, "sampler_run_ex: bad sizes");
  if (a->n == 0) return GLDM_OK;
  GLDM_REQUIRE(a->x_init && a->z_obj && a->x_out && a->coef, "sampler_run_ex: null pointer");
  const bool edm = a->sched_kind == GLDM_SCHED_EDM;
  GLDM_REQUIRE(a->sched_kind == GLDM_SCHED_DDPM || a->sched_kind == GLDM_SCHED_DDIM || edm, "sampler_run_ex: bad scheduler");
  GLDM_REQUIRE(edm ? a->times != nullptr : (a->timesteps != nullptr || a->times != nullptr),
               "sampler_run_ex: the fp32 kernel needs the (device) time of every evaluation");
  p.mode = 0; p.n = a->n; p.gpo = a->grasps_per_obj; p.x_in = a->x_init; p.z_cond = a->z_obj;
  p.n_steps = a->n_steps; p.timesteps = a->timesteps; p.times_f = a->times; p.coef = a->coef; p.sched_kind = a->sched_kind;
  p.clip = a->clip_sample; p.noise = a->noise; p.seed = a->seed; p.cls_emb = a->cls_emb; p.x_out = a->x_out; p.x_all = a->x_all;
  return launch_resnet(p, (cudaStream_t)stream);
}

extern "C" int gldm_denoiser_forward_ex_f32(const GldmResNetCfg* cfg, const float* prepared, const float* x, const int* t,
                                            const float* tf, const float* z_cond, const float* cls_emb, int n, float* eps,
                                            void* stream) {
  ResNetParams p = {};
  int rc = fill_common(p, cfg, prepared);
  if (rc) return rc;
  GLDM_REQUIRE(n >= 0, "denoiser_forward_ex: bad n");
  if (n == 0) return GLDM_OK;
  GLDM_REQUIRE(x && z_cond && eps && (t || tf || !cfg->time_cond), "denoiser_forward_ex: null pointer");
  p.mode = 1; p.n = n; p.gpo = 1; p.x_in = x; p.z_cond = z_cond; p.t_sample = t; p.tf_sample = tf; p.cls_emb = cls_emb;
  p.x_out = eps;
  return launch_resnet(p, (cudaStream_t)stream);
}

// cls_embed of the class-conditioned denoiser: SiLU(Linear(1 -> emb))   R/grasp_ldm/models/modules/class_conditioned_resnet.py:43-46
__global__ void class_embed_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ cls,
                                   int n, int emb, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * emb) return;
  const int e = i % emb;
  const float a = fmaf(__ldg(w + e), __ldg(cls + i / emb), __ldg(b + e));
  out[i] = silu_acc(a);
}

extern "C" int gldm_class_embed(const float* w, const float* b, const float* cls, int n, int emb, float* out, void* stream) {
  GLDM_REQUIRE(n <= 0 || (w && b && cls && out), "class_embed: null pointer");
  GLDM_REQUIRE(n >= 0 && emb > 0, "class_embed: bad sizes");
  if (n == 0) return GLDM_OK;
  class_embed_kernel<<<ceil_div(n * emb, 256), 256, 0, (cudaStream_t)stream>>>(w, b, cls, n, emb, out);
  return check_launch("class_embed_kernel");
}

extern "C" int gldm_denoiser_forward_f32(const GldmResNetCfg* cfg, const float* prepared, const float* x,
                                         const int* t, const float* z_cond, int n, float* eps, void* stream) {
  ResNetParams p = {};
  int rc = fill_common(p, cfg, prepared);
  if (rc) return rc;
  GLDM_REQUIRE(x && z_cond && eps && (t || !cfg->time_cond), "denoiser_forward: null pointer");
  GLDM_REQUIRE(n >= 0, "denoiser_forward: bad n");
  p.mode = 1; p.n = n; p.gpo = 1; p.x_in = x; p.z_cond = z_cond; p.t_sample = t; p.x_out = eps;
  return launch_resnet(p, (cudaStream_t)stream);
}

extern "C" int gldm_denoiser_forward_f32_ftime(const GldmResNetCfg* cfg, const float* prepared, const float* x,
                                               const float* t, const float* z_cond, int n, float* eps, void* stream) {
  ResNetParams p = {};
  int rc = fill_common(p, cfg, prepared);
  if (rc) return rc;
  GLDM_REQUIRE(cfg->time_cond, "denoiser_forward_ftime: needs a time-conditioned configuration");
  GLDM_REQUIRE(n <= 0 || (x && z_cond && eps && t), "denoiser_forward_ftime: null pointer");
  GLDM_REQUIRE(n >= 0, "denoiser_forward_ftime: bad n");
  if (n == 0) return GLDM_OK;
  p.mode = 1; p.n = n; p.gpo = 1; p.x_in = x; p.z_cond = z_cond; p.tf_sample = t; p.x_out = eps;
  return launch_resnet(p, (cudaStream_t)stream);
}

extern "C" int gldm_decoder_forward_f32(const GldmResNetCfg* cfg, const float* prepared, const float* head,
                                        int D, const float* z_h, const float* z_obj, int n, int grasps_per_obj,
                                        float* tmrp, float* logit, void* stream) {
  ResNetParams p = {};
  int rc = fill_common(p, cfg, prepared);
  if (rc) return rc;
  GLDM_REQUIRE(!cfg->time_cond, "decoder_forward: the decoder trunk is not time conditioned");
  GLDM_REQUIRE(head && z_h && z_obj && tmrp && logit, "decoder_forward: null pointer");
  GLDM_REQUIRE(n >= 0 && grasps_per_obj > 0 && D > 0 && D <= 64, "decoder_forward: bad sizes");
  p.mode = 2; p.n = n; p.gpo = grasps_per_obj; p.x_in = z_h; p.z_cond = z_obj; p.head = head; p.D = D;
  p.tmrp = tmrp; p.logit = logit;
  return launch_resnet(p, (cudaStream_t)stream);
}
