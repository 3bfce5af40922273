// PointNet++ set abstraction, fused: ball-query neighbours -> gather (coords - centre | features) -> SharedMLP(dim=2)
// (Conv2d k1 + BatchNorm(eval) + ReLU per layer) -> max over the neighbours.
//   PointNetSAModule.forward   R/grasp_ldm/models/modules/ext/pvcnn/modules/pointnet.py:100-111
//   BallQuery.forward          R/.../modules/ball_query.py:16-34
//   SharedMLP                  R/.../modules/shared_mlp.py:18-28
// The reference materialises the grouped tensor [B, C+3, M, U] (and every MLP activation of that shape) in HBM before
// the max; here a CTA owns one centre, the grouped rows and the activations of all layers live in shared memory and only
// [B, C_out, M] is written.  fp32 SIMT (the set-abstraction family is the PVCNN2 / PointNet++ harness side of the operator
// extension, SURVEY.md finding 1; its GEMMs are small: K <= 512).
#include "common.cuh"

namespace gldm {

constexpr int SA_TU = 32;          // neighbours per tile (the max over U is taken across tiles)
constexpr int SA_CMAX = 512;       // widest layer
constexpr int SA_THREADS = 256;
constexpr int SA_MAX_LAYERS = 4;

struct SaParams {
  const float* coords;     // [b][3][n]
  const float* centers;    // [b][3][m]
  const float* feats;      // [b][c][n] or NULL
  const int* idx;          // [b][m][u]
  int n, m, u, c, include_coords;
  int n_layers;
  int width[SA_MAX_LAYERS + 1];             // width[0] = input channels (c + 3), width[l + 1] = outputs of layer l
  const float* wt[SA_MAX_LAYERS];           // [ci][co] (transposed once on the host side: coalesced over co)
  const float* scale[SA_MAX_LAYERS];        // folded BatchNorm
  const float* shift[SA_MAX_LAYERS];
  float* out;              // [b][c_out][m]
};

// activations as [channel][SA_TU] rows: a thread owns output channels and keeps the 32 neighbours in registers;
// the input row of a channel is read as broadcast float4s
__global__ void __launch_bounds__(SA_THREADS) sa_mlp_max_kernel(const SaParams p) {
  extern __shared__ float smem[];
  float* bufA = smem;
  float* bufB = smem + SA_CMAX * SA_TU;
  const int mi = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int c_in = p.width[0];
  const int c_out = p.width[p.n_layers];
  const float* cb = p.coords + (size_t)b * 3 * p.n;
  const float* fb = p.feats ? p.feats + (size_t)b * p.c * p.n : nullptr;
  const int* ib = p.idx + ((size_t)b * p.m + mi) * p.u;
  float best[2];                                   // running max of this thread's output channels (tid, tid + 256)
  best[0] = best[1] = -INFINITY;
  float ctr[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) ctr[k] = __ldg(p.centers + ((size_t)b * 3 + k) * p.m + mi);
  for (int u0 = 0; u0 < p.u; u0 += SA_TU) {
    const int nu = min(SA_TU, p.u - u0);
    __syncthreads();
    // ---- gather the tile: rows = channels [coords - centre (3) | features (c)], columns = neighbours
    for (int e = tid; e < c_in * SA_TU; e += SA_THREADS) {
      const int ch = e / SA_TU, uu = e % SA_TU;
      float v = 0.f;
      if (uu < nu) {
        const int j = __ldg(ib + u0 + uu);
        if (p.include_coords && ch < 3) v = __ldg(cb + (size_t)ch * p.n + j) - ctr[ch];
        else v = __ldg(fb + (size_t)(ch - (p.include_coords ? 3 : 0)) * p.n + j);
      }
      bufA[e] = v;
    }
    __syncthreads();
    float* src = bufA;
    float* dst = bufB;
    for (int l = 0; l < p.n_layers; ++l) {
      const int ci = p.width[l], co = p.width[l + 1];
      const bool last = l + 1 == p.n_layers;
#pragma unroll
      for (int slot = 0; slot < 2; ++slot) {
        const int o = tid + slot * SA_THREADS;
        if (o >= co) continue;
        float acc[SA_TU];
#pragma unroll
        for (int uu = 0; uu < SA_TU; ++uu) acc[uu] = 0.f;
        const float* w = p.wt[l] + o;
        for (int k = 0; k < ci; ++k) {
          const float wk = __ldg(w + (size_t)k * co);
          const float4* x = reinterpret_cast<const float4*>(src + k * SA_TU);
#pragma unroll
          for (int q = 0; q < SA_TU / 4; ++q) {
            const float4 xv = x[q];
            acc[4 * q] = fmaf(wk, xv.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(wk, xv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(wk, xv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(wk, xv.w, acc[4 * q + 3]);
          }
        }
        const float sc = __ldg(p.scale[l] + o), sh = __ldg(p.shift[l] + o);
        if (last) {
          float mx = best[slot];
#pragma unroll
          for (int uu = 0; uu < SA_TU; ++uu)
            if (uu < nu) mx = fmaxf(mx, fmaxf(fmaf(acc[uu], sc, sh), 0.f));
          best[slot] = mx;
        } else {
#pragma unroll
          for (int uu = 0; uu < SA_TU; ++uu) dst[o * SA_TU + uu] = fmaxf(fmaf(acc[uu], sc, sh), 0.f);
        }
      }
      __syncthreads();
      float* t = src; src = dst; dst = t;
    }
  }
#pragma unroll
  for (int slot = 0; slot < 2; ++slot) {
    const int o = tid + slot * SA_THREADS;
    if (o < c_out) p.out[((size_t)b * c_out + o) * p.m + mi] = best[slot];
  }
}

// SE excite with ReLU: gate = sigmoid(W2 relu(W1 mean))   (R/.../modules/se.py:12-25 with use_relu=True, PVCNN2's PVConv)
__global__ void __launch_bounds__(128) se_gate_relu_kernel(const float* __restrict__ mean, const float* __restrict__ w1,
                                                           const float* __restrict__ w2, int c, int cr,
                                                           float* __restrict__ gate) {
  extern __shared__ float s_h[];
  const int b = blockIdx.x;
  const float* mb = mean + (size_t)b * c;
  for (int j = threadIdx.x; j < cr; j += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < c; ++k) a = fmaf(w1[j * c + k], mb[k], a);
    s_h[j] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < c; o += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < cr; ++k) a = fmaf(w2[o * cr + k], s_h[k], a);
    gate[(size_t)b * c + o] = 1.0f / (1.0f + expf(-a));
  }
}

}  // namespace gldm

using namespace gldm;

extern "C" int gldm_sa_mlp_max_f32(const float* coords, const float* centers, const float* feats, const int* idx, int b,
                                   int c, int n, int m, int u, int include_coords, int n_layers, const int* widths,
                                   const float* const* wt, const float* const* scale, const float* const* shift,
                                   float* out, void* stream) {
  GLDM_REQUIRE(b <= 0 || m <= 0 || (coords && centers && idx && out && widths && wt && scale && shift), "sa_mlp_max_f32: null pointer");
  GLDM_REQUIRE(b >= 0 && c >= 0 && n > 0 && m >= 0 && u > 0, "sa_mlp_max_f32: bad sizes");
  GLDM_REQUIRE(n_layers >= 1 && n_layers <= SA_MAX_LAYERS, "sa_mlp_max_f32: 1..%d layers", SA_MAX_LAYERS);
  GLDM_REQUIRE(include_coords || c > 0, "sa_mlp_max_f32: no features for grouping");
  GLDM_REQUIRE(c == 0 || feats, "sa_mlp_max_f32: null features");
  if (b == 0 || m == 0) return GLDM_OK;
  SaParams p = {};
  p.coords = coords; p.centers = centers; p.feats = feats; p.idx = idx;
  p.n = n; p.m = m; p.u = u; p.c = c; p.include_coords = include_coords; p.n_layers = n_layers; p.out = out;
  p.width[0] = c + (include_coords ? 3 : 0);
  for (int l = 0; l < n_layers; ++l) {
    p.width[l + 1] = widths[l];
    p.wt[l] = wt[l]; p.scale[l] = scale[l]; p.shift[l] = shift[l];
    GLDM_REQUIRE(widths[l] > 0 && widths[l] <= SA_CMAX && wt[l] && scale[l] && shift[l], "sa_mlp_max_f32: layer %d width %d (<= %d)", l, widths[l], SA_CMAX);
  }
  GLDM_REQUIRE(p.width[0] <= SA_CMAX, "sa_mlp_max_f32: %d input channels (<= %d)", p.width[0], SA_CMAX);
  const int smem = 2 * SA_CMAX * SA_TU * (int)sizeof(float);
  static SmemOptIn attr;
  if (int rc = opt_in_smem(attr, sa_mlp_max_kernel, smem, "sa_mlp_max_kernel")) return rc;
  sa_mlp_max_kernel<<<dim3(m, b), SA_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("sa_mlp_max_kernel");
}

extern "C" int gldm_se_gate_relu_f32(const float* mean, const float* w1, const float* w2, int b, int c, int cr, float* gate,
                                     void* stream) {
  GLDM_REQUIRE(b <= 0 || (mean && w1 && w2 && gate), "se_gate_relu_f32: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && cr > 0, "se_gate_relu_f32: bad sizes");
  if (b == 0) return GLDM_OK;
  se_gate_relu_kernel<<<b, 128, sizeof(float) * cr, (cudaStream_t)stream>>>(mean, w1, w2, c, cr, gate);
  return check_launch("se_gate_relu_kernel");
}
