// Strict-fp32 (SIMT FFMA) dense kernels of the PVCNN encoder: the parity-mode path.
// The tensor-core (tcgen05, bf16) path lives in encoder_tc.cu; both sit behind the same host module.
//
//   pointwise conv (+folded BN, ReLU, residual add)   R/../pvcnn/modules/shared_mlp.py:18-28,
//                                                     R/models/modules/pc_encoders.py:60-75
//   Conv3d k3 p1                                      R/../pvcnn/modules/pvconv.py:48-67
//   GroupNorm(8) + Swish (+ SE squeeze)               pvconv.py:56-58,66-68, R/../pvcnn/modules/se.py:22-25
//   SE excite                                         se.py:14-19
//   trilinear devoxelize x gate + point branch        pvconv.py:79-83, trilinear_devox.cu:21-105
//   Linear over the point axis                        pc_encoders.py:76-79
#include <cuda_bf16.h>

#include "common.cuh"

namespace gldm {

// ------------------------------------------------------------------------------------------------
// SGEMM: y[b, m, n] = epi( sum_k W[m,k] * x[b,k,n] ),  128x128x8 tiles, 8x8 register tile per thread
// ------------------------------------------------------------------------------------------------
constexpr int PW_BM = 128, PW_BN = 128, PW_BK = 8;

template <bool RELU>
__global__ void __launch_bounds__(256) pw_gemm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ scale,
                                                      const float* __restrict__ shift,
                                                      const float* __restrict__ add, int ci, int co, int n,
                                                      float* __restrict__ y) {
  __shared__ __align__(16) float As[2][PW_BK][PW_BM];
  __shared__ __align__(16) float Bs[2][PW_BK][PW_BN];
  const int b = blockIdx.z, m0 = blockIdx.y * PW_BM, n0 = blockIdx.x * PW_BN;
  const float* xb = x + (size_t)b * ci * n;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // global->smem assignment
  const int a_row = tid >> 1, a_k = (tid & 1) * 4;     // W[m0+a_row][k0+a_k .. +4)
  const int b_k = tid >> 5, b_col = (tid & 31) * 4;    // x[k0+b_k][n0+b_col .. +4)
  const bool a_vec = (ci % 4 == 0), b_vec = (n % 4 == 0);
  float4 ra, rb;

  auto load_tiles = [&](int k0) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int m = m0 + a_row;
    if (m < co) {
      if (a_vec && k0 + a_k + 3 < ci) {
        float4 t = __ldg(reinterpret_cast<const float4*>(w + (size_t)m * ci + k0 + a_k));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (k0 + a_k + q < ci) v[q] = __ldg(w + (size_t)m * ci + k0 + a_k + q);
      }
    }
    ra = make_float4(v[0], v[1], v[2], v[3]);
    float u[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = k0 + b_k;
    if (k < ci) {
      if (b_vec && n0 + b_col + 3 < n) {
        float4 t = __ldg(reinterpret_cast<const float4*>(xb + (size_t)k * n + n0 + b_col));
        u[0] = t.x; u[1] = t.y; u[2] = t.z; u[3] = t.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (n0 + b_col + q < n) u[q] = __ldg(xb + (size_t)k * n + n0 + b_col + q);
      }
    }
    rb = make_float4(u[0], u[1], u[2], u[3]);
  };
  auto store_tiles = [&](int buf) {
    As[buf][a_k + 0][a_row] = ra.x;
    As[buf][a_k + 1][a_row] = ra.y;
    As[buf][a_k + 2][a_row] = ra.z;
    As[buf][a_k + 3][a_row] = ra.w;
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_col]) = rb;
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (ci + PW_BK - 1) / PW_BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * PW_BK);
#pragma unroll
    for (int kk = 0; kk < PW_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue
  float* yb = y + (size_t)b * co * n;
  const float* ab = add ? add + (size_t)b * co * n : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= co) continue;
    const float sc = scale ? __ldg(scale + m) : 1.f, sh = shift ? __ldg(shift + m) : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = n0 + h * 64 + tx * 4;
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = fmaf(acc[i][h * 4 + q], sc, sh);
        if (RELU) v = fmaxf(v, 0.f);
        o[q] = v;
      }
      if (b_vec && col + 3 < n) {
        if (ab) {
          float4 t = __ldg(reinterpret_cast<const float4*>(ab + (size_t)m * n + col));
          o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
        }
        *reinterpret_cast<float4*>(yb + (size_t)m * n + col) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (col + q < n) yb[(size_t)m * n + col + q] = o[q] + (ab ? ab[(size_t)m * n + col + q] : 0.f);
      }
    }
  }
}

// few output channels (co <= 8): one thread per point, streaming over ci (HBM/L2-bound)
template <int CO>
__global__ void __launch_bounds__(256) pw_small_co_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ w,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int ci, int n, int act,
                                                          float* __restrict__ y) {
  extern __shared__ float s_w[];   // [CO][ci]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < CO * ci; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* xb = x + (size_t)b * ci * n + j;
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = 0.f;
  for (int k = 0; k < ci; ++k) {
    const float v = __ldg(xb + (size_t)k * n);
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = fmaf(s_w[o * ci + k], v, acc[o]);
  }
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    float v = fmaf(acc[o], scale ? scale[o] : 1.f, shift ? shift[o] : 0.f);
    if (act == 1) v = fmaxf(v, 0.f);
    y[((size_t)b * CO + o) * n + j] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Conv3d k3 p1: implicit GEMM, tile = 256 consecutive voxels x 48 output channels, one input channel
// (27 taps) staged per iteration.  w is pre-permuted to [ci][27][co].
// ------------------------------------------------------------------------------------------------
// CL = true (fused voxel branch, co == 48, GroupNorm(8)): the result goes to the zero-padded channels-last bf16 grid
// [b * (r+2)^3][y_stride] the tensor-core Conv3d reads, and the GroupNorm statistics are accumulated on the way (a warp
// owns exactly the 6 channels of one group): stats f64[b][8][2] += (sum, sum of squares).
constexpr int C3_VT = 256, C3_CT = 48;
template <bool CL>
__global__ void __launch_bounds__(256) conv3d_k3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, int ci, int co, int r,
                                                        float* __restrict__ y, __nv_bfloat16* __restrict__ y_cl,
                                                        int y_stride, double* __restrict__ stats) {
  __shared__ __align__(16) float As[27][C3_VT];
  __shared__ __align__(16) float Ws[27][C3_CT];
  const int b = blockIdx.z, co0 = blockIdx.y * C3_CT, v0 = blockIdx.x * C3_VT;
  const int r2 = r * r, r3 = r2 * r;
  const int tid = threadIdx.x;
  const int vg = tid & 31, cg = tid >> 5;         // 32 voxel groups x 8 voxels, 8 channel groups x 6
  // each thread gathers the taps of one voxel (tid) for the staged input channel
  const int gv = v0 + tid;
  const bool gvalid = gv < r3;
  const int gx = gv / r2, gy = (gv / r) % r, gz = gv % r;
  const float* xb = x + (size_t)b * ci * r3;
  float acc[8][6];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;

  for (int c = 0; c < ci; ++c) {
    __syncthreads();
    const float* xc = xb + (size_t)c * r3;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int dx = t / 9 - 1, dy = (t / 3) % 3 - 1, dz = t % 3 - 1;
      const int xx = gx + dx, yy = gy + dy, zz = gz + dz;
      float v = 0.f;
      if (gvalid && (unsigned)xx < (unsigned)r && (unsigned)yy < (unsigned)r && (unsigned)zz < (unsigned)r)
        v = __ldg(xc + xx * r2 + yy * r + zz);
      As[t][tid] = v;
    }
    for (int i = tid; i < 27 * C3_CT; i += 256) {
      const int t = i / C3_CT, o = i - t * C3_CT;
      Ws[t][o] = (co0 + o < co) ? __ldg(w + ((size_t)c * 27 + t) * co + co0 + o) : 0.f;
    }
    __syncthreads();
#pragma unroll 3
    for (int t = 0; t < 27; ++t) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[t][vg * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[t][128 + vg * 4]);
      const float2 w0 = *reinterpret_cast<const float2*>(&Ws[t][cg * 6]);
      const float2 w1 = *reinterpret_cast<const float2*>(&Ws[t][cg * 6 + 2]);
      const float2 w2 = *reinterpret_cast<const float2*>(&Ws[t][cg * 6 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
  }
  if (CL) {
    const int rp = r + 2;
    float gs = 0.f, gq = 0.f;
    float bz[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) bz[j] = bias ? __ldg(bias + cg * 6 + j) : 0.f;
    // stage the [256 voxels][48 channels] bf16 tile in shared memory (the tap buffer is free now) so that the grid rows
    // are written as whole 16-byte pieces instead of 4-byte scatters
    __syncthreads();
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(&As[0][0]);      // [256][48]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int vl = (i >> 2) * 128 + vg * 4 + (i & 3);
      const bool ok = v0 + vl < r3;
      float o[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        o[j] = acc[i][j] + bz[j];
        if (ok) { gs += o[j]; gq = fmaf(o[j], o[j], gq); }
      }
      __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(tile + vl * 48 + cg * 6);
      dst[0] = __floats2bfloat162_rn(o[0], o[1]);
      dst[1] = __floats2bfloat162_rn(o[2], o[3]);
      dst[2] = __floats2bfloat162_rn(o[4], o[5]);
    }
    __syncthreads();
    for (int idx = tid; idx < 256 * 6; idx += 256) {
      const int vl = idx / 6, piece = idx - vl * 6, v = v0 + vl;
      if (v >= r3) continue;
      const int vx = v / r2, vy = (v / r) % r, vz = v % r;
      const uint4 val = *reinterpret_cast<const uint4*>(tile + vl * 48 + piece * 8);
      *reinterpret_cast<uint4*>(y_cl + ((size_t)b * rp * rp * rp + ((size_t)(vx + 1) * rp + (vy + 1)) * rp + (vz + 1)) * y_stride +
                                piece * 8) = val;
    }
    gs = warp_sum(gs);
    gq = warp_sum(gq);
    if (vg == 0) {      // per-block partials [b][block][16] (0..7 sums, 8..15 sums of squares): no atomics, summed in order later
      double* dst = stats + ((size_t)b * gridDim.x + blockIdx.x) * 16;
      dst[cg] = (double)gs;
      dst[8 + cg] = (double)gq;
    }
    return;
  }
  float* yb = y + (size_t)b * co * r3;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int o = co0 + cg * 6 + j;
    if (o >= co) continue;
    const float bz = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int v = v0 + h * 128 + vg * 4;
      if (v + 3 < r3 && (r3 % 4 == 0)) {
        *reinterpret_cast<float4*>(yb + (size_t)o * r3 + v) =
            make_float4(acc[h * 4 + 0][j] + bz, acc[h * 4 + 1][j] + bz, acc[h * 4 + 2][j] + bz,
                        acc[h * 4 + 3][j] + bz);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (v + q < r3) yb[(size_t)o * r3 + v + q] = acc[h * 4 + q][j] + bz;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm + Swish, one block per (cloud, group); two-pass statistics; optional SE squeeze
// ------------------------------------------------------------------------------------------------
__device__ float block_sum_1024(float v, float* s_buf) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) s_buf[wid] = v;
  __syncthreads();
  float t = lane < nw ? s_buf[lane] : 0.f;
  return warp_sum(t);
}

__global__ void __launch_bounds__(512) groupnorm_swish_kernel(float* __restrict__ x,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int c, int s,
                                                              int groups, float eps, float* __restrict__ se_mean) {
  __shared__ float s_buf[32];
  const int b = blockIdx.y, g = blockIdx.x;
  const int cpg = c / groups;
  float* xg = x + ((size_t)b * c + (size_t)g * cpg) * s;
  const size_t tot = (size_t)cpg * s;
  float sum = 0.f;
  for (size_t i = threadIdx.x; i < tot; i += blockDim.x) sum += xg[i];
  const float mean = block_sum_1024(sum, s_buf) / (float)tot;
  float sq = 0.f;
  for (size_t i = threadIdx.x; i < tot; i += blockDim.x) {
    const float d = xg[i] - mean;
    sq = fmaf(d, d, sq);
  }
  const float rstd = rsqrtf(block_sum_1024(sq, s_buf) / (float)tot + eps);
  for (int ch = 0; ch < cpg; ++ch) {
    const float ga = gamma[g * cpg + ch] * rstd, be = beta[g * cpg + ch] - mean * ga;
    float* xc = xg + (size_t)ch * s;
    float part = 0.f;
    for (int i = threadIdx.x; i < s; i += blockDim.x) {
      const float v = fmaf(xc[i], ga, be);
      const float o = v * (1.0f / (1.0f + expf(-v)));    // Swish, R/models/modules/modules.py:5-7
      xc[i] = o;
      part += o;
    }
    if (se_mean) {
      const float tsum = block_sum_1024(part, s_buf);
      if (threadIdx.x == 0) se_mean[(size_t)b * c + g * cpg + ch] = tsum / (float)s;
    }
  }
}

// SE excite: gate = sigmoid(W2 swish(W1 mean)); one block per cloud
__global__ void __launch_bounds__(128) se_gate_kernel(const float* __restrict__ mean, const float* __restrict__ w1,
                                                      const float* __restrict__ w2, int c, int cr,
                                                      float* __restrict__ gate) {
  extern __shared__ float s_h[];   // [cr]
  const int b = blockIdx.x;
  const float* mb = mean + (size_t)b * c;
  for (int j = threadIdx.x; j < cr; j += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < c; ++k) a = fmaf(w1[j * c + k], mb[k], a);
    s_h[j] = a * (1.0f / (1.0f + expf(-a)));
  }
  __syncthreads();
  for (int o = threadIdx.x; o < c; o += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < cr; ++k) a = fmaf(w2[o * cr + k], s_h[k], a);
    gate[(size_t)b * c + o] = 1.0f / (1.0f + expf(-a));
  }
}

// trilinear devoxelize of (grid * gate) + point branch
__global__ void __launch_bounds__(128) devox_gate_add_kernel(const float* __restrict__ coords,
                                                             const float* __restrict__ grid,
                                                             const float* __restrict__ gate,
                                                             const float* __restrict__ point, int c, int n, int r,
                                                             float* __restrict__ out) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r2 = r * r, r3 = r2 * r;
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
  const int xh = xd1 > 0 ? r2 : 0, yh = yd1 > 0 ? r : 0, zh = zd1 > 0 ? 1 : 0;
  int id[8];
  id[0] = (int)xl * r2 + (int)yl * r + (int)zl;
  id[1] = id[0] + zh; id[2] = id[0] + yh; id[3] = id[2] + zh;
  id[4] = id[0] + xh; id[5] = id[4] + zh; id[6] = id[4] + yh; id[7] = id[6] + zh;
  const int c0 = blockIdx.y * 8, c1 = min(c, c0 + 8);
  for (int ch = c0; ch < c1; ++ch) {
    const float* f = grid + ((size_t)b * c + ch) * r3;
    const float g = gate ? __ldg(gate + (size_t)b * c + ch) : 1.f;
    float acc = __fmul_rn(wgt[1], __fmul_rn(__ldg(f + id[1]), g));
    acc = __fmaf_rn(wgt[0], __fmul_rn(__ldg(f + id[0]), g), acc);
#pragma unroll
    for (int k = 2; k < 8; ++k) acc = __fmaf_rn(wgt[k], __fmul_rn(__ldg(f + id[k]), g), acc);
    const size_t o = ((size_t)b * c + ch) * n + i;
    out[o] = acc + (point ? point[o] : 0.f);
  }
}

// y[row, f] = sum_n x[row, n] * W[f, n] + bias[f]; one block per row, one warp per output feature
__global__ void __launch_bounds__(256) linear_lastdim_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias, int n, int f,
                                                             float* __restrict__ y) {
  extern __shared__ float s_x[];
  const int row = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_x[i] = x[(size_t)row * n + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = wid; o < f; o += nw) {
    const float* wr = w + (size_t)o * n;
    float a = 0.f;
    for (int i = lane; i < n; i += 32) a = fmaf(__ldg(wr + i), s_x[i], a);
    a = warp_sum(a);
    if (lane == 0) y[(size_t)row * f + o] = a + (bias ? bias[o] : 0.f);
  }
}

}  // namespace gldm

using namespace gldm;

extern "C" int gldm_pointwise_conv_f32(const float* x, const float* w, const float* scale, const float* shift,
                                       const float* add, int b, int ci, int co, int n, int act, float* y,
                                       void* stream) {
  GLDM_REQUIRE(b <= 0 || (x && w && y), "pointwise_conv_f32: null pointer");
  GLDM_REQUIRE(b >= 0 && ci > 0 && co > 0 && n > 0, "pointwise_conv_f32: bad sizes");
  GLDM_REQUIRE(act == 0 || act == 1, "pointwise_conv_f32: act must be 0 or 1");
  if (b == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (co <= 8 && !add && ci <= 4096) {
    dim3 grid(ceil_div(n, 256), b);
    size_t smem = sizeof(float) * co * ci;
    switch (co) {
#define SMALL(C) case C: pw_small_co_kernel<C><<<grid, 256, smem, s>>>(x, w, scale, shift, ci, n, act, y); break;
      SMALL(1) SMALL(2) SMALL(3) SMALL(4) SMALL(5) SMALL(6) SMALL(7) SMALL(8)
#undef SMALL
    }
    return check_launch("pw_small_co_kernel");
  }
  dim3 grid(ceil_div(n, PW_BN), ceil_div(co, PW_BM), b);
  if (act == 1) pw_gemm_kernel<true><<<grid, 256, 0, s>>>(x, w, scale, shift, add, ci, co, n, y);
  else pw_gemm_kernel<false><<<grid, 256, 0, s>>>(x, w, scale, shift, add, ci, co, n, y);
  return check_launch("pw_gemm_kernel");
}

extern "C" int gldm_conv3d_k3_f32(const float* x, const float* w, const float* bias, int b, int ci, int co, int r,
                                  float* y, void* stream) {
  GLDM_REQUIRE(b <= 0 || (x && w && y), "conv3d_k3_f32: null pointer");
  GLDM_REQUIRE(b >= 0 && ci > 0 && co > 0 && r > 0, "conv3d_k3_f32: bad sizes");
  if (b == 0) return GLDM_OK;
  dim3 grid(ceil_div(r * r * r, C3_VT), ceil_div(co, C3_CT), b);
  conv3d_k3_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, ci, co, r, y, nullptr, 0, nullptr);
  return check_launch("conv3d_k3_kernel");
}

extern "C" int gldm_block_partials_to_stats(const double* part, int b, int nblk, double* stats, void* stream);

extern "C" int gldm_conv3d_k3_f32_cl(const float* x, const float* w, const float* bias, int b, int ci, int r, void* y_cl,
                                     int y_stride, double* stats, void* ws, void* stream) {
  GLDM_REQUIRE(b <= 0 || (x && w && y_cl && stats && ws), "conv3d_k3_f32_cl: null pointer");
  GLDM_REQUIRE(b >= 0 && ci > 0 && r > 0, "conv3d_k3_f32_cl: bad sizes");
  GLDM_REQUIRE(y_stride >= C3_CT && y_stride % 8 == 0, "conv3d_k3_f32_cl: bad row stride");
  if (b == 0) return GLDM_OK;
  dim3 grid(ceil_div(r * r * r, C3_VT), 1, b);
  conv3d_k3_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, ci, C3_CT, r, nullptr,
                                                                 reinterpret_cast<__nv_bfloat16*>(y_cl), y_stride,
                                                                 reinterpret_cast<double*>(ws));
  int rc = check_launch("conv3d_k3_kernel");
  if (rc) return rc;
  return gldm_block_partials_to_stats(reinterpret_cast<const double*>(ws), b, (int)grid.x, stats, stream);
}

extern "C" int gldm_groupnorm_swish_f32(float* x, const float* gamma, const float* beta, int b, int c, int s,
                                        int groups, float eps, float* se_mean, void* stream) {
  GLDM_REQUIRE(b <= 0 || (x && gamma && beta), "groupnorm_swish_f32: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && s > 0 && groups > 0 && c % groups == 0, "groupnorm_swish_f32: bad sizes");
  if (b == 0) return GLDM_OK;
  groupnorm_swish_kernel<<<dim3(groups, b), 512, 0, (cudaStream_t)stream>>>(x, gamma, beta, c, s, groups, eps,
                                                                            se_mean);
  return check_launch("groupnorm_swish_kernel");
}

extern "C" int gldm_se_gate_f32(const float* mean, const float* w1, const float* w2, int b, int c, int cr,
                                float* gate, void* stream) {
  GLDM_REQUIRE(b <= 0 || (mean && w1 && w2 && gate), "se_gate_f32: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && cr > 0, "se_gate_f32: bad sizes");
  if (b == 0) return GLDM_OK;
  se_gate_kernel<<<b, 128, sizeof(float) * cr, (cudaStream_t)stream>>>(mean, w1, w2, c, cr, gate);
  return check_launch("se_gate_kernel");
}

extern "C" int gldm_devox_gate_add_f32(const float* coords, const float* grid, const float* gate,
                                       const float* point, int b, int c, int n, int r, float* out, void* stream) {
  GLDM_REQUIRE(b <= 0 || (coords && grid && out), "devox_gate_add_f32: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && n > 0 && r > 0, "devox_gate_add_f32: bad sizes");
  if (b == 0) return GLDM_OK;
  dim3 g(ceil_div(n, 128), ceil_div(c, 8), b);
  devox_gate_add_kernel<<<g, 128, 0, (cudaStream_t)stream>>>(coords, grid, gate, point, c, n, r, out);
  return check_launch("devox_gate_add_kernel");
}

extern "C" int gldm_linear_lastdim_f32(const float* x, const float* w, const float* bias, int rows, int n, int f,
                                       float* y, void* stream) {
  GLDM_REQUIRE(x && w && y, "linear_lastdim_f32: null pointer");
  GLDM_REQUIRE(rows >= 0 && n > 0 && f > 0 && n <= 12000, "linear_lastdim_f32: bad sizes");
  if (rows == 0) return GLDM_OK;
  linear_lastdim_kernel<<<rows, 256, sizeof(float) * n, (cudaStream_t)stream>>>(x, w, bias, n, f, y);
  return check_launch("linear_lastdim_kernel");
}
