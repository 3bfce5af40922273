// Point-cloud operators of the GraspLDM encoder backbone, written for sm_100a.
//
// C-ABI replacements for the reference's `_pvcnn_backend` extension
// (R = /root/reference/grasp_ldm/models/modules/ext/pvcnn/modules/functional/src).
// The reference launches one 512-thread block per cloud for every op, uses global float atomics
// and ~10 block barriers per FPS round; here
//   * FPS keeps every point and its running min-distance in registers, picks the round winner with
//     two REDUX (warp max of the distance bits, warp min of the tie-break key) per level and one
//     block barrier per round;
//   * ball query is a warp-per-centre ballot/popc compaction with early exit;
//   * grouping / gather are 128-bit vectorised, write-coalesced gathers;
//   * voxelisation ranks the points of a voxel in index order (match.any + ordered warp rounds) and
//     accumulates in that order, so results are deterministic.
#include "common.cuh"

namespace gldm {

// ------------------------------------------------------------------------------------------------
// furthest point sampling                        R/sampling/sampling.cu:86-167
// ------------------------------------------------------------------------------------------------
// Tie-break contract of the reference (512-thread strided scan + pairwise tree that keeps the lower
// slot unless strictly greater): winner = max min-distance; ties -> smallest (k mod 512), then
// smallest k.  Encoded as a 32-bit key whose minimum wins.
__device__ __forceinline__ unsigned fps_key(int k) { return ((unsigned)(k & 511) << 22) | (unsigned)k; }

template <int PPT>
__global__ void __launch_bounds__(1024) fps_kernel(const float* __restrict__ coords, int n, int m,
                                                   int* __restrict__ indices) {
  if (m <= 0) return;
  const int b = blockIdx.x;
  const float* cx = coords + (size_t)b * 3 * n;
  const float* cy = cx + n;
  const float* cz = cy + n;
  int* out = indices + (size_t)b * m;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  __shared__ unsigned s_bits[2][32];
  __shared__ unsigned s_key[2][32];

  float px[PPT], py[PPT], pz[PPT], dist[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    int k = tid + j * blockDim.x;
    bool v = k < n;
    px[j] = v ? cx[k] : 0.f;
    py[j] = v ? cy[k] : 0.f;
    pz[j] = v ? cz[k] : 0.f;
    dist[j] = 1e38f;   // sampling.cpp:53-54
  }
  int old = 0;
  if (tid == 0) out[0] = 0;
  for (int r = 1; r < m; ++r) {
    const float x1 = __ldg(cx + old), y1 = __ldg(cy + old), z1 = __ldg(cz + old);
    unsigned bits = 0u, key = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      int k = tid + j * blockDim.x;
      if (k < n) {
        float d = sqdist_ref(px[j] - x1, py[j] - y1, pz[j] - z1);
        float d2 = fminf(d, dist[j]);
        dist[j] = d2;
        unsigned db = __float_as_uint(d2);        // d2 >= +0: bit order == value order
        if (key == 0xffffffffu || db > bits) {    // first strictly greater within the thread
          bits = db;
          key = fps_key(k);
        }
      }
    }
    // warp level: max distance, then min key among the maxima
    unsigned wb = __reduce_max_sync(0xffffffffu, bits);
    unsigned wk = __reduce_min_sync(0xffffffffu, bits == wb ? key : 0xffffffffu);
    const int buf = r & 1;
    if (lane == 0) {
      s_bits[buf][wid] = wb;
      s_key[buf][wid] = wk;
    }
    __syncthreads();
    unsigned b2 = lane < nwarps ? s_bits[buf][lane] : 0u;
    unsigned k2 = lane < nwarps ? s_key[buf][lane] : 0xffffffffu;
    unsigned gb = __reduce_max_sync(0xffffffffu, b2);
    unsigned gk = __reduce_min_sync(0xffffffffu, (b2 == gb) ? k2 : 0xffffffffu);
    old = (gk == 0xffffffffu) ? 0 : (int)(gk & 0x3fffffu);
    if (tid == 0) out[r] = old;
  }
}

// ------------------------------------------------------------------------------------------------
// ball query                                     R/ball_query/ball_query.cu:19-50
// ------------------------------------------------------------------------------------------------
constexpr int kBqCentersPerBlock = 32;
__global__ void __launch_bounds__(256) ball_query_kernel(const float* __restrict__ centers,
                                                         const float* __restrict__ points, int n, int m,
                                                         float r2, int u, int* __restrict__ out) {
  const int b = blockIdx.y;
  const float* px = points + (size_t)b * 3 * n;
  const float* py = px + n;
  const float* pz = py + n;
  const float* cc = centers + (size_t)b * 3 * m;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int j = blockIdx.x * kBqCentersPerBlock + wid; j < min(m, (blockIdx.x + 1) * kBqCentersPerBlock);
       j += nw) {
    const float cx = __ldg(cc + j), cy = __ldg(cc + m + j), cz = __ldg(cc + 2 * m + j);
    int* o = out + ((size_t)b * m + j) * u;
    int cnt = 0, first = 0;
    for (int base = 0; base < n && cnt < u; base += 32) {
      int k = base + lane;
      bool hit = false;
      if (k < n) {
        float d2 = sqdist_ref(cx - __ldg(px + k), cy - __ldg(py + k), cz - __ldg(pz + k));
        hit = d2 < r2;
      }
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask) {
        if (cnt == 0) first = base + __ffs(mask) - 1;
        int pos = cnt + __popc(mask & lt);
        if (hit && pos < u) o[pos] = k;
        cnt += __popc(mask);
      }
    }
    // the first hit pre-fills every slot; no hit leaves zeros (ball_query.cu:39-44)
    cnt = min(cnt, u);
    for (int v = cnt + lane; v < u; v += 32) o[v] = first;
  }
}

// ------------------------------------------------------------------------------------------------
// grouping / gather                              R/grouping/grouping.cu:18-36, R/sampling/sampling.cu:17-39
// out[b,c,q] = feat[b,c,idx[b,q]],  q in [0, mu)
// ------------------------------------------------------------------------------------------------
constexpr int kGroupChanTile = 8;
template <bool VEC4>
__global__ void __launch_bounds__(256) group_gather_kernel(const float* __restrict__ feat,
                                                           const int* __restrict__ idx, int c, int n, int mu,
                                                           float* __restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kGroupChanTile;
  const int c1 = min(c, c0 + kGroupChanTile);
  const int* ib = idx + (size_t)b * mu;
  const float* fb = feat + (size_t)b * c * n;
  float* ob = out + (size_t)b * c * mu;
  if (VEC4) {
    int q = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (q >= mu) return;
    int4 i4 = *reinterpret_cast<const int4*>(ib + q);
    for (int ch = c0; ch < c1; ++ch) {
      const float* f = fb + (size_t)ch * n;
      float4 v = make_float4(__ldg(f + i4.x), __ldg(f + i4.y), __ldg(f + i4.z), __ldg(f + i4.w));
      __stcs(reinterpret_cast<float4*>(ob + (size_t)ch * mu + q), v);
    }
  } else {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= mu) return;
    int i = ib[q];
    for (int ch = c0; ch < c1; ++ch) ob[(size_t)ch * mu + q] = __ldg(fb + (size_t)ch * n + i);
  }
}

// grad_x[b,c,idx[b,q]] += grad_y[b,c,q]          (grouping.cu:51-72, sampling.cu:54-73)
__global__ void __launch_bounds__(256) group_scatter_add_kernel(const float* __restrict__ gy,
                                                                const int* __restrict__ idx, int c, int n,
                                                                int mu, float* __restrict__ gx) {
  const int b = blockIdx.z;
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= mu) return;
  int i = idx[(size_t)b * mu + q];
  for (int ch = blockIdx.y * kGroupChanTile; ch < min(c, (int)(blockIdx.y + 1) * kGroupChanTile); ++ch)
    atomicAdd(gx + ((size_t)b * c + ch) * n + i, gy[((size_t)b * c + ch) * mu + q]);
}

// ------------------------------------------------------------------------------------------------
// three nearest neighbours + interpolation       R/interpolate/neighbor_interpolate.cu:20-116
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) three_nn_kernel(const float* __restrict__ points,
                                                       const float* __restrict__ centers,
                                                       const float* __restrict__ feats, int c, int m, int n,
                                                       float* __restrict__ out, int* __restrict__ idx,
                                                       float* __restrict__ wts) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* pp = points + (size_t)b * 3 * n;
  const float* cc = centers + (size_t)b * 3 * m;
  const float ux = pp[j], uy = pp[j + n], uz = pp[j + 2 * n];
  double best0 = 1e40, best1 = 1e40, best2 = 1e40;
  int i0 = 0, i1 = 0, i2 = 0;
  for (int k = 0; k < m; ++k) {
    float d = sqdist_ref(ux - __ldg(cc + k), uy - __ldg(cc + m + k), uz - __ldg(cc + 2 * m + k));
    if (d < best2) {
      best2 = d; i2 = k;
      if (d < best1) {
        best2 = best1; i2 = i1; best1 = d; i1 = k;
        if (d < best0) { best1 = best0; i1 = i0; best0 = d; i0 = k; }
      }
    }
  }
  best0 = fmax(fmin((double)1e10f, best0), (double)1e-10f);
  best1 = fmax(fmin((double)1e10f, best1), (double)1e-10f);
  best2 = fmax(fmin((double)1e10f, best2), (double)1e-10f);
  float d0d1 = (float)(best0 * best1), d0d2 = (float)(best0 * best2), d1d2 = (float)(best1 * best2);
  float inv = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(d0d1, d0d2), d1d2));
  float w0 = __fmul_rn(d1d2, inv), w1 = __fmul_rn(d0d2, inv), w2 = __fmul_rn(d0d1, inv);
  float* wb = wts + (size_t)b * 3 * n;
  int* ib = idx + (size_t)b * 3 * n;
  wb[j] = w0; wb[j + n] = w1; wb[j + 2 * n] = w2;
  ib[j] = i0; ib[j + n] = i1; ib[j + 2 * n] = i2;
  const float* fb = feats + (size_t)b * c * m;
  float* ob = out + (size_t)b * c * n;
  for (int ch = 0; ch < c; ++ch) {
    const float* f = fb + (size_t)ch * m;
    // reference SASS order: FMUL(2nd term), FFMA(1st), FFMA(3rd)
    float acc = __fmul_rn(__ldg(f + i1), w1);
    acc = __fmaf_rn(__ldg(f + i0), w0, acc);
    acc = __fmaf_rn(__ldg(f + i2), w2, acc);
    ob[(size_t)ch * n + j] = acc;
  }
}

__global__ void __launch_bounds__(256) three_nn_grad_kernel(const float* __restrict__ gy,
                                                            const int* __restrict__ idx,
                                                            const float* __restrict__ wts, int c, int n, int m,
                                                            float* __restrict__ gx) {
  const int b = blockIdx.z, ch = blockIdx.y;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int* ib = idx + (size_t)b * 3 * n;
  const float* wb = wts + (size_t)b * 3 * n;
  float g = gy[((size_t)b * c + ch) * n + j];
  float* o = gx + ((size_t)b * c + ch) * m;
  atomicAdd(o + ib[j], g * wb[j]);
  atomicAdd(o + ib[j + n], g * wb[j + n]);
  atomicAdd(o + ib[j + 2 * n], g * wb[j + 2 * n]);
}

// ------------------------------------------------------------------------------------------------
// voxelisation                                   R/voxelization/vox.cu:18-72 (+ voxelization.py:16-35 when FUSED)
// One block per cloud; per-voxel state (count, list tail) lives in shared memory as 16-bit values.
// Phase B threads the points of each voxel into a linked list in ascending point index: 32-point chunks
// are processed in order (one warp per chunk), match.any finds the in-chunk predecessor and s_tail
// carries the link across chunks.  Phase C gives every (list head, channel) pair to one thread, which
// walks the list and sums in index order - the order the oracle uses - so the averages are
// bit-reproducible (the reference's float atomics are not) and no global atomics are issued.
// ------------------------------------------------------------------------------------------------
constexpr int kVoxMaxPPT = 16;      // 512 threads x 16 -> n <= 8192 points per cloud
constexpr int kVoxMaxPoints = 8192;
template <bool FUSED>
__global__ void __launch_bounds__(512) voxelize_kernel(const float* __restrict__ feat,
                                                        const void* __restrict__ coords_in, int c, int n, int r,
                                                        float* __restrict__ out, int* __restrict__ ind_out,
                                                        int* __restrict__ cnt_out, float* __restrict__ norm_out,
                                                        int* __restrict__ vox_out) {
  extern __shared__ unsigned char s_raw[];
  __shared__ double s_red[3][16];
  __shared__ float s_mean[3];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int r2 = r * r, r3 = r2 * r;
  // carve: cnt u16[r3] | tail i16[r3] | next i16[n] | vox i32[n]
  unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_raw);
  short* s_tail = reinterpret_cast<short*>(s_cnt + ((r3 + 1) & ~1));
  short* s_next = s_tail + ((r3 + 1) & ~1);
  int* s_vox = reinterpret_cast<int*>(s_next + ((n + 1) & ~1));
  // blockIdx.y owns a slice of the channels (the point binning is cheap and repeated per slice); slice 0 also writes
  // the per-point / per-voxel side outputs
  const int cs = (c + gridDim.y - 1) / gridDim.y, ch_lo = blockIdx.y * cs, ch_n = max(0, min(c, ch_lo + cs) - ch_lo);
  const bool side = blockIdx.y == 0;
  const float* fb = feat + ((size_t)b * c + ch_lo) * n;
  float* ob = out + ((size_t)b * c + ch_lo) * r3;

  for (int i = tid; i < r3; i += nthreads) { s_cnt[i] = 0; s_tail[i] = -1; }
  for (int i = tid; i < n; i += nthreads) s_next[i] = -1;
  // zero the output grid (the reference relies on torch::zeros); 128-bit stores when aligned
  {
    size_t tot = (size_t)ch_n * r3;
    if ((tot & 3) == 0 && (reinterpret_cast<uintptr_t>(ob) & 15) == 0) {
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (size_t i = tid; i < tot / 4; i += nthreads) reinterpret_cast<float4*>(ob)[i] = z;
    } else {
      for (size_t i = tid; i < tot; i += nthreads) ob[i] = 0.f;
    }
  }
  const int ppt = (n + nthreads - 1) / nthreads;
  if (FUSED) {
    const float* cf = reinterpret_cast<const float*>(coords_in) + (size_t)b * 3 * n;
    // mean over the point axis, accumulated in double (voxelization.py:18)
    double sx = 0, sy = 0, sz = 0;
    for (int i = tid; i < n; i += nthreads) { sx += cf[i]; sy += cf[i + n]; sz += cf[i + 2 * n]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) { s_red[0][wid] = sx; s_red[1][wid] = sy; s_red[2][wid] = sz; }
    __syncthreads();
    if (wid == 0) {
      double a = lane < nwarps ? s_red[0][lane] : 0, bb = lane < nwarps ? s_red[1][lane] : 0,
             cc = lane < nwarps ? s_red[2][lane] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        bb += __shfl_xor_sync(0xffffffffu, bb, o);
        cc += __shfl_xor_sync(0xffffffffu, cc, o);
      }
      if (lane == 0) { s_mean[0] = (float)(a / n); s_mean[1] = (float)(bb / n); s_mean[2] = (float)(cc / n); }
    }
    __syncthreads();
    const float rf = (float)r, hi = (float)(r - 1);
    for (int i = tid; i < n; i += nthreads) {
      int v3[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float x = __fsub_rn(cf[i + a * n], s_mean[a]);
        x = __fmul_rn(__fadd_rn(x, 1.0f), 0.5f);               // (x + 1) / 2.0
        x = fminf(fmaxf(__fmul_rn(x, rf), 0.f), hi);             // clamp(x * r, 0, r-1)
        if (side) norm_out[((size_t)b * 3 + a) * n + i] = x;
        v3[a] = (int)rintf(x);                                   // torch.round: half to even
        if (vox_out && side) vox_out[((size_t)b * 3 + a) * n + i] = v3[a];
      }
      s_vox[i] = v3[0] * r2 + v3[1] * r + v3[2];
    }
  } else {
    const int* ci = reinterpret_cast<const int*>(coords_in) + (size_t)b * 3 * n;
    for (int i = tid; i < n; i += nthreads) s_vox[i] = ci[i] * r2 + ci[i + n] * r + ci[i + 2 * n];
  }
  __syncthreads();
  // Phase B: ordered chunks.  Chunk q covers points [32q, 32q+32); warp (q mod nwarps) owns it and the
  // block barrier between consecutive groups of nwarps chunks keeps the order.
  const int nchunks = (n + 31) >> 5;
  for (int q0 = 0; q0 < nchunks; q0 += nwarps) {
    for (int w = 0; w < nwarps; ++w) {
      if (wid == w && q0 + w < nchunks) {
        const int i = ((q0 + w) << 5) + lane;
        const bool valid = i < n;
        const int v = valid ? s_vox[i] : -1;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
          const unsigned same = __match_any_sync(act, v);
          const unsigned below = same & ((1u << lane) - 1u);
          const int base = s_cnt[v];
          const int tail = s_tail[v];
          __syncwarp(act);
          const int pred = below ? i - (lane - (31 - __clz(below))) : tail;
          if (pred >= 0) s_next[pred] = (short)i;
          if (below == 0) s_cnt[v] = (unsigned short)(base + __popc(same));
          if ((same >> lane) == 1u) s_tail[v] = (short)i;       // highest lane of the group
        }
      }
      __syncthreads();
    }
  }
  if (ind_out && side)
    for (int i = tid; i < n; i += nthreads) ind_out[(size_t)b * n + i] = s_vox[i];
  if (cnt_out && side)
    for (int i = tid; i < r3; i += nthreads) cnt_out[(size_t)b * r3 + i] = s_cnt[i];
  __syncthreads();
  // Phase C: item (ch, i) with i a list head (the first point of its voxel, i.e. no point links to it).
  // Successors are marked by setting bit 30 of their s_vox entry.
  for (int i = tid; i < n; i += nthreads) {
    int nx = s_next[i];
    if (nx >= 0) atomicOr(&s_vox[nx], 0x40000000);               // successor is not a head
  }
  __syncthreads();
  const int items = ch_n * n;
  for (int it = tid; it < items; it += nthreads) {
    const int ch = it / n, i = it - ch * n;
    const int vv = s_vox[i];
    if (vv & 0x40000000) continue;
    const float div = (float)(1.0 / (double)(float)s_cnt[vv]);    // vox.cu:65
    const float* f = fb + (size_t)ch * n;
    float acc = 0.f;
    for (int p = i; p >= 0; p = s_next[p]) acc = __fadd_rn(acc, __fmul_rn(__ldg(f + p), div));
    ob[(size_t)ch * r3 + vv] = acc;
  }
}

__global__ void __launch_bounds__(256) avg_voxelize_grad_kernel(const float* __restrict__ gy,
                                                                const int* __restrict__ ind,
                                                                const int* __restrict__ cnt, int c, int n,
                                                                int r3, float* __restrict__ gx) {
  const int b = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int pos = ind[(size_t)b * n + i];
  int cc = cnt[(size_t)b * r3 + pos];
  float div = cc > 0 ? (float)(1.0 / (double)(float)cc) : 0.f;
  for (int ch = 0; ch < c; ++ch)
    gx[((size_t)b * c + ch) * n + i] = cc > 0 ? gy[((size_t)b * c + ch) * r3 + pos] * div : 0.f;
}

// ------------------------------------------------------------------------------------------------
// trilinear devoxelisation                       R/interpolate/trilinear_devox.cu:21-105
// ------------------------------------------------------------------------------------------------
struct TriCorner {
  int idx[8];
  float w[8];
};
__device__ __forceinline__ TriCorner tri_setup(float x, float y, float z, int r, int r2) {
  TriCorner t;
  float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  float xd1 = x - xl, yd1 = y - yl, zd1 = z - zl;
  float xd0 = 1.0f - xd1, yd0 = 1.0f - yd1, zd0 = 1.0f - zd1;
  t.w[0] = __fmul_rn(__fmul_rn(xd0, yd0), zd0);
  t.w[1] = __fmul_rn(__fmul_rn(xd0, yd0), zd1);
  t.w[2] = __fmul_rn(__fmul_rn(xd0, yd1), zd0);
  t.w[3] = __fmul_rn(__fmul_rn(xd0, yd1), zd1);
  t.w[4] = __fmul_rn(__fmul_rn(xd1, yd0), zd0);
  t.w[5] = __fmul_rn(__fmul_rn(xd1, yd0), zd1);
  t.w[6] = __fmul_rn(__fmul_rn(xd1, yd1), zd0);
  t.w[7] = __fmul_rn(__fmul_rn(xd1, yd1), zd1);
  int xh = xd1 > 0 ? r2 : 0, yh = yd1 > 0 ? r : 0, zh = zd1 > 0 ? 1 : 0;
  t.idx[0] = (int)xl * r2 + (int)yl * r + (int)zl;
  t.idx[1] = t.idx[0] + zh;
  t.idx[2] = t.idx[0] + yh;
  t.idx[3] = t.idx[2] + zh;
  t.idx[4] = t.idx[0] + xh;
  t.idx[5] = t.idx[4] + zh;
  t.idx[6] = t.idx[4] + yh;
  t.idx[7] = t.idx[6] + zh;
  return t;
}
// reference SASS order: FMUL w001*f001, FFMA w000*f000, then FFMA 010 ... 111
__device__ __forceinline__ float tri_eval(const TriCorner& t, const float* __restrict__ f) {
  float acc = __fmul_rn(t.w[1], __ldg(f + t.idx[1]));
  acc = __fmaf_rn(t.w[0], __ldg(f + t.idx[0]), acc);
#pragma unroll
  for (int k = 2; k < 8; ++k) acc = __fmaf_rn(t.w[k], __ldg(f + t.idx[k]), acc);
  return acc;
}

constexpr int kDevoxChanTile = 8;
__global__ void __launch_bounds__(128) devoxelize_kernel(const float* __restrict__ coords,
                                                         const float* __restrict__ feat, int c, int n, int r,
                                                         bool training, float* __restrict__ outs,
                                                         int* __restrict__ inds, float* __restrict__ wgts) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r2 = r * r, r3 = r2 * r;
  const float* cb = coords + (size_t)b * 3 * n;
  TriCorner t = tri_setup(cb[i], cb[i + n], cb[i + 2 * n], r, r2);
  if (training && blockIdx.y == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      wgts[((size_t)b * 8 + k) * n + i] = t.w[k];
      inds[((size_t)b * 8 + k) * n + i] = t.idx[k];
    }
  }
  const int c0 = blockIdx.y * kDevoxChanTile, c1 = min(c, c0 + kDevoxChanTile);
  for (int ch = c0; ch < c1; ++ch)
    outs[((size_t)b * c + ch) * n + i] = tri_eval(t, feat + ((size_t)b * c + ch) * r3);
}

__global__ void __launch_bounds__(128) devoxelize_grad_kernel(const float* __restrict__ gy,
                                                              const int* __restrict__ inds,
                                                              const float* __restrict__ wgts, int c, int n,
                                                              int r3, float* __restrict__ gx) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int id[8];
  float w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    id[k] = inds[((size_t)b * 8 + k) * n + i];
    w[k] = wgts[((size_t)b * 8 + k) * n + i];
  }
  for (int ch = blockIdx.y * kDevoxChanTile; ch < min(c, (int)(blockIdx.y + 1) * kDevoxChanTile); ++ch) {
    float g = gy[((size_t)b * c + ch) * n + i];
    float* o = gx + ((size_t)b * c + ch) * r3;
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(o + id[k], w[k] * g);
  }
}

}  // namespace gldm

// =================================================================================================
// C ABI
// =================================================================================================
using namespace gldm;

extern "C" int gldm_furthest_point_sampling(const float* coords, int b, int n, int m, int* indices,
                                            void* stream) {
  GLDM_REQUIRE(b <= 0 || m <= 0 || (coords && indices), "furthest_point_sampling: null pointer");
  GLDM_REQUIRE(b >= 0 && n > 0 && m >= 0, "furthest_point_sampling: bad sizes b=%d n=%d m=%d", b, n, m);
  GLDM_REQUIRE(n < (1 << 22), "furthest_point_sampling: n=%d exceeds 2^22", n);
  if (b == 0 || m == 0) return GLDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int threads = min(1024, ceil_div(n, 32) * 32);
  int ppt = ceil_div(n, threads);
#define FPS_LAUNCH(P) fps_kernel<P><<<b, threads, 0, s>>>(coords, n, m, indices)
  if (ppt <= 1) FPS_LAUNCH(1);
  else if (ppt <= 2) FPS_LAUNCH(2);
  else if (ppt <= 4) FPS_LAUNCH(4);
  else if (ppt <= 8) FPS_LAUNCH(8);
  else if (ppt <= 16) FPS_LAUNCH(16);
  else if (ppt <= 32) FPS_LAUNCH(32);
  else {
    set_error("furthest_point_sampling: n=%d > 32768 points per cloud is not supported", n);
    return GLDM_ENOSUP;
  }
#undef FPS_LAUNCH
  return check_launch("fps_kernel");
}

extern "C" int gldm_ball_query(const float* centers, const float* points, int b, int n, int m, float radius,
                               int u, int* neighbors, void* stream) {
  GLDM_REQUIRE(b <= 0 || (centers && points && neighbors), "ball_query: null pointer");
  GLDM_REQUIRE(b >= 0 && n > 0 && m > 0 && u > 0, "ball_query: bad sizes");
  if (b == 0) return GLDM_OK;
  float r2 = radius * radius;   // ball_query.cpp:24 (float multiply)
  dim3 grid(ceil_div(m, kBqCentersPerBlock), b);
  ball_query_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(centers, points, n, m, r2, u, neighbors);
  return check_launch("ball_query_kernel");
}

static int launch_group_gather(const float* feat, const int* idx, int b, int c, int n, int mu, float* out,
                               cudaStream_t s, const char* what) {
  bool vec = (mu % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  dim3 grid(ceil_div(vec ? mu / 4 : mu, 256), ceil_div(c, kGroupChanTile), b);
  if (vec) group_gather_kernel<true><<<grid, 256, 0, s>>>(feat, idx, c, n, mu, out);
  else group_gather_kernel<false><<<grid, 256, 0, s>>>(feat, idx, c, n, mu, out);
  return check_launch(what);
}

extern "C" int gldm_grouping_forward(const float* features, const int* indices, int b, int c, int n, int m,
                                     int u, float* out, void* stream) {
  GLDM_REQUIRE(b <= 0 || (features && indices && out), "grouping_forward: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && n > 0 && m > 0 && u > 0, "grouping_forward: bad sizes");
  if (b == 0) return GLDM_OK;
  return launch_group_gather(features, indices, b, c, n, m * u, out, (cudaStream_t)stream, "grouping_kernel");
}

extern "C" int gldm_gather_features_forward(const float* features, const int* indices, int b, int c, int n,
                                            int m, float* out, void* stream) {
  GLDM_REQUIRE(b <= 0 || (features && indices && out), "gather_features_forward: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && n > 0 && m > 0, "gather_features_forward: bad sizes");
  if (b == 0) return GLDM_OK;
  return launch_group_gather(features, indices, b, c, n, m, out, (cudaStream_t)stream, "gather_kernel");
}

static int launch_scatter_add(const float* gy, const int* idx, int b, int c, int n, int mu, float* gx,
                              cudaStream_t s, const char* what) {
  if (cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)b * c * n, s) != cudaSuccess) {
    set_error("%s: memset failed", what);
    return GLDM_ECUDA;
  }
  dim3 grid(ceil_div(mu, 256), ceil_div(c, kGroupChanTile), b);
  group_scatter_add_kernel<<<grid, 256, 0, s>>>(gy, idx, c, n, mu, gx);
  return check_launch(what);
}

extern "C" int gldm_grouping_backward(const float* grad_y, const int* indices, int b, int c, int n, int m,
                                      int u, float* grad_x, void* stream) {
  GLDM_REQUIRE(grad_y && indices && grad_x, "grouping_backward: null pointer");
  GLDM_REQUIRE(b > 0 && c > 0 && n > 0 && m > 0 && u > 0, "grouping_backward: bad sizes");
  return launch_scatter_add(grad_y, indices, b, c, n, m * u, grad_x, (cudaStream_t)stream, "grouping_grad");
}

extern "C" int gldm_gather_features_backward(const float* grad_y, const int* indices, int b, int c, int n,
                                             int m, float* grad_x, void* stream) {
  GLDM_REQUIRE(grad_y && indices && grad_x, "gather_features_backward: null pointer");
  GLDM_REQUIRE(b > 0 && c > 0 && n > 0 && m > 0, "gather_features_backward: bad sizes");
  return launch_scatter_add(grad_y, indices, b, c, n, m, grad_x, (cudaStream_t)stream, "gather_grad");
}

extern "C" int gldm_three_nn_interpolate_forward(const float* points, const float* centers, const float* feats,
                                                 int b, int c, int m, int n, float* out, int* idx, float* w,
                                                 void* stream) {
  GLDM_REQUIRE(b <= 0 || (points && centers && feats && out && idx && w), "three_nn_interpolate_forward: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && m > 0 && n > 0, "three_nn_interpolate_forward: bad sizes");
  if (b == 0) return GLDM_OK;
  dim3 grid(ceil_div(n, 128), b);
  three_nn_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(points, centers, feats, c, m, n, out, idx, w);
  return check_launch("three_nn_kernel");
}

extern "C" int gldm_three_nn_interpolate_backward(const float* grad_y, const int* idx, const float* w, int b,
                                                  int c, int n, int m, float* grad_x, void* stream) {
  GLDM_REQUIRE(grad_y && idx && w && grad_x, "three_nn_interpolate_backward: null pointer");
  GLDM_REQUIRE(b > 0 && c > 0 && m > 0 && n > 0, "three_nn_interpolate_backward: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * m, s) != cudaSuccess) {
    set_error("three_nn_interpolate_backward: memset failed");
    return GLDM_ECUDA;
  }
  dim3 grid(ceil_div(n, 256), c, b);
  three_nn_grad_kernel<<<grid, 256, 0, s>>>(grad_y, idx, w, c, n, m, grad_x);
  return check_launch("three_nn_grad_kernel");
}

static int launch_voxelize(bool fused, const float* feat, const void* coords, int b, int c, int n, int r,
                           float* out, int* ind, int* cnt, float* norm, int* vox, cudaStream_t s) {
  GLDM_REQUIRE(b <= 0 || (feat && coords && out), "voxelize: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && n > 0 && r > 0, "voxelize: bad sizes b=%d c=%d n=%d r=%d", b, c, n, r);
  GLDM_REQUIRE(n <= kVoxMaxPoints, "voxelize: n=%d > %d points per cloud not supported", n, kVoxMaxPoints);
  GLDM_REQUIRE(r <= 36, "voxelize: resolution %d > 36 not supported (shared-memory voxel table)", r);
  if (b == 0) return GLDM_OK;
  const size_t r3 = (size_t)r * r * r;
  size_t smem = 2 * ((r3 + 1) & ~(size_t)1) * 2 + (((size_t)n + 1) & ~(size_t)1) * 2 + (size_t)n * 4;
  int threads = min(512, ceil_div(n, 32) * 32);
  const int slices = min(8, ceil_div(c, 8));      // channel slices per cloud: more CTAs than clouds for wide features
  if (fused) {
    static SmemOptIn attr;
    if (int rc = opt_in_smem(attr, voxelize_kernel<true>, 200 * 1024, "voxelize_kernel<fused>")) return rc;
    voxelize_kernel<true><<<dim3(b, slices), threads, smem, s>>>(feat, coords, c, n, r, out, ind, cnt, norm, vox);
  } else {
    static SmemOptIn attr;
    if (int rc = opt_in_smem(attr, voxelize_kernel<false>, 200 * 1024, "voxelize_kernel")) return rc;
    voxelize_kernel<false><<<dim3(b, slices), threads, smem, s>>>(feat, coords, c, n, r, out, ind, cnt, norm, vox);
  }
  return check_launch("voxelize_kernel");
}

extern "C" int gldm_avg_voxelize_forward(const float* features, const int* coords, int b, int c, int n, int r,
                                         float* out, int* ind, int* cnt, void* stream) {
  return launch_voxelize(false, features, coords, b, c, n, r, out, ind, cnt, nullptr, nullptr,
                         (cudaStream_t)stream);
}

extern "C" int gldm_voxelize_fused(const float* features, const float* coords, int b, int c, int n, int r,
                                   float* grid, float* norm_coords, int* vox, void* stream) {
  GLDM_REQUIRE(norm_coords, "voxelize_fused: null norm_coords");
  return launch_voxelize(true, features, coords, b, c, n, r, grid, nullptr, nullptr, norm_coords, vox,
                         (cudaStream_t)stream);
}

extern "C" int gldm_avg_voxelize_backward(const float* grad_y, const int* ind, const int* cnt, int b, int c,
                                          int n, int r3, float* grad_x, void* stream) {
  GLDM_REQUIRE(grad_y && ind && cnt && grad_x, "avg_voxelize_backward: null pointer");
  GLDM_REQUIRE(b > 0 && c > 0 && n > 0 && r3 > 0, "avg_voxelize_backward: bad sizes");
  dim3 grid(ceil_div(n, 256), b);
  avg_voxelize_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_y, ind, cnt, c, n, r3, grad_x);
  return check_launch("avg_voxelize_grad_kernel");
}

extern "C" int gldm_trilinear_devoxelize_forward(const float* coords, const float* features, int b, int c,
                                                 int n, int r, int is_training, float* outs, int* inds,
                                                 float* wgts, void* stream) {
  GLDM_REQUIRE(b <= 0 || (coords && features && outs), "trilinear_devoxelize_forward: null pointer");
  GLDM_REQUIRE(!is_training || (inds && wgts), "trilinear_devoxelize_forward: training needs inds/wgts");
  GLDM_REQUIRE(b >= 0 && c > 0 && n > 0 && r > 0, "trilinear_devoxelize_forward: bad sizes");
  if (b == 0) return GLDM_OK;
  dim3 grid(ceil_div(n, 128), ceil_div(c, kDevoxChanTile), b);
  devoxelize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(coords, features, c, n, r, is_training != 0, outs,
                                                            inds, wgts);
  return check_launch("devoxelize_kernel");
}

extern "C" int gldm_trilinear_devoxelize_backward(const float* grad_y, const int* inds, const float* wgts,
                                                  int b, int c, int n, int r3, float* grad_x, void* stream) {
  GLDM_REQUIRE(grad_y && inds && wgts && grad_x, "trilinear_devoxelize_backward: null pointer");
  GLDM_REQUIRE(b > 0 && c > 0 && n > 0 && r3 > 0, "trilinear_devoxelize_backward: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * r3, s) != cudaSuccess) {
    set_error("trilinear_devoxelize_backward: memset failed");
    return GLDM_ECUDA;
  }
  dim3 grid(ceil_div(n, 128), ceil_div(c, kDevoxChanTile), b);
  devoxelize_grad_kernel<<<grid, 128, 0, s>>>(grad_y, inds, wgts, c, n, r3, grad_x);
  return check_launch("devoxelize_grad_kernel");
}
