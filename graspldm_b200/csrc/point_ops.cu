// Point-cloud operators of the GraspLDM encoder backbone, written for sm_100a.
//
// C-ABI replacements for the reference's `_pvcnn_backend` extension
// (R = /root/reference/grasp_ldm/models/modules/ext/pvcnn/modules/functional/src).
// The reference launches one 512-thread block per cloud for every op, uses global float atomics
// and ~10 block barriers per FPS round; here
//   * FPS keeps every point and its running min-distance in registers, picks the round winner with
//     two REDUX (warp max of the distance bits, warp min of the tie-break key) per level and one
//     block barrier per round;
//   * ball query has two kernels: small batches take four centres per warp (ballot / popc compaction, early exit),
//     large ones a thread per centre that scans 32 points per round branch-free into a hit mask and walks only the hits;
//   * 3-NN keeps the three best distances in fp32 registers with a branch-free insertion (same decisions as the
//     reference's double-typed bests, which only ever hold floats);
//   * grouping / gather are 128-bit vectorised, write-coalesced gathers;
//   * voxelisation buckets the points by voxel with a shared-memory counting sort, orders every bucket by point index
//     and accumulates in that order, so results are deterministic.
#include <cuda_bf16.h>

#include "common.cuh"

namespace gldm {

// ------------------------------------------------------------------------------------------------
// furthest point sampling                        R/sampling/sampling.cu:86-167
// ------------------------------------------------------------------------------------------------
// Tie-break contract of the reference (512-thread strided scan + pairwise tree that keeps the lower
// slot unless strictly greater): winner = max min-distance; ties -> smallest (k mod 512), then
// smallest k.  Encoded as a 32-bit key whose minimum wins.
__device__ __forceinline__ unsigned fps_key(int k) { return ((unsigned)(k & 511) << 22) | (unsigned)k; }

// One CTA per cloud with FOUR points per thread where the cloud allows it (256 threads for 1024 points): a round is
// distance update -> warp arg-max (two REDUX) -> one block barrier -> second-level arg-max over <= 32 warp results.
// The winner's coordinates come from a shared-memory copy of the cloud (broadcast LDS, ~30 cycles) instead of a
// global load (~an L2 round trip per round), and the small CTAs let 8 clouds share an SM at large batch.
template <int PPT, bool SMEM>
__global__ void __launch_bounds__(1024) fps_kernel(const float* __restrict__ coords, int n, int m,
                                                   int* __restrict__ indices) {
  if (m <= 0) return;
  extern __shared__ float s_pts[];           // SMEM: [3][n]
  const int b = blockIdx.x;
  const float* cx = coords + (size_t)b * 3 * n;
  const float* cy = cx + n;
  const float* cz = cy + n;
  int* out = indices + (size_t)b * m;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  __shared__ unsigned s_bits[2][32];
  __shared__ unsigned s_key[2][32];
  // slots of warps that do not exist hold the neutral element, so the second level reads all 32 without a branch; the
  // slot addresses are formed once (inside the round loop the compiler rebuilt them, ~20 instructions per round)
  if (tid < 64) {
    (&s_bits[0][0])[tid] = 0u;
    (&s_key[0][0])[tid] = 0xffffffffu;
  }
  // (32-bit shared-space addresses used through ld / st.shared below: with C++ pointers the cluster-window base was
  // rebuilt from %cluster_ctarank inside every round)
  unsigned wr_bits = (unsigned)__cvta_generic_to_shared(&s_bits[0][wid]);
  unsigned wr_key = (unsigned)__cvta_generic_to_shared(&s_key[0][wid]);
  unsigned rd_bits = (unsigned)__cvta_generic_to_shared(&s_bits[0][lane]);
  unsigned rd_key = (unsigned)__cvta_generic_to_shared(&s_key[0][lane]);
  asm volatile("" : "+r"(wr_bits), "+r"(wr_key), "+r"(rd_bits), "+r"(rd_key));   // opaque: keep them in registers

  float px[PPT], py[PPT], pz[PPT], dist[PPT];
  unsigned kk[PPT];                            // tie-break keys of this thread's points (0xffffffff: no point)
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    int k = tid + j * blockDim.x;
    bool v = k < n;
    px[j] = v ? cx[k] : 0.f;
    py[j] = v ? cy[k] : 0.f;
    pz[j] = v ? cz[k] : 0.f;
    dist[j] = v ? 1e38f : 0.f;   // sampling.cpp:53-54; a missing point can never exceed a real distance (>= 0) on bits
    kk[j] = v ? fps_key(k) : 0xffffffffu;
    if (SMEM && v) { s_pts[k] = px[j]; s_pts[n + k] = py[j]; s_pts[2 * n + k] = pz[j]; }
  }
  __syncthreads();                             // the staged cloud and the neutral slots above
  int old = 0;
  if (tid == 0) out[0] = 0;
  for (int r = 1; r < m; ++r) {
    const float x1 = SMEM ? s_pts[old] : __ldg(cx + old), y1 = SMEM ? s_pts[n + old] : __ldg(cy + old),
                z1 = SMEM ? s_pts[2 * n + old] : __ldg(cz + old);
    // running min-distances (d2 >= +0: bit order == value order), the thread's maximum first, then the smallest key
    // among the points that reach it (ties inside a thread follow the same key order as across threads)
    unsigned bits = 0u;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const float d = sqdist_ref(px[j] - x1, py[j] - y1, pz[j] - z1);
      dist[j] = fminf(d, dist[j]);
      bits = max(bits, __float_as_uint(dist[j]));
    }
    unsigned key = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < PPT; ++j) key = min(key, __float_as_uint(dist[j]) == bits ? kk[j] : 0xffffffffu);
    // warp level: max distance, then min key among the maxima
    unsigned gb = __reduce_max_sync(0xffffffffu, bits);
    unsigned gk = __reduce_min_sync(0xffffffffu, bits == gb ? key : 0xffffffffu);
    if (nwarps > 1) {
      const unsigned buf = (r & 1) * 128u;
      if (lane == 0) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(wr_bits + buf), "r"(gb) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(wr_key + buf), "r"(gk) : "memory");
      }
      __syncthreads();
      unsigned b2, k2;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b2) : "r"(rd_bits + buf) : "memory");
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(k2) : "r"(rd_key + buf) : "memory");
      gb = __reduce_max_sync(0xffffffffu, b2);
      gk = __reduce_min_sync(0xffffffffu, (b2 == gb) ? k2 : 0xffffffffu);
    }
    old = (gk == 0xffffffffu) ? 0 : (int)(gk & 0x3fffffu);
    if (tid == 0) out[r] = old;
  }
}

// ------------------------------------------------------------------------------------------------
// ball query                                     R/ball_query/ball_query.cu:19-50
// ------------------------------------------------------------------------------------------------
// Thread = centre: a CTA stages the cloud once in shared memory (12 n bytes) and every thread scans it in ascending
// index for its own centre.  The scan is branch-free: 32 points per round, four points per broadcast LDS.128 and
// coordinate, the distance chain as packed f32x2 operations (FADD2 / FMUL2 / FFMA2 round per element like the scalar
// chain, so the indices stay bit-exact), and the test "d < r2" as the SIGN of the packed difference d - r2 shifted
// into a per-thread hit mask (one funnel shift per point; d, r2 >= 0 and x - x = +0, so the sign is set exactly when
// d < r2).  Only the hits are then walked (count-leading-zeros over the mask, ascending index): the divergent append
// runs once per hit of the busiest lane of a 32-point round instead of once per point pair with a hit in any lane
// (which was 40 % of the pairs at r = 0.2).  The tail of every list (slots the reference pre-fills with the first
// hit) is written by the whole CTA, coalesced.  ~6 thread instructions per (centre, point) pair (was 16).
constexpr int kBqThreads = 128;
constexpr int kBqMaxSmemPoints = 16384;
template <bool SMEM>
__global__ void __launch_bounds__(kBqThreads) ball_query_kernel(const float* __restrict__ centers,
                                                                const float* __restrict__ points, int n, int m,
                                                                float r2, int u, int* __restrict__ out) {
  extern __shared__ __align__(16) float s_pts[];   // SMEM: [3][np], np = n rounded up to 32 (padding masked out below)
  __shared__ int s_cnt[kBqThreads], s_first[kBqThreads];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int np = (n + 31) & ~31;
  const float* gx = points + (size_t)b * 3 * n;
  if (SMEM) {
    for (int k = tid; k < np; k += kBqThreads) {
      const bool v = k < n;
      s_pts[k] = v ? gx[k] : 0.f;
      s_pts[np + k] = v ? gx[n + k] : 0.f;
      s_pts[2 * np + k] = v ? gx[2 * n + k] : 0.f;
    }
    __syncthreads();
  }
  const int j = blockIdx.x * kBqThreads + tid;
  const bool valid = j < m;
  const float* cc = centers + (size_t)b * 3 * m;
  const float c0 = valid ? __ldg(cc + j) : 0.f, c1 = valid ? __ldg(cc + m + j) : 0.f, c2 = valid ? __ldg(cc + 2 * m + j) : 0.f;
  const float2 cx = make_float2(c0, c0), cy = make_float2(c1, c1), cz = make_float2(c2, c2);
  const float2 nr2 = make_float2(-r2, -r2);
  int* o = out + ((size_t)b * m + (valid ? j : 0)) * u;
  int cnt = valid ? 0 : u, first = 0;
  for (int base = 0; base < np; base += 32) {
    unsigned mask = 0;                                  // point base + i -> bit 31 - i
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = base + 4 * q;
      float4 x, y, z;
      if (SMEM) {
        x = *reinterpret_cast<const float4*>(s_pts + k);
        y = *reinterpret_cast<const float4*>(s_pts + np + k);
        z = *reinterpret_cast<const float4*>(s_pts + 2 * np + k);
      } else {
        float t[12];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int e = 0; e < 4; ++e) t[4 * a + e] = k + e < n ? __ldg(gx + (size_t)a * n + k + e) : 0.f;
        x = make_float4(t[0], t[1], t[2], t[3]);
        y = make_float4(t[4], t[5], t[6], t[7]);
        z = make_float4(t[8], t[9], t[10], t[11]);
      }
      // d = fma(dz, dz, fma(dx, dx, dy * dy)) with d* = centre - point (ball_query.cu:36-40, FMA chain of the reference build)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 px = h ? make_float2(x.z, x.w) : make_float2(x.x, x.y);
        const float2 py = h ? make_float2(y.z, y.w) : make_float2(y.x, y.y);
        const float2 pz = h ? make_float2(z.z, z.w) : make_float2(z.x, z.y);
        const float2 dx = __fadd2_rn(cx, make_float2(-px.x, -px.y)), dy = __fadd2_rn(cy, make_float2(-py.x, -py.y)),
                     dz = __fadd2_rn(cz, make_float2(-pz.x, -pz.y));
        const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
        const float2 e = __fadd2_rn(d, nr2);            // sign(e) = (d < r2)
        mask = __funnelshift_l(__float_as_uint(e.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(e.y), mask, 1);
      }
    }
    if (base + 32 > n) mask &= ~0u << (base + 32 - n);  // padding points are never hits
    if (cnt >= u) mask = 0;
    while (mask) {
      const int i = __clz(mask);
      mask &= ~(0x80000000u >> i);
      if (cnt == 0) first = base + i;
      o[cnt] = base + i;
      if (++cnt >= u) mask = 0;
    }
    if (__all_sync(0xffffffffu, cnt >= u)) break;
  }
  // the first hit pre-fills every slot; no hit leaves zeros (ball_query.cu:39-44): tails written by the whole CTA
  s_cnt[tid] = cnt;
  s_first[tid] = first;
  __syncthreads();
  const int nc = min(kBqThreads, m - (int)blockIdx.x * kBqThreads);
  int* ob = out + ((size_t)b * m + (size_t)blockIdx.x * kBqThreads) * u;
  for (int idx = tid; idx < nc * u; idx += kBqThreads) {
    const int c = idx / u, slot = idx - c * u;
    if (slot >= s_cnt[c]) ob[idx] = s_first[c];
  }
}

// Small batches (latency): the warps of a CTA take the centres FOUR at a time; a 32-point chunk is read once (three LDS)
// and tested against four centres held in registers (two packed pairs), ballot / popc compaction of the hits in ascending
// point index, early exit when all four lists are full.  More parallelism per cloud than thread = centre, more
// instructions per test.
constexpr int kBqCentersPerBlock = 32;
template <bool SMEM>
__global__ void __launch_bounds__(256) ball_query_warp_kernel(const float* __restrict__ centers,
                                                              const float* __restrict__ points, int n, int m,
                                                              float r2, int u, int* __restrict__ out) {
  extern __shared__ __align__(16) float s_pts[];   // SMEM: [3][np], np = n rounded up to 32, the padding at +inf (never a hit)
  const int b = blockIdx.y;
  const int np = (n + 31) & ~31;
  const float* gx = points + (size_t)b * 3 * n;
  if (SMEM) {
    for (int k = threadIdx.x; k < np; k += blockDim.x) {
      const bool v = k < n;
      s_pts[k] = v ? gx[k] : INFINITY;
      s_pts[np + k] = v ? gx[n + k] : INFINITY;
      s_pts[2 * np + k] = v ? gx[2 * n + k] : INFINITY;
    }
    __syncthreads();
  }
  const float* cc = centers + (size_t)b * 3 * m;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int j_end = min(m, (int)(blockIdx.x + 1) * kBqCentersPerBlock);
  for (int j0 = blockIdx.x * kBqCentersPerBlock + 4 * wid; j0 < j_end; j0 += 4 * nw) {
    float2 cxa, cya, cza, cxb, cyb, czb;     // centres (j0, j0+1) and (j0+2, j0+3); a missing centre repeats the last one
    int jj[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) jj[q] = min(j0 + q, j_end - 1);
    cxa = make_float2(__ldg(cc + jj[0]), __ldg(cc + jj[1]));
    cya = make_float2(__ldg(cc + m + jj[0]), __ldg(cc + m + jj[1]));
    cza = make_float2(__ldg(cc + 2 * m + jj[0]), __ldg(cc + 2 * m + jj[1]));
    cxb = make_float2(__ldg(cc + jj[2]), __ldg(cc + jj[3]));
    cyb = make_float2(__ldg(cc + m + jj[2]), __ldg(cc + m + jj[3]));
    czb = make_float2(__ldg(cc + 2 * m + jj[2]), __ldg(cc + 2 * m + jj[3]));
    int cnt[4] = {0, 0, 0, 0}, first[4] = {0, 0, 0, 0};
    for (int base = 0; base < np; base += 32) {
      const int k = base + lane;
      float x, y, z;
      if (SMEM) {
        x = s_pts[k]; y = s_pts[np + k]; z = s_pts[2 * np + k];
      } else {
        const bool valid = k < n;
        x = valid ? __ldg(gx + k) : INFINITY;
        y = valid ? __ldg(gx + n + k) : INFINITY;
        z = valid ? __ldg(gx + 2 * n + k) : INFINITY;
      }
      const float2 nx = make_float2(-x, -x), ny = make_float2(-y, -y), nz = make_float2(-z, -z);
      const float2 dxa = __fadd2_rn(cxa, nx), dya = __fadd2_rn(cya, ny), dza = __fadd2_rn(cza, nz);
      const float2 dxb = __fadd2_rn(cxb, nx), dyb = __fadd2_rn(cyb, ny), dzb = __fadd2_rn(czb, nz);
      const float2 da = __ffma2_rn(dza, dza, __ffma2_rn(dxa, dxa, __fmul2_rn(dya, dya)));
      const float2 db = __ffma2_rn(dzb, dzb, __ffma2_rn(dxb, dxb, __fmul2_rn(dyb, dyb)));
      unsigned mk[4];
      mk[0] = __ballot_sync(0xffffffffu, da.x < r2);
      mk[1] = __ballot_sync(0xffffffffu, da.y < r2);
      mk[2] = __ballot_sync(0xffffffffu, db.x < r2);
      mk[3] = __ballot_sync(0xffffffffu, db.y < r2);
      if ((mk[0] | mk[1] | mk[2] | mk[3]) == 0u) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned mask = mk[q];
        if (mask && cnt[q] < u && j0 + q < j_end) {
          if (cnt[q] == 0) first[q] = base + __ffs(mask) - 1;
          const int pos = cnt[q] + __popc(mask & lt);
          if (((mask >> lane) & 1u) && pos < u) out[((size_t)b * m + j0 + q) * u + pos] = k;
          cnt[q] += __popc(mask);
        }
      }
      if (cnt[0] >= u && cnt[1] >= u && cnt[2] >= u && cnt[3] >= u) break;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (j0 + q < j_end) {
        int* o = out + ((size_t)b * m + j0 + q) * u;
        for (int v = min(cnt[q], u) + lane; v < u; v += 32) o[v] = first[q];
      }
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
// Thread = point.  The reference keeps the three best distances as doubles, but only ever stores floats in them and
// compares a float against them: float comparisons give the same decisions (the 1e40 start value behaves like +inf: any
// finite float is smaller, +inf and NaN are not), so the scan runs in fp32.  The centres are staged in shared memory
// (four per broadcast LDS.128 and coordinate), the distance chain is packed f32x2 (same rounding per element), and the
// insertion is branch-free (three compares, five selects, five min / max): with a divergent insert some lane of a warp
// took the branch on more than half of the centres.  Strict "<" as in the reference: an equal distance stays behind the
// earlier centre.  20 instead of 44 thread instructions per (point, centre) pair.
constexpr int kNnMaxSmemCenters = 4096;          // 3 x 16 KB: within the default dynamic shared-memory limit
template <bool SMEM>
__global__ void __launch_bounds__(128) three_nn_kernel(const float* __restrict__ points,
                                                       const float* __restrict__ centers,
                                                       const float* __restrict__ feats, int c, int m, int n,
                                                       float* __restrict__ out, int* __restrict__ idx,
                                                       float* __restrict__ wts) {
  extern __shared__ __align__(16) float s_ctr[];   // SMEM: [3][mp], mp = m rounded up to 4
  const int b = blockIdx.y;
  const int mp = (m + 3) & ~3;
  const float* cc = centers + (size_t)b * 3 * m;
  if (SMEM) {
    for (int k = threadIdx.x; k < mp; k += blockDim.x) {
      const bool v = k < m;
      s_ctr[k] = v ? cc[k] : 0.f;
      s_ctr[mp + k] = v ? cc[m + k] : 0.f;
      s_ctr[2 * mp + k] = v ? cc[2 * m + k] : 0.f;
    }
    __syncthreads();
  }
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* pp = points + (size_t)b * 3 * n;
  const float ux = pp[j], uy = pp[j + n], uz = pp[j + 2 * n];
  const float2 ux2 = make_float2(ux, ux), uy2 = make_float2(uy, uy), uz2 = make_float2(uz, uz);
  float b0 = INFINITY, b1 = INFINITY, b2 = INFINITY;
  int i0 = 0, i1 = 0, i2 = 0;
  auto insert = [&](float d, int k) {
    const bool lt0 = d < b0, lt1 = d < b1, lt2 = d < b2;
    i2 = lt1 ? i1 : (lt2 ? k : i2);
    i1 = lt0 ? i0 : (lt1 ? k : i1);
    i0 = lt0 ? k : i0;
    const float n2 = fmaxf(fminf(d, b2), b1), n1 = fmaxf(fminf(d, b1), b0);
    b0 = fminf(d, b0);
    b1 = n1;
    b2 = n2;
  };
  const int m4 = m & ~3;
  for (int k = 0; k < m4; k += 4) {
    float4 x, y, z;
    if (SMEM) {
      x = *reinterpret_cast<const float4*>(s_ctr + k);
      y = *reinterpret_cast<const float4*>(s_ctr + mp + k);
      z = *reinterpret_cast<const float4*>(s_ctr + 2 * mp + k);
    } else {
      x = make_float4(__ldg(cc + k), __ldg(cc + k + 1), __ldg(cc + k + 2), __ldg(cc + k + 3));
      y = make_float4(__ldg(cc + m + k), __ldg(cc + m + k + 1), __ldg(cc + m + k + 2), __ldg(cc + m + k + 3));
      z = make_float4(__ldg(cc + 2 * m + k), __ldg(cc + 2 * m + k + 1), __ldg(cc + 2 * m + k + 2), __ldg(cc + 2 * m + k + 3));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 dx = __fadd2_rn(ux2, h ? make_float2(-x.z, -x.w) : make_float2(-x.x, -x.y));
      const float2 dy = __fadd2_rn(uy2, h ? make_float2(-y.z, -y.w) : make_float2(-y.x, -y.y));
      const float2 dz = __fadd2_rn(uz2, h ? make_float2(-z.z, -z.w) : make_float2(-z.x, -z.y));
      const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));   // sqdist_ref, two centres
      insert(d.x, k + 2 * h);
      insert(d.y, k + 2 * h + 1);
    }
  }
  for (int k = m4; k < m; ++k)
    insert(sqdist_ref(ux - (SMEM ? s_ctr[k] : __ldg(cc + k)), uy - (SMEM ? s_ctr[mp + k] : __ldg(cc + m + k)),
                      uz - (SMEM ? s_ctr[2 * mp + k] : __ldg(cc + 2 * m + k))), k);
  double best0 = b0, best1 = b1, best2 = b2;
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
#pragma unroll 4
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
// One block per (cloud, channel slice, cell range); per-voxel state lives in shared memory as 16-bit values.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned gldm_pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&t);
}
constexpr int kVoxMaxPoints = 8192;  // point indices, bucket offsets and voxel ids (r <= 36) are 16-bit values in shared memory
// Bucket version (round 2).  The first version threaded the points of a voxel into a linked list, 32-point chunk after
// chunk (one block barrier per chunk: the order is what makes the sums reproducible), and its averaging phase gave a
// thread four voxels and all channels of the slice - ncu: 55 % of the executed instructions in divergent list walks whose
// length is set by the fullest voxels, 20 % of the stall samples in the serial chunk loop.  Now
//   B  counting sort: per-voxel counts by shared-memory atomics (the arrival slot is arbitrary), exclusive scan over the
//      voxels in place, scatter into buckets, then every point ranks itself inside its bucket by counting the smaller
//      indices - the bucket is in ascending index whatever order the atomics ran in: six barriers instead of n / 32;
//   C  empty cells are zero-filled coalesced; the non-empty ones are compacted and averaged as (cell, channel group) items
//      spread evenly over the threads: a planar item is four consecutive voxels of one channel (one 16-byte store), a
//      channels-last item is eight channels of one voxel (one 16-byte bf16 chunk).  Sums run in ascending point index, the
//      order the oracle uses, so results stay bit-reproducible (the reference's float atomics are not).
template <bool FUSED, int CH>
__global__ void __launch_bounds__(512) voxelize_kernel(const float* __restrict__ feat,
                                                        const void* __restrict__ coords_in, int c, int n, int r,
                                                        float* __restrict__ out, int* __restrict__ ind_out,
                                                        int* __restrict__ cnt_out, float* __restrict__ norm_out,
                                                        int* __restrict__ vox_out, __nv_bfloat16* __restrict__ cl_out,
                                                        int cl_stride, int stage_feat) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ double s_red[3][16];
  __shared__ float s_mean[3];
  __shared__ unsigned s_wsum[16];
  __shared__ int s_ncells;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int r2 = r * r, r3 = r2 * r;
  // carve (16-bit entries): off[r3 + 1] counts, then exclusive bucket offsets (off[r3] = n) | vox[n] | perm[n] buckets in
  // arrival order | sorted[n] arrival slot, then the buckets in ascending index | cells[n] non-empty cells
  const int offw = (r3 + 2 + 1) & ~1, ne = (n + 1) & ~1;
  unsigned short* s_off = reinterpret_cast<unsigned short*>(s_raw);
  unsigned short* s_vox = s_off + offw;
  unsigned short* s_perm = s_vox + ne;
  unsigned short* s_sorted = s_perm + ne;
  unsigned short* s_cells = s_sorted + ne;
  float* s_feat = reinterpret_cast<float*>(s_raw + (((size_t)(offw + 4 * ne) * 2 + 15) & ~(size_t)15));   // [ch_n][n] when staged
  // blockIdx.y owns a slice of the channels (the point binning is cheap and repeated per slice); slice 0 also writes
  // the per-point / per-voxel side outputs
  const int cs = (c + gridDim.y - 1) / gridDim.y, ch_lo = blockIdx.y * cs, ch_n = max(0, min(c, ch_lo + cs) - ch_lo);
  const bool side = blockIdx.y == 0 && blockIdx.z == 0;
  const float* fb = feat + ((size_t)b * c + ch_lo) * n;

  // the slice's features on their way into shared memory (asynchronous copies, consumed in phase C): the bucket sums
  // are chains of dependent loads - from global memory they were 44 % of the stall samples
  if (stage_feat) {
    const int nvec = ch_n * n / 4;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(s_feat);
    for (int i = tid; i < nvec; i += nthreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(fb + (size_t)i * 4));
    asm volatile("cp.async.commit_group;" ::);
  }
  for (int i = tid; i < offw / 2; i += nthreads) reinterpret_cast<unsigned*>(s_off)[i] = 0u;
  if (tid == 0) s_ncells = 0;
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
      s_vox[i] = (unsigned short)(v3[0] * r2 + v3[1] * r + v3[2]);
    }
  } else {
    const int* ci = reinterpret_cast<const int*>(coords_in) + (size_t)b * 3 * n;
    for (int i = tid; i < n; i += nthreads) s_vox[i] = (unsigned short)(ci[i] * r2 + ci[i + n] * r + ci[i + 2 * n]);
  }
  __syncthreads();
  // ---- B1: counts (two 16-bit counters per 32-bit word; n <= 8192 cannot carry into the upper half)
  for (int i = tid; i < n; i += nthreads) {
    const int v = s_vox[i], sh = (v & 1) << 4;
    const unsigned old = atomicAdd(reinterpret_cast<unsigned*>(s_off) + (v >> 1), 1u << sh);
    s_sorted[i] = (unsigned short)((old >> sh) & 0xffffu);
  }
  __syncthreads();
  // ---- B2: exclusive scan over the voxels, in place (a thread owns a contiguous segment)
  {
    const int seg = (r3 + nthreads - 1) / nthreads, lo = min(r3, tid * seg), hi = min(r3, lo + seg);
    unsigned sum = 0;
    for (int v = lo; v < hi; ++v) sum += s_off[v];
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    unsigned run = incl - sum;
    for (int w = 0; w < wid; ++w) run += s_wsum[w];
    for (int v = lo; v < hi; ++v) {
      const unsigned cv = s_off[v];
      s_off[v] = (unsigned short)run;
      run += cv;
    }
    if (tid == 0) s_off[r3] = (unsigned short)n;
  }
  __syncthreads();
  // ---- B3: buckets in arrival order;  B4: rank inside the bucket = number of smaller indices
  for (int i = tid; i < n; i += nthreads) s_perm[s_off[s_vox[i]] + s_sorted[i]] = (unsigned short)i;
  __syncthreads();
  for (int i = tid; i < n; i += nthreads) {
    const int v = s_vox[i], o = s_off[v], cv = s_off[v + 1] - o;
    int rank = 0;
    for (int j = 0; j < cv; ++j) rank += (int)(s_perm[o + j] < (unsigned short)i);
    s_sorted[o + rank] = (unsigned short)i;
  }
  if (ind_out && side)
    for (int i = tid; i < n; i += nthreads) ind_out[(size_t)b * n + i] = s_vox[i];
  __syncthreads();
  if (cnt_out && side)
    for (int i = tid; i < r3; i += nthreads) cnt_out[(size_t)b * r3 + i] = (int)s_off[i + 1] - (int)s_off[i];

  // ---- C: cells.  planar: four consecutive voxels (one voxel when the grid cannot be written in 16-byte pieces);
  //      channels-last: one voxel = one row of the zero-padded grid of this cloud
  float* ob = out ? out + ((size_t)b * c + ch_lo) * r3 : nullptr;
  const bool quads = !cl_out && (r3 & 3) == 0 && (reinterpret_cast<uintptr_t>(ob) & 15) == 0;
  const int cw = quads ? 4 : 1, ncell = r3 / cw;
  const int rp = r + 2;
  __nv_bfloat16* cb = cl_out ? cl_out + (size_t)b * rp * rp * rp * cl_stride + ch_lo : nullptr;
  auto cl_row = [&](int v) {
    const int x = v / r2, y = (v / r) % r, z = v % r;
    return reinterpret_cast<uint4*>(cb + (size_t)(((x + 1) * rp + (y + 1)) * rp + (z + 1)) * cl_stride);
  };
  const int groups = cl_out ? (CH == 4 ? 1 : (ch_n + 7) >> 3) : ch_n;
  // blockIdx.z owns a contiguous range of the cells (small batches: more CTAs than clouds x channel slices)
  const int cell_per = (ncell + gridDim.z - 1) / gridDim.z, cell_lo = blockIdx.z * cell_per, cell_hi = min(ncell, cell_lo + cell_per);
  for (int base = cell_lo + wid * 32; base < cell_hi; base += nthreads) {
    const int cell = base + lane;
    const bool valid = cell < cell_hi;
    const bool empty = valid && s_off[cell * cw + cw] == s_off[cell * cw];
    if (empty) {
      if (cl_out) {
        uint4* dst = cl_row(cell);
        for (int g = 0; g < groups; ++g) dst[g] = make_uint4(0u, 0u, 0u, 0u);
      } else if (quads) {
        const float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < ch_n; ++k) *reinterpret_cast<float4*>(ob + (size_t)k * r3 + 4 * cell) = zz;
      } else {
        for (int k = 0; k < ch_n; ++k) ob[(size_t)k * r3 + cell] = 0.f;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid && !empty);
    int pos = 0;
    if (lane == 0 && m) pos = atomicAdd(&s_ncells, __popc(m));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (valid && !empty) s_cells[pos + __popc(m & ((1u << lane) - 1u))] = (unsigned short)cell;
  }
  if (stage_feat) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int ncells = s_ncells, items = ncells * groups;
  const float* fsrc = stage_feat ? s_feat : fb;       // (generic loads: shared when staged)
  // divisor (float)(1.0 / (double)(float)cnt) of vox.cu:65 == the correctly rounded fp32 reciprocal for every count up to
  // 8192 (checked exhaustively), so no double-precision division here
  auto bucket_sum = [&](const float* f, int v) {      // one channel of voxel v: ascending point index
    const int o = s_off[v], cv = s_off[v + 1] - o;
    float acc = 0.f;
    if (cv) {
      const float div = __frcp_rn((float)cv);
      for (int j = 0; j < cv; ++j) acc = __fadd_rn(acc, __fmul_rn(f[s_sorted[o + j]], div));
    }
    return acc;
  };
  for (int it = tid; it < items; it += nthreads) {
    const int g = it / ncells, cell = s_cells[it - g * ncells];
    if (cl_out) {
      const int o = s_off[cell], cv = s_off[cell + 1] - o;
      const float div = __frcp_rn((float)cv);
      constexpr int CK = CH == 4 ? 4 : 8;
      float acc[CK];
#pragma unroll
      for (int k = 0; k < CK; ++k) acc[k] = 0.f;
      const float* f = fsrc + (size_t)(g * 8) * n;
      const int kn = min(CK, ch_n - g * 8);
      for (int j = 0; j < cv; ++j) {
        const int p = s_sorted[o + j];
#pragma unroll
        for (int k = 0; k < CK; ++k)
          if (k < kn) acc[k] = __fadd_rn(acc[k], __fmul_rn(f[(size_t)k * n + p], div));
      }
      uint4* dst = cl_row(cell) + g;
      if (CK == 4) *dst = make_uint4(gldm_pack_bf16(acc[0], acc[1]), gldm_pack_bf16(acc[2], acc[3]), 0u, 0u);
      else *dst = make_uint4(gldm_pack_bf16(acc[0], acc[1]), gldm_pack_bf16(acc[2], acc[3]),
                             gldm_pack_bf16(acc[4 % CK], acc[5 % CK]), gldm_pack_bf16(acc[6 % CK], acc[7 % CK]));
    } else if (quads) {
      const float* f = fsrc + (size_t)g * n;
      const float4 val = make_float4(bucket_sum(f, 4 * cell), bucket_sum(f, 4 * cell + 1), bucket_sum(f, 4 * cell + 2),
                                     bucket_sum(f, 4 * cell + 3));
      *reinterpret_cast<float4*>(ob + (size_t)g * r3 + 4 * cell) = val;
    } else {
      ob[(size_t)g * r3 + cell] = bucket_sum(fsrc + (size_t)g * n, cell);
    }
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
  // four points per thread up to 4096 points (256 threads for the 1024-point clouds of the model), more beyond
  int threads = min(1024, max(32, ceil_div(ceil_div(n, 4), 32) * 32));
  int ppt = ceil_div(n, threads);
  const bool use_smem = n <= 8192;                         // 12 n bytes: the winner's coordinates by broadcast LDS
  const size_t smem = use_smem ? sizeof(float) * 3 * (size_t)n : 0;
  if (smem > 48 * 1024) {
    static SmemOptIn a4, a8;
    if (int rc = opt_in_smem(a4, fps_kernel<4, true>, 96 * 1024, "fps_kernel")) return rc;
    if (int rc = opt_in_smem(a8, fps_kernel<8, true>, 96 * 1024, "fps_kernel")) return rc;
  }
#define FPS_LAUNCH(P)                                                                   \
  do {                                                                                  \
    if (use_smem) fps_kernel<P, true><<<b, threads, smem, s>>>(coords, n, m, indices);  \
    else fps_kernel<P, false><<<b, threads, 0, s>>>(coords, n, m, indices);             \
  } while (0)
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
  cudaStream_t s = (cudaStream_t)stream;
  const bool smem_ok = n <= kBqMaxSmemPoints;
  static SmemOptIn attr_t, attr_w;
  if ((long long)b * m >= 131072) {
    // throughput: thread = centre (fewest instructions per distance test)
    dim3 grid(ceil_div(m, kBqThreads), b);
    const int smem = 3 * ((n + 31) & ~31) * (int)sizeof(float);
    if (smem_ok && smem > 48 * 1024)
      if (int rc = opt_in_smem(attr_t, ball_query_kernel<true>, 3 * kBqMaxSmemPoints * (int)sizeof(float), "ball_query_kernel")) return rc;
    if (smem_ok) ball_query_kernel<true><<<grid, kBqThreads, smem, s>>>(centers, points, n, m, r2, u, neighbors);
    else ball_query_kernel<false><<<grid, kBqThreads, 0, s>>>(centers, points, n, m, r2, u, neighbors);
  } else {
    // latency: warp-cooperative scan, 32 centres per CTA
    dim3 grid(ceil_div(m, kBqCentersPerBlock), b);
    const int smem = 3 * ((n + 31) & ~31) * (int)sizeof(float);
    if (smem_ok && smem > 48 * 1024)
      if (int rc = opt_in_smem(attr_w, ball_query_warp_kernel<true>, 3 * kBqMaxSmemPoints * (int)sizeof(float), "ball_query_warp_kernel")) return rc;
    if (smem_ok) ball_query_warp_kernel<true><<<grid, 256, smem, s>>>(centers, points, n, m, r2, u, neighbors);
    else ball_query_warp_kernel<false><<<grid, 256, 0, s>>>(centers, points, n, m, r2, u, neighbors);
  }
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
  if (m <= kNnMaxSmemCenters)
    three_nn_kernel<true><<<grid, 128, 3 * ((m + 3) & ~3) * sizeof(float), (cudaStream_t)stream>>>(points, centers, feats, c, m, n,
                                                                                                 out, idx, w);
  else
    three_nn_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(points, centers, feats, c, m, n, out, idx, w);
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
                           float* out, int* ind, int* cnt, float* norm, int* vox, cudaStream_t s,
                           __nv_bfloat16* cl_out = nullptr, int cl_stride = 0) {
  GLDM_REQUIRE(b <= 0 || (feat && coords && (out || cl_out)), "voxelize: null pointer");
  GLDM_REQUIRE(b >= 0 && c > 0 && n > 0 && r > 0, "voxelize: bad sizes b=%d c=%d n=%d r=%d", b, c, n, r);
  GLDM_REQUIRE(n <= kVoxMaxPoints, "voxelize: n=%d > %d points per cloud not supported", n, kVoxMaxPoints);
  GLDM_REQUIRE(r <= 36, "voxelize: resolution %d > 36 not supported (shared-memory voxel table)", r);
  if (b == 0) return GLDM_OK;
  const size_t r3 = (size_t)r * r * r;
  size_t smem = ((2 * (((r3 + 3) & ~(size_t)1) + 4 * (((size_t)n + 1) & ~(size_t)1)) + 15) & ~(size_t)15) + 16;
  GLDM_REQUIRE(smem <= 200 * 1024, "voxelize: r=%d with n=%d points needs %zu bytes of shared memory (> 200 KB)", r, n, smem);
  int threads = min(512, ceil_div(n, 32) * 32);
  // channels per CTA: 4 for the coordinate-only first block, 16 for wide features in large batches, 8 otherwise
  static const int ch_env = getenv("GLDM_VOX_CH") ? atoi(getenv("GLDM_VOX_CH")) : 0;
  const int ch = c <= 4 ? 4 : ch_env ? ch_env : (c >= 32 && (long long)b * ceil_div(c, 16) >= 2 * kNumSMs) ? 16 : 8;
  const int slices = ceil_div(c, ch);
  // the slice's features are staged in shared memory when that keeps >= 3 CTAs per SM (16-byte copies: n % 4, alignment)
  const size_t feat_bytes = (size_t)min(c, ch) * n * 4;
  static const bool no_stage = getenv("GLDM_VOX_NOSTAGE") != nullptr;      // development knobs: GLDM_VOX_NOSTAGE, GLDM_VOX_CH
  const int stage = (n % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && smem + feat_bytes <= 80 * 1024 && !no_stage) ? 1 : 0;
  if (stage) smem += feat_bytes;
  // small batches of large grids: split the cells of a cloud over up to 8 CTAs (each repeats the binning)
  const int vsl = (r3 >= 4096) ? max(1, min(8, (2 * kNumSMs) / max(1, b * slices))) : 1;
  if (cl_out) {
    // a slice must be whole 16-byte chunks of the rows (or the single narrow slice of a <= 4-channel input)
    GLDM_REQUIRE((slices == 1 && c <= 4) || (c % ch == 0), "voxelize (channels-last): %d channels are not a multiple of %d", c, ch);
    GLDM_REQUIRE(cl_stride % 8 == 0 && cl_stride >= (slices == 1 && c <= 4 ? 8 : c), "voxelize (channels-last): bad row stride %d", cl_stride);
  }
#define VOX_LAUNCH(F, C_)                                                                                          \
  do {                                                                                                             \
    static SmemOptIn attr;                                                                                         \
    if (int rc = opt_in_smem(attr, voxelize_kernel<F, C_>, 200 * 1024, "voxelize_kernel")) return rc;              \
    voxelize_kernel<F, C_><<<dim3(b, slices, vsl), threads, smem, s>>>(feat, coords, c, n, r, out, ind, cnt, norm, vox, \
                                                                  cl_out, cl_stride, stage);                       \
  } while (0)
  if (fused) {
    if (ch == 4) VOX_LAUNCH(true, 4); else if (ch == 16) VOX_LAUNCH(true, 16); else VOX_LAUNCH(true, 8);
  } else {
    if (ch == 4) VOX_LAUNCH(false, 4); else if (ch == 16) VOX_LAUNCH(false, 16); else VOX_LAUNCH(false, 8);
  }
#undef VOX_LAUNCH
  return check_launch("voxelize_kernel");
}

extern "C" int gldm_avg_voxelize_forward(const float* features, const int* coords, int b, int c, int n, int r,
                                         float* out, int* ind, int* cnt, void* stream) {
  return launch_voxelize(false, features, coords, b, c, n, r, out, ind, cnt, nullptr, nullptr,
                         (cudaStream_t)stream);
}

extern "C" int gldm_voxelize_fused_cl(const float* features, const float* coords, int b, int c, int n, int r, void* x_cl,
                                      int stride, float* norm_coords, void* stream) {
  GLDM_REQUIRE(norm_coords && x_cl, "voxelize_fused_cl: null pointer");
  return launch_voxelize(true, features, coords, b, c, n, r, nullptr, nullptr, nullptr, norm_coords, nullptr,
                         (cudaStream_t)stream, reinterpret_cast<__nv_bfloat16*>(x_cl), stride);
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
