// Pre- and post-processing either side of the generation path:
//   * cloud normalisation of R/grasp_ldm/inference/inference_base.py:182-212 / R/tools/inference.py:570-591
//     (centre on the cloud mean, dataset shift / scale, per-object un-normalisation statistics),
//   * pose post-processing of R/tools/inference.py:627-656 and R/grasp_ldm/utils/rotations.py:171-252, 298-302
//     (un-normalise, MRP -> quaternion -> rotation matrix -> 4x4, sigmoid of the class logit).
// Both are HBM-bound streaming kernels; fp32 throughout, operator order of the reference.
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"

namespace gldm {

// One CTA per cloud: fp64 block reduction of the mean (the reference's torch.mean is an fp32 cascade sum; fp64 is
// within 1 ulp of it and order independent), then a second pass over the (L1/L2 resident) cloud.
__global__ void __launch_bounds__(256) normalize_cloud_kernel(const float* __restrict__ pc, const float* __restrict__ pc_shift,
                                                             const float* __restrict__ pc_scale,
                                                             const float* __restrict__ grasp_shift, int n,
                                                             float* __restrict__ pc_out, float* __restrict__ pc_mean_out,
                                                             float* __restrict__ grasp_mean_out) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* src = pc + (size_t)b * n * 3;
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = tid; i < n; i += blockDim.x) {
    s[0] += (double)src[i * 3 + 0];
    s[1] += (double)src[i * 3 + 1];
    s[2] += (double)src[i * 3 + 2];
  }
  __shared__ double red[3][8];
  __shared__ float mean[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
    if ((tid & 31) == 0) red[c][tid >> 5] = s[c];
  }
  __syncthreads();
  if (tid < 3) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[tid][w];
    const float m = (float)(t / (double)n);
    mean[tid] = m;
    if (pc_mean_out) pc_mean_out[b * 3 + tid] = __fadd_rn(pc_shift[tid], m);              // metas["pc_mean"]
  }
  if (grasp_mean_out && tid >= 32 && tid < 38) {                                            // metas["grasp_mean"]
    const int j = tid - 32;
    grasp_mean_out[b * 6 + j] = grasp_shift[j];
  }
  __syncthreads();
  if (grasp_mean_out && tid < 3) grasp_mean_out[b * 6 + tid] = __fadd_rn(grasp_shift[tid], mean[tid]);
  float* dst = pc_out + (size_t)b * n * 3;
  for (int i = tid; i < n * 3; i += blockDim.x) {
    const int c = i % 3;
    // (pc - pc_mean - shift) / scale, in the reference's two steps
    dst[i] = __fdiv_rn(__fsub_rn(__fsub_rn(src[i], mean[c]), pc_shift[c]), pc_scale[c]);
  }
}

__global__ void pose_post_kernel(const float* __restrict__ tmrp, const float* __restrict__ logit,
                                 const float* __restrict__ gmean, const float* __restrict__ gstd, int n, int gpo,
                                 int mean_stride, int std_stride, float* __restrict__ gt, float* __restrict__ Hm,
                                 float* __restrict__ conf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int obj = i / gpo;
  const float* gm = gmean + (size_t)obj * mean_stride;
  const float* gs = gstd + (size_t)obj * std_stride;
  float g[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    g[j] = __fadd_rn(__fmul_rn(tmrp[(size_t)i * 6 + j], gs[j]), gm[j]);   // tools/inference.py:91-94
    if (gt) gt[(size_t)i * 6 + j] = g[j];
  }
  if (Hm) {
    // rotations.py:218-252 (mrp_to_quat) and :171-215 (quat_to_rotmat), same operator order
    const float m0 = g[3], m1 = g[4], m2 = g[5];
    const float magsq = __fadd_rn(__fadd_rn(__fmul_rn(m0, m0), __fmul_rn(m1, m1)), __fmul_rn(m2, m2));
    const float den = __fadd_rn(1.0f, magsq);
    const float x = __fdiv_rn(__fmul_rn(2.0f, m0), den), y = __fdiv_rn(__fmul_rn(2.0f, m1), den),
                z = __fdiv_rn(__fmul_rn(2.0f, m2), den), w = __fdiv_rn(__fsub_rn(1.0f, magsq), den);
    const float x2 = __fmul_rn(x, x), y2 = __fmul_rn(y, y), z2 = __fmul_rn(z, z), w2 = __fmul_rn(w, w);
    const float xy = __fmul_rn(x, y), zw = __fmul_rn(z, w), xz = __fmul_rn(x, z), yw = __fmul_rn(y, w),
                yz = __fmul_rn(y, z), xw = __fmul_rn(x, w);
    float* o = Hm + (size_t)i * 16;
    o[0] = __fadd_rn(__fsub_rn(__fsub_rn(x2, y2), z2), w2);
    o[1] = __fmul_rn(2.0f, __fsub_rn(xy, zw));
    o[2] = __fmul_rn(2.0f, __fadd_rn(xz, yw));
    o[3] = g[0];
    o[4] = __fmul_rn(2.0f, __fadd_rn(xy, zw));
    o[5] = __fadd_rn(__fsub_rn(__fadd_rn(-x2, y2), z2), w2);
    o[6] = __fmul_rn(2.0f, __fsub_rn(yz, xw));
    o[7] = g[1];
    o[8] = __fmul_rn(2.0f, __fsub_rn(xz, yw));
    o[9] = __fmul_rn(2.0f, __fadd_rn(yz, xw));
    o[10] = __fadd_rn(__fadd_rn(__fsub_rn(-x2, y2), z2), w2);
    o[11] = g[2];
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
  }
  if (conf && logit) conf[i] = 1.0f / (1.0f + expf(-logit[i]));
}
}  // namespace gldm

using namespace gldm;

extern "C" int gldm_normalize_clouds(const float* pc, const float* pc_shift, const float* pc_scale,
                                     const float* grasp_shift, int b, int n, float* pc_out, float* pc_mean,
                                     float* grasp_mean, void* stream) {
  GLDM_REQUIRE(b <= 0 || (pc && pc_shift && pc_scale && pc_out), "normalize_clouds: null pointer");
  GLDM_REQUIRE(grasp_shift || !grasp_mean, "normalize_clouds: grasp_mean needs grasp_shift");
  GLDM_REQUIRE(b >= 0 && n > 0, "normalize_clouds: bad sizes");
  if (b == 0) return GLDM_OK;
  normalize_cloud_kernel<<<b, 256, 0, (cudaStream_t)stream>>>(pc, pc_shift, pc_scale, grasp_shift, n, pc_out, pc_mean,
                                                             grasp_mean);
  return check_launch("normalize_cloud_kernel");
}

extern "C" int gldm_pose_postprocess_rows(const float* tmrp, const float* logit, const float* grasp_mean,
                                          const float* grasp_std, int n, int grasps_per_obj, int mean_rows,
                                          int std_rows, float* grasp_tmrp, float* H, float* conf, void* stream) {
  GLDM_REQUIRE(n <= 0 || (tmrp && grasp_mean && grasp_std), "pose_postprocess: null pointer");
  GLDM_REQUIRE(n >= 0 && grasps_per_obj > 0, "pose_postprocess: bad n");
  GLDM_REQUIRE(n % grasps_per_obj == 0, "pose_postprocess: n is not a multiple of grasps_per_obj");
  const int n_obj = n / grasps_per_obj;
  GLDM_REQUIRE(mean_rows == 1 || mean_rows == n_obj, "pose_postprocess: grasp_mean needs 1 row or one per object");
  GLDM_REQUIRE(std_rows == 1 || std_rows == n_obj, "pose_postprocess: grasp_std needs 1 row or one per object");
  if (n == 0) return GLDM_OK;
  pose_post_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(
      tmrp, logit, grasp_mean, grasp_std, n, grasps_per_obj, mean_rows == 1 ? 0 : 6, std_rows == 1 ? 0 : 6, grasp_tmrp,
      H, conf);
  return check_launch("pose_post_kernel");
}

extern "C" int gldm_pose_postprocess(const float* tmrp, const float* logit, const float* grasp_mean,
                                     const float* grasp_std, int n, float* grasp_tmrp, float* H, float* conf,
                                     void* stream) {
  GLDM_REQUIRE(n >= 0, "pose_postprocess: bad n");
  return gldm_pose_postprocess_rows(tmrp, logit, grasp_mean, grasp_std, n, n > 0 ? n : 1, 1, 1, grasp_tmrp, H, conf,
                                    stream);
}
