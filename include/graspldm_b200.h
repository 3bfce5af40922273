/*
 * graspldm_b200 - C ABI of the B200-native (sm_100a) GraspLDM generation path.
 *
 * Every entry point takes plain device pointers, sizes and a cudaStream_t (as void*), launches
 * asynchronously on that stream and returns 0 on success or a negative GLDM_E* code; the text of
 * the last error of the calling thread is available from gldm_last_error().  No torch types.
 *
 * Section A replaces, one for one, the pybind11 surface of the reference's operator extension
 * `_pvcnn_backend` (R = /root/reference/grasp_ldm/models/modules/ext/pvcnn/modules/functional/src):
 * the reference-side binding a maintainer would add is shown in INTEGRATION.md.
 * Section B are the fused entry points behind the model-class API (boundary A of SURVEY.md 8b).
 */
#ifndef GRASPLDM_B200_H_
#define GRASPLDM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLDM_OK 0
#define GLDM_EINVAL (-1)   /* bad argument (null pointer, non-positive size, unsupported shape) */
#define GLDM_ECUDA (-2)    /* CUDA runtime error at launch */
#define GLDM_ENOSUP (-3)   /* configuration not supported by this build */

const char* gldm_last_error(void);
int gldm_version(void);
/* number of kernels this library has launched since load (bench.py reports the delta) */
unsigned long long gldm_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Section A - operator FFI (replaces R/bindings.cpp:10-37)
 * ------------------------------------------------------------------------------------------ */

/* avg_voxelize_forward            R/voxelization/vox.cpp:17-43, vox.cu:18-72
 * features f32[b,c,n], coords i32[b,3,n] (voxel coords in [0,r)) ->
 * out f32[b,c,r^3], ind i32[b,n], cnt i32[b,r^3].  All outputs are fully written (no pre-zeroing
 * needed).  Sums run in ascending point order, so the result is run-to-run deterministic
 * (the reference uses float atomics). */
int gldm_avg_voxelize_forward(const float* features, const int* coords, int b, int c, int n, int r,
                              float* out, int* ind, int* cnt, void* stream);
/* avg_voxelize_backward           R/voxelization/vox.cpp:54-77, vox.cu:86-110
 * grad_y f32[b,c,r^3], ind i32[b,n], cnt i32[b,r^3] -> grad_x f32[b,c,n] */
int gldm_avg_voxelize_backward(const float* grad_y, const int* ind, const int* cnt, int b, int c, int n,
                               int r3, float* grad_x, void* stream);

/* trilinear_devoxelize_forward    R/interpolate/trilinear_devox.cpp:18-55, trilinear_devox.cu:21-105
 * coords f32[b,3,n] (voxel units), features f32[b,c,r^3] -> outs f32[b,c,n];
 * inds i32[b,8,n] / wgts f32[b,8,n] are written only when is_training != 0 (may be NULL otherwise). */
int gldm_trilinear_devoxelize_forward(const float* coords, const float* features, int b, int c, int n, int r,
                                      int is_training, float* outs, int* inds, float* wgts, void* stream);
/* trilinear_devoxelize_backward   R/interpolate/trilinear_devox.cpp:67-92
 * grad_y f32[b,c,n], inds i32[b,8,n], wgts f32[b,8,n] -> grad_x f32[b,c,r3] (zeroed by the call) */
int gldm_trilinear_devoxelize_backward(const float* grad_y, const int* inds, const float* wgts, int b, int c,
                                       int n, int r3, float* grad_x, void* stream);

/* furthest_point_sampling         R/sampling/sampling.cpp:43-58, sampling.cu:86-167
 * coords f32[b,3,n] -> indices i32[b,m]; bit-exact index contract (start 0, FMA chain
 * fma(dz,dz,fma(dx,dx,dy*dy)), ties -> smallest (k mod 512) then smallest k). */
int gldm_furthest_point_sampling(const float* coords, int b, int n, int m, int* indices, void* stream);

/* gather_features_forward/backward R/sampling/sampling.cpp:6-41, sampling.cu:17-73
 * features f32[b,c,n], indices i32[b,m] -> out f32[b,c,m];  backward: grad_y f32[b,c,m] -> grad_x f32[b,c,n] */
int gldm_gather_features_forward(const float* features, const int* indices, int b, int c, int n, int m,
                                 float* out, void* stream);
int gldm_gather_features_backward(const float* grad_y, const int* indices, int b, int c, int n, int m,
                                  float* grad_x, void* stream);

/* ball_query                      R/ball_query/ball_query.cpp:6-30, ball_query.cu:19-50
 * centers f32[b,3,m], points f32[b,3,n], radius, u -> neighbors i32[b,m,u] (bit-exact) */
int gldm_ball_query(const float* centers, const float* points, int b, int n, int m, float radius, int u,
                    int* neighbors, void* stream);

/* grouping_forward/backward       R/grouping/grouping.cpp:6-45, grouping.cu:18-72
 * features f32[b,c,n], indices i32[b,m,u] -> out f32[b,c,m,u];  backward -> grad_x f32[b,c,n] */
int gldm_grouping_forward(const float* features, const int* indices, int b, int c, int n, int m, int u,
                          float* out, void* stream);
int gldm_grouping_backward(const float* grad_y, const int* indices, int b, int c, int n, int m, int u,
                           float* grad_x, void* stream);

/* three_nearest_neighbors_interpolate_forward/backward
 *                                 R/interpolate/neighbor_interpolate.cpp:6-64, neighbor_interpolate.cu:20-160
 * points f32[b,3,n], centers f32[b,3,m], feats f32[b,c,m] -> out f32[b,c,n], idx i32[b,3,n], w f32[b,3,n] */
int gldm_three_nn_interpolate_forward(const float* points, const float* centers, const float* feats, int b,
                                      int c, int m, int n, float* out, int* idx, float* w, void* stream);
int gldm_three_nn_interpolate_backward(const float* grad_y, const int* idx, const float* w, int b, int c,
                                       int n, int m, float* grad_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * Section B - fused generation path (boundary A: model classes call these)
 * ------------------------------------------------------------------------------------------ */

/* Voxelization.forward + avg_voxelize in one launch
 *   (R/../pvcnn/modules/voxelization.py:16-35, normalize=False branch, + vox.cu:18-72)
 * features f32[b,c,n], coords f32[b,3,n] -> grid f32[b,c,r^3], norm_coords f32[b,3,n]
 * (the clamped float coordinates devoxelize consumes), vox i32[b,3,n] (optional, may be NULL). */
int gldm_voxelize_fused(const float* features, const float* coords, int b, int c, int n, int r, float* grid,
                        float* norm_coords, int* vox, void* stream);
/* The same, writing the averaged features straight into the zero-padded channels-last bf16 grid the Conv3d kernels read
 * (x_cl [b * (r+2)^3][stride] bf16, halo rows and padding channels are left untouched: allocate the grid zeroed).
 * Requires c <= 4 or c a multiple of 8.  Replaces gldm_voxelize_fused + gldm_cl_pad of the tensor-core path. */
int gldm_voxelize_fused_cl(const float* features, const float* coords, int b, int c, int n, int r, void* x_cl, int stride,
                           float* norm_coords, void* stream);

/* ---- dense fp32 building blocks of the encoder (SIMT, strict-fp32 parity mode) ---- */
/* y[b,co,n] = act(scale[co] * (sum_ci W[co,ci] x[b,ci,n]) + shift[co]) (+ add[b,co,n] when add != NULL)
 * act: 0 none, 1 relu.  Conv1d k=1 + folded bias/BatchNorm(eval) + ReLU == SharedMLP
 * (R/../pvcnn/modules/shared_mlp.py:18-28), conv_downscale, out_layer.0 (R/models/modules/pc_encoders.py:60-75) */
int gldm_pointwise_conv_f32(const float* x, const float* w, const float* scale, const float* shift,
                            const float* add, int b, int ci, int co, int n, int act, float* y, void* stream);
/* Conv3d k=3 pad=1 + bias on [b,ci,r,r,r] -> [b,co,r,r,r]   (R/../pvcnn/modules/pvconv.py:48-67) */
int gldm_conv3d_k3_f32(const float* x, const float* w, const float* bias, int b, int ci, int co, int r,
                       float* y, void* stream);
/* GroupNorm(groups, eps) + Swish in place on [b,c,s]; when se_mean != NULL also writes the per-(b,c) mean of
 * the activated output (SE3d squeeze, R/../pvcnn/modules/se.py:22-25). */
int gldm_groupnorm_swish_f32(float* x, const float* gamma, const float* beta, int b, int c, int s, int groups,
                             float eps, float* se_mean, void* stream);
/* SE3d excite: gate[b,c] = sigmoid(W2 swish(W1 mean[b,:])) ; x[b,c,:] *= gate[b,c]  (se.py:12-25) */
int gldm_se_gate_f32(const float* mean, const float* w1, const float* w2, int b, int c, int cr, float* gate,
                     void* stream);
/* trilinear devoxelize of (grid * gate) fused with the residual add of the point branch:
 * out[b,c,n] = devox(grid[b,c,:] * gate[b,c], coords)[n] + point[b,c,n]    (pvconv.py:79-83) */
int gldm_devox_gate_add_f32(const float* coords, const float* grid, const float* gate, const float* point,
                            int b, int c, int n, int r, float* out, void* stream);
/* Linear over the point axis: y[b,c,f] = sum_n x[b,c,n] W[f,n] + bias[f]   (pc_encoders.py:76-79) */
int gldm_linear_lastdim_f32(const float* x, const float* w, const float* bias, int rows, int n, int f, float* y,
                            void* stream);

/* ---- ResNet1D family (denoiser / decoder), SURVEY.md 8a rows a17-a19 ---- */
typedef struct GldmResNetCfg {
  int L;             /* sequence length: latent dim D (denoiser) / feature_resolution (decoder) */
  int n_stages;      /* 4 */
  int ch[6];         /* ch[0] = dim, ch[1..n_stages] = block_channels */
  int emb_dim;       /* 4 * dim */
  int cond_ch;       /* conditioning channels (3) */
  int cond_dim;      /* conditioning features (64 fpc / 256 ppc) */
  int groups;        /* GroupNorm groups (4) */
  int time_cond;     /* 1: TimeConditionedResNet1D, 0: ResNet1D */
  int fourier_half;  /* 8 (learned_sinusoidal_dim / 2) */
  int heads;         /* 4 */
  int dim_head;      /* 32 */
} GldmResNetCfg;

/* Number of floats in the canonical raw parameter blob (state_dict tensors, fixed key order documented in
 * graspldm_b200/engine.py:resnet_param_keys) and in the prepared blob the kernels read. */
long long gldm_resnet_raw_floats(const GldmResNetCfg* cfg);
long long gldm_resnet_prepared_floats(const GldmResNetCfg* cfg);
/* One-time weight preparation on the device: weight standardisation (R/models/modules/resnets.py:85-91),
 * [k][co] transposition.  raw and prepared are device pointers. */
int gldm_resnet_prepare(const GldmResNetCfg* cfg, const float* raw, float* prepared, void* stream);

/* Sampler descriptor: per executed step i (loop order, t descending) the integer timestep and the
 * scheduler coefficients, computed on the host in diffusers' operator order (oracle/schedulers.py
 * restates them): coef[i] = {sqrt(1-abar_t), sqrt(abar_t), c_x0, c_xt_or_eps, sigma, 0,0,0}. */
#define GLDM_SCHED_DDPM 0
#define GLDM_SCHED_DDIM 1
#define GLDM_SCHED_EDM 2   /* evaluation program of the elucidated samplers, see GldmSamplerArgs */
/* Whole T-step reverse diffusion in ONE launch  (GaussianDiffusion1D.sample,
 * R/models/diffusion/gaussian_diffusion.py:232-277 + TimeConditionedResNet1D.forward resnets.py:558-616):
 *   x_T f32[n,D] initial latents, z_obj f32[n_obj,cond_ch,cond_dim] per-object conditioning,
 *   sample s uses object s / grasps_per_obj.
 *   noise f32[n_steps,n,D] pre-drawn per-step noise (parity mode) or NULL -> in-kernel Philox4x32-10
 *   + Box-Muller keyed by (seed, sample, step).
 *   x_out f32[n,D]; x_all f32[n_steps+1,n,D] optional trajectory (may be NULL). */
int gldm_sampler_run_f32(const GldmResNetCfg* cfg, const float* prepared, const float* x_T, const float* z_obj,
                         int n, int grasps_per_obj, int n_steps, const int* timesteps_host,
                         const float* coef_host, int sched_kind, int clip_sample, const float* noise,
                         unsigned long long seed, float* x_out, float* x_all, void* stream);
/* One denoiser evaluation (eps prediction) with per-sample timesteps: TimeConditionedResNet1D.forward.
 * x f32[n,D], t i32[n], z_cond f32[n,cond_ch,cond_dim] (per sample) -> eps f32[n,D] */
int gldm_denoiser_forward_f32(const GldmResNetCfg* cfg, const float* prepared, const float* x, const int* t,
                              const float* z_cond, int n, float* eps, void* stream);
int gldm_denoiser_forward_f32_ftime(const GldmResNetCfg* cfg, const float* prepared, const float* x, const float* t,
                                    const float* z_cond, int n, float* eps, void* stream);
/* ConditionalGraspPoseDecoder.forward (R/models/grasp_vae.py:401-436): in_layer -> ResNet1D -> heads.
 * head weights: in_w f32[L,D] in_b[L]  tmrp_w[6,L] tmrp_b[6]  cls_w[1,L] cls_b[1] packed in `head` in
 * that order.  z_h f32[n,D], z_obj f32[n_obj,cond_ch,cond_dim] -> tmrp f32[n,6], logit f32[n,1] */
int gldm_decoder_forward_f32(const GldmResNetCfg* cfg, const float* prepared, const float* head, int D,
                             const float* z_h, const float* z_obj, int n, int grasps_per_obj, float* tmrp,
                             float* logit, void* stream);

/* ---- tensor-core (tcgen05 / TMEM) sampler: bf16 operands, fp32 accumulation ----
 * Same contract as gldm_sampler_run_f32 / gldm_denoiser_forward_f32, but the GEMMs run on the 5th-generation
 * tensor cores.  `raw` is the canonical fp32 parameter blob (per-channel parameters are read from it), `pack`
 * the bf16 UMMA weight images produced once by gldm_sampler_tc_prepare (gldm_sampler_tc_pack_bytes bytes,
 * 1024-byte aligned).  Supported: the fpc latent denoiser family (L = 4, emb 16, time conditioned) and the grasp
 * decoder trunk (L = 16, emb 64), 4 stages of width <= 128, final width <= 256; anything else returns GLDM_ENOSUP. */
/* sample sets per sampler CTA: 0 = automatic (two sets of 16 samples once the batch exceeds one wave of 16-sample
 * CTAs: the tensor-core phase of one set then overlaps the epilogue of the other), 1 or 2 to force */
int gldm_sampler_tc_set_sets(int sets);
/* Kernel choice for the fpc latent denoiser: 1 = row-major kernel (activations as the M = 128 operand, 32 samples per
 * CTA: best throughput per SM), 0 = channel-major kernel (weights as the M operand, 16 or 32 samples per CTA; it also
 * serves the ppc denoiser and the decoder), -1 (default) = by batch size (row-major above one wave of 16-sample CTAs) */
int gldm_sampler_tc_set_rows(int on);
long long gldm_sampler_tc_pack_bytes(const GldmResNetCfg* cfg);
/* development aid: when dev_buf != NULL (>= 512 int64 on the device) CTA 0 stamps clock64() around every
 * accumulator wait of its second denoising step; NULL (default) disables it */
int gldm_sampler_tc_set_profile(long long* dev_buf);
int gldm_sampler_tc_prepare(const GldmResNetCfg* cfg, const float* raw, void* pack, void* stream);
int gldm_sampler_run_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x_T,
                        const float* z_obj, int n, int grasps_per_obj, int n_steps, const int* timesteps_host,
                        const float* coef_host, int sched_kind, int clip_sample, const float* noise,
                        unsigned long long seed, float* x_out, float* x_all, void* stream);
/* Same sampler with device-resident tables (no per-call allocation or host copy): coef_dev f32[n_steps,8],
 * te_dev f32[n_steps,emb] produced once per schedule by gldm_time_embed_table(timesteps_dev i32[n_steps]). */
int gldm_time_embed_table(const GldmResNetCfg* cfg, const float* raw, const int* timesteps_dev, int count,
                          float* te_dev, void* stream);
int gldm_sampler_run_tc_dev(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x_T,
                            const float* z_obj, int n, int grasps_per_obj, int n_steps, const float* coef_dev,
                            const float* te_dev, int sched_kind, int clip_sample, const float* noise,
                            unsigned long long seed, float* x_out, float* x_all, void* stream);
int gldm_denoiser_forward_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x,
                             const int* t, const float* z_cond, int n, float* eps, void* stream);
/* the same with continuous per-sample times (elucidated sampler: time = c_noise(sigma) = log(sigma) / 4,
 * R/grasp_ldm/models/diffusion/elucidated_diffusion.py:120-147) */
int gldm_denoiser_forward_tc_ftime(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x,
                                   const float* t, const float* z_cond, int n, float* eps, void* stream);
/* ConditionalGraspPoseDecoder.forward on the tensor cores (same contract as gldm_decoder_forward_f32); cfg is the
 * decoder trunk (L = 16, emb 64, not time conditioned), pack from gldm_sampler_tc_prepare with that cfg. */
int gldm_decoder_forward_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* head, int D,
                            const float* z_h, const float* z_obj, int n, int grasps_per_obj, float* tmrp,
                            float* logit, void* stream);

/* ---- tensor-core GEMM for the encoder's point-wise layers (bf16 operands, fp32 accumulation) ----
 * Operands are "UMMA images": [row tile of 128][K block of 64][128 rows x 128 bytes, SWIZZLE_128B] bf16.
 * gldm_gemm_tc_image_bytes(rows, k): size of such an image.  gldm_gemm_tc_pack_weight: fp32 W[n_out,k] -> image.
 * gldm_gemm_tc_to_image: fp32 activations [b,c,n] (channel-major) -> image with rows m = b*n + point.
 * gldm_gemm_tc_run: out_img[m, n] = act(scale[n] * sum_k A[m,k] W[n,k] + shift[n]) written as the next layer's image
 *   (SharedMLP / conv_downscale, R/../pvcnn/modules/shared_mlp.py:18-28, R/models/modules/pc_encoders.py:60-67).
 * gldm_gemm_tc_image_small_co: y f32[b,co,n] = W[co,k] * image + bias, co <= 4 (out_layer.0, pc_encoders.py:68-75). */
long long gldm_gemm_tc_image_bytes(long long rows, int k);
int gldm_gemm_tc_pack_weight(const float* w, int n_out, int k, void* img, void* stream);
int gldm_gemm_tc_to_image(const float* x, int b, int c, int n, void* img, void* stream);
int gldm_gemm_tc_run(const void* a_img, const void* w_img, const float* scale, const float* shift, long long rows,
                     int k, int n_out, int relu, void* out_img, void* stream);
int gldm_gemm_tc_image_small_co(const void* img, const float* w, const float* bias, long long rows, int k, int co,
                                int n, float* y, void* stream);
/* gldm_gemm_tc_run with a fused projection instead of an output image:
 *   y f32[b, co, n] = sum_j proj_w[j][c] * act(scale[j] * sum_k A[m,k] W[j,k] + shift[j]) + proj_bias[c],   co <= 4.
 * proj_w f32[n_out][4] (output channel fastest, unused entries zero, 16-byte aligned); partials: 16 * rows * n_out / 128
 * bytes of scratch (per-tile partial sums, added in tile order: deterministic).  Used for the last SharedMLP of the
 * encoder followed by conv_downscale and out_layer.0, which are two affine maps in a row without a non-linearity
 * (R/models/modules/pc_encoders.py:60-75, 104-112 with use_global_attention=False) and are composed once into
 * proj_w = (W_out W_down)^T, proj_bias = W_out b_down + b_out: the [rows, 1536] and [rows, 768] activations never exist. */
int gldm_gemm_tc_run_proj(const void* a_img, const void* w_img, const float* scale, const float* shift, long long rows,
                          int k, int n_out, int relu, const float* proj_w, const float* proj_bias, int co, int n,
                          void* partials, float* y, void* stream);

/* ---- tensor-core Conv3d k3 p1 (bf16 operands, fp32 accumulation), implicit GEMM over a zero-padded channels-last
 * grid loaded with TMA tensor copies (R/../pvcnn/modules/pvconv.py:48-67).  16 <= ci, co <= 128.
 * w_img: gldm_conv3d_tc_weight_bytes(ci) bytes from gldm_conv3d_tc_pack_weight(w f32[co,ci,3,3,3]);
 * scratch: gldm_conv3d_tc_grid_bytes(b, ci, r) bytes, 256-byte aligned.  x f32[b,ci,r^3] -> y f32[b,co,r^3] (+ bias). */
long long gldm_conv3d_tc_weight_bytes(int ci);
long long gldm_conv3d_tc_grid_bytes(int b, int ci, int r);
int gldm_conv3d_tc_pack_weight(const float* w, int co, int ci, void* img, void* stream);
int gldm_conv3d_k3_tc(const float* x, const void* w_img, const float* bias, int b, int ci, int co, int r, void* scratch,
                      float* y, void* stream);

/* ---- fused voxel branch of PVConv (R/../pvcnn/modules/pvconv.py:48-67, 84-96; se.py:10-21): the voxel grids stay in HBM
 * as zero-padded channels-last rows [b * (r+2)^3][stride] (bf16 with stride = channels rounded up to 64 where the next
 * Conv3d reads them through TMA, fp32 with stride = channels for the devoxelize input).  Halo rows of an output grid
 * are never written and must be zero on entry.
 *   gldm_cl_pad:        x f32[b,c,r^3] -> bf16 padded grid (stride = c rounded up to 64), halo rows written as zero
 *   gldm_conv3d_tc_cl:  padded bf16 grid -> padded grid y_cl (+ bias) and GroupNorm(8) statistics
 *                       stats f64[b][8][2] = (sum, sum of squares) per group
 *   gldm_gn_swish_cl:   GroupNorm(8) + Swish in place from those statistics; se_sum f64[b][c] = channel sums (or NULL)
 * All sums are bit-reproducible: blocks write partials into the workspace `ws` (gldm_voxel_ws_bytes(b, c, r) bytes,
 * 256-byte aligned, contents irrelevant on entry) and a second small kernel adds them in a fixed order - no atomics.
 *   gldm_se_gate_sum:   gate f32[b][c] = sigmoid(W2 swish(W1 (se_sum / count)))
 *   gldm_devox_cl:      out f32[b,c,n] = trilinear(grid * gate)(coords f32[b,3,n] in voxel units) + point f32[b,c,n] */
int gldm_cl_pad(const float* x, int b, int c, int r, void* out_cl, void* stream);
/* strict-fp32 SIMT Conv3d k3 p1 for few input channels (the 3 -> 48 first layer) writing the padded bf16 grid
 * y_cl [b * (r+2)^3][y_stride] (+ bias) and the GroupNorm(8) statistics of its 48 output channels; w f32[ci][27][48] */
int gldm_conv3d_k3_f32_cl(const float* x, const float* w, const float* bias, int b, int ci, int r, void* y_cl,
                          int y_stride, double* stats, void* ws, void* stream);
long long gldm_voxel_ws_bytes(int b, int c, int r);
int gldm_conv3d_tc_cl(const void* x_cl, const void* w_img, const float* bias, int b, int ci, int co, int r, void* y_cl,
                      int out_fp32, int out_stride, double* stats, void* ws, void* stream);
int gldm_gn_swish_cl(void* y_cl, int is_fp32, int stride, const double* stats, const float* gamma, const float* beta, int b,
                     int c, int r, float eps, double* se_sum, void* ws, void* stream);
int gldm_block_partials_to_stats(const double* part, int b, int nblk, double* stats, void* stream);
/* narrow-input Conv3d k3 p1 (ci <= 16: the 3 -> 48 first layer) on the tensor cores: x f32[b,ci,r^3] -> padded bf16 grid
 * y_cl [b * (r+2)^3][out_stride] (+ bias) and its GroupNorm(8) statistics.  w_img: gldm_conv3d_tc16_weight_bytes() bytes
 * from gldm_conv3d_tc16_pack_weight(w f32[co,ci,3,3,3]); scratch: b * (r+2)^3 * 32 bytes, 256-byte aligned */
long long gldm_conv3d_tc16_weight_bytes(void);
int gldm_conv3d_tc16_pack_weight(const float* w, int co, int ci, void* img, void* stream);
int gldm_conv3d_tc16_cl(const float* x, const void* w_img, const float* bias, int b, int ci, int co, int r, void* scratch,
                        void* y_cl, int out_stride, double* stats, void* ws, void* stream);
int gldm_se_gate_sum(const double* sum, int count, const float* w1, const float* w2, int b, int c, int cr, float* gate,
                     void* stream);
int gldm_devox_cl(const float* coords, const void* grid_cl, int is_fp32, int stride, const float* gate, const float* point,
                  int b, int c, int n, int r, float* out, void* stream);

/* Pose post-processing (R/../tools/inference.py:627-656, R/utils/rotations.py:298-302):
 * tmrp f32[n,6], logit f32[n], grasp_mean/std f32[6] -> grasp_tmrp f32[n,6], H f32[n,4,4], conf f32[n] */
int gldm_pose_postprocess(const float* tmrp, const float* logit, const float* grasp_mean, const float* grasp_std,
                          int n, float* grasp_tmrp, float* H, float* conf, void* stream);
/* Same with per-object statistics, as the reference's datasets and normalize_input produce them
 * (R/grasp_ldm/dataset/acronym/acronym_pointclouds.py:232-243, R/tools/inference.py:581-589): grasp i belongs to object
 * i / grasps_per_obj; grasp_mean f32[mean_rows,6], grasp_std f32[std_rows,6], rows = 1 (shared) or n / grasps_per_obj. */
int gldm_pose_postprocess_rows(const float* tmrp, const float* logit, const float* grasp_mean, const float* grasp_std,
                               int n, int grasps_per_obj, int mean_rows, int std_rows, float* grasp_tmrp, float* H,
                               float* conf, void* stream);

/* Raw-cloud normalisation (R/grasp_ldm/inference/inference_base.py:182-212, R/tools/inference.py:570-591):
 * pc f32[b,n,3]; pc_shift, pc_scale f32[3]; grasp_shift f32[6] (may be NULL when grasp_mean is NULL).
 * pc_out[b,n,3] = (pc - mean_n(pc) - pc_shift) / pc_scale; pc_mean[b,3] = pc_shift + mean_n(pc) (metas["pc_mean"]);
 * grasp_mean[b,6] = grasp_shift with the cloud mean added to the translation part (metas["grasp_mean"]).
 * pc_out must not alias pc (the reference centres its argument in place; this entry leaves it untouched). */
int gldm_normalize_clouds(const float* pc, const float* pc_shift, const float* pc_scale, const float* grasp_shift, int b,
                          int n, float* pc_out, float* pc_mean, float* grasp_mean, void* stream);

/* ------------------------------------------------------------------------------------------
 * Extended sampler / denoiser entry points: class conditioning and the elucidated samplers
 * ------------------------------------------------------------------------------------------ */
/* All pointers are DEVICE pointers.  sched_kind GLDM_SCHED_EDM runs an evaluation program: one 16-float row per network
 * evaluation - [0] c_in [1] c_skip [2] c_out [3] kind [4..7] k0..k3 [8] noise slot (-1: none) [9] x_all slot (-1: none) -
 *   kind 0  stochastic Heun, first evaluation   R/grasp_ldm/models/diffusion/elucidated_diffusion.py:214-237
 *   kind 1  second-order correction             :239-256
 *   kind 2  DPM-Solver++(2M) step               :282-313
 * (formulas next to eval_update in csrc/resnet_layout.cuh).  x_all f32[slots][n][L]: slot 0 receives x_init. */
typedef struct {
  const float* x_init;       /* f32[n][L] */
  const float* z_obj;        /* f32[n_obj][cond_ch][cond_dim], sample i uses row i / grasps_per_obj */
  int n, grasps_per_obj;
  int sched_kind;            /* GLDM_SCHED_DDPM / DDIM / EDM */
  int n_steps;               /* network evaluations */
  const float* coef;         /* f32[n_steps][8] (DDPM / DDIM, as gldm_sampler_run_*) or f32[n_steps][16] (EDM) */
  const int* timesteps;      /* i32[n_steps]  (DDPM / DDIM; fp32 kernel only) */
  const float* times;        /* f32[n_steps] continuous time c_noise(sigma) (EDM; fp32 kernel only) */
  const float* te;           /* f32[n_steps][emb] time-embedding table (tensor-core kernels only, gldm_time_embed_table_f) */
  int clip_sample;
  const float* noise;        /* f32[noise slots][n][L] pre-drawn N(0,1) or NULL (in-kernel Philox keyed by seed) */
  unsigned long long seed;
  const float* cls_emb;      /* f32[n_obj][emb] added to the time embedding (class_conditioned_resnet.py:96-98) or NULL */
  float* x_out;              /* f32[n][L] */
  float* x_all;              /* NULL or f32[slots][n][L] */
} GldmSamplerArgs;
int gldm_sampler_run_ex_f32(const GldmResNetCfg* cfg, const float* prepared, const GldmSamplerArgs* args, void* stream);
int gldm_sampler_run_ex_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const GldmSamplerArgs* args,
                           void* stream);
/* single evaluation with integer (t) or continuous (tf) time and an optional class embedding f32[n][emb] */
int gldm_denoiser_forward_ex_f32(const GldmResNetCfg* cfg, const float* prepared, const float* x, const int* t,
                                 const float* tf, const float* z_cond, const float* cls_emb, int n, float* eps, void* stream);
int gldm_denoiser_forward_ex_tc(const GldmResNetCfg* cfg, const float* raw, const void* pack, const float* x, const int* t,
                                const float* tf, const float* z_cond, const float* cls_emb, int n, float* eps, void* stream);
/* time-embedding table from continuous times f32[count] -> f32[count][emb] */
int gldm_time_embed_table_f(const GldmResNetCfg* cfg, const float* raw, const float* times_dev, int count, float* te_dev,
                            void* stream);
/* cls_embed: out f32[n][emb] = SiLU(w * cls + b)   (class_conditioned_resnet.py:43-46) */
int gldm_class_embed(const float* w, const float* b, const float* cls, int n, int emb, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Set abstraction (PointNet++ / PVCNN2 side of the operator family), fused
 * ------------------------------------------------------------------------------------------ */
/* PointNetSAModule.forward for one radius, after the centre selection   R/../pvcnn/modules/pointnet.py:100-111,
 * ball_query.py:16-34, shared_mlp.py:18-28: ball-query neighbours idx i32[b,m,u] -> gather (coords - centre | features)
 * -> n_layers x (1x1 conv, folded BatchNorm scale / shift, ReLU) -> max over the u neighbours -> out f32[b, widths[last], m].
 * The grouped tensor [b, c+3, m, u] never exists in HBM.  wt[l] = layer weight TRANSPOSED to [c_in][c_out] (device
 * pointers in HOST arrays of n_layers entries); every width <= 512, c + 3 <= 512. */
int gldm_sa_mlp_max_f32(const float* coords, const float* centers, const float* feats, const int* idx, int b, int c, int n,
                        int m, int u, int include_coords, int n_layers, const int* widths, const float* const* wt,
                        const float* const* scale, const float* const* shift, float* out, void* stream);
/* SE excite with ReLU (se.py:12-25, use_relu=True): gate f32[b,c] = sigmoid(W2 relu(W1 mean)) */
int gldm_se_gate_relu_f32(const float* mean, const float* w1, const float* w2, int b, int c, int cr, float* gate,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRASPLDM_B200_H_ */
