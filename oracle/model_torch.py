"""CPU oracle for the dense part of the generation path.  TEST INFRASTRUCTURE ONLY.

A plain PyTorch (fp32, CPU-runnable) functional restatement of the reference's
forward passes, operating directly on a reference-keyed `state_dict` instead of
nn.Module trees.  R = /root/reference/grasp_ldm.

It is pinned against the reference itself: tests/golden/make_golden.py imports the
unmodified reference modules in the build container (they cannot travel to the GPU
box), runs them on seeded random-init weights of `fpc_1a_latentc3_z4_pc64`, and
stores the outputs under tests/golden/; tests/test_oracle_golden.py replays the same
seeds through the functions below.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import math

import torch
import torch.nn.functional as F

from . import ops_np
from .schedulers import SchedulerOracle, timestep_list


# ----------------------------------------------------------------------------- resnets.py
def _ws_conv1d(x, w, b, padding):
    """R/models/modules/resnets.py:79-101 (WeightStandardizedConv2d, eps 1e-5 for fp32)."""
    mean = w.mean(dim=(1, 2), keepdim=True)
    var = w.var(dim=(1, 2), unbiased=False, keepdim=True)
    return F.conv1d(x, (w - mean) * (var + 1e-5).rsqrt(), b, padding=padding)


def _chan_layernorm(x, g):
    """resnets.py:104-113: normalise over the channel axis of [B,C,L]."""
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


def _block(sd, p, x, groups, scale_shift=None):
    """resnets.py:127-177."""
    x = _ws_conv1d(x, sd[p + "proj.weight"], sd[p + "proj.bias"], 1)
    x = F.group_norm(x, groups, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        if scale.shape[-1] == 1:
            x = x * (scale + 1) + shift
        else:  # multi-channel FiLM: tile over the r conditioning channels and sum (:172-175)
            x = (x.unsqueeze(-1) * (scale.unsqueeze(-2) + 1) + shift.unsqueeze(-2)).sum(-1)
    return F.silu(x)


def _resnet_block(sd, p, x, emb, groups):
    """resnets.py:180-208 (dim == dim_out everywhere on this path, so res_conv is Identity)."""
    scale_shift = None
    if emb is not None and (p + "mlp.1.weight") in sd:
        e = F.linear(F.silu(emb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])
        e = e.unsqueeze(-1) if e.ndim == 2 else e.transpose(1, 2)
        scale_shift = e.chunk(2, dim=1)
    h = _block(sd, p + "block1.", x, groups, scale_shift)
    h = _block(sd, p + "block2.", h, groups)
    if (p + "res_conv.weight") in sd:
        x = F.conv1d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return h + x


def _linear_attention(sd, p, x, heads=4, dim_head=32):
    """resnets.py:211-235 wrapped in Residual(PreNorm(...)) (:59-66,116-124).
    p addresses `blocks.i.2.` so keys are p+'fn.norm.g', p+'fn.fn.to_qkv.weight', ..."""
    B, C, n = x.shape
    xn = _chan_layernorm(x, sd[p + "fn.norm.g"])
    qkv = F.conv1d(xn, sd[p + "fn.fn.to_qkv.weight"]).view(B, 3, heads, dim_head, n)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    q = q.softmax(dim=-2) * dim_head ** -0.5
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q).reshape(B, heads * dim_head, n)
    out = F.conv1d(out, sd[p + "fn.fn.to_out.0.weight"], sd[p + "fn.fn.to_out.0.bias"])
    out = _chan_layernorm(out, sd[p + "fn.fn.to_out.1.g"])
    return out + x


def _resnet_trunk(sd, p, x, emb, groups):
    """Shared body of ResNet1D.forward (resnets.py:397-424) and
    TimeConditionedResNet1D.forward (:584-616); Dropout is identity in eval."""
    x = F.conv1d(x, sd[p + "init_conv.weight"], sd[p + "init_conv.bias"], padding=3)
    i = 0
    while (p + f"blocks.{i}.3.weight") in sd:
        b = p + f"blocks.{i}."
        x = _resnet_block(sd, b + "0.", x, emb, groups)
        x = _resnet_block(sd, b + "1.", x, emb, groups)
        x = _linear_attention(sd, b + "2.", x)
        x = F.conv1d(x, sd[b + "3.weight"], sd[b + "3.bias"], padding=1)
        i += 1
    x = _resnet_block(sd, p + "final_res_block.", x, emb, groups)
    return F.conv1d(x, sd[p + "final_conv.weight"], sd[p + "final_conv.bias"])


def resnet1d_forward(sd, p, x, z_cond, groups=4):
    """ResNet1D.forward, resnets.py:373-424: x [B,1,D], z_cond [B,3,Dc] -> [B,1,D]."""
    emb = F.silu(F.linear(z_cond, sd[p + "input_emb_layers.0.weight"], sd[p + "input_emb_layers.0.bias"]))
    return _resnet_trunk(sd, p, x, emb, groups)


def time_embedding(sd, p, time):
    """time_mlp = RandomOrLearnedSinusoidalPosEmb -> Linear -> GELU -> Linear
    (resnets.py:44-56, :517-522).  `time` is int64 [B]."""
    x = time.view(-1, 1)
    freqs = x * sd[p + "time_mlp.0.weights"].view(1, -1) * 2 * math.pi
    four = torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)
    h = F.linear(four, sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
    return F.linear(F.gelu(h), sd[p + "time_mlp.3.weight"], sd[p + "time_mlp.3.bias"])


def denoiser_forward(sd, p, x, time, z_cond, groups=4, cls_cond=None):
    """TimeConditionedResNet1D.forward, resnets.py:558-616: eps prediction [B,1,D].  With cls_cond [B,1] the class-conditioned
    variant (class_conditioned_resnet.py:48-122): SiLU(Linear(1 -> emb)) of the class is added to the time embedding."""
    emb = time_embedding(sd, p, time)
    if cls_cond is not None:
        emb = emb + F.silu(F.linear(cls_cond.reshape(-1, 1).to(x.dtype), sd[p + "cls_embed.0.weight"], sd[p + "cls_embed.0.bias"]))
    inp = F.silu(F.linear(z_cond, sd[p + "input_emb_layers.0.weight"], sd[p + "input_emb_layers.0.bias"]))
    if inp.ndim == 3:
        emb = emb.unsqueeze(-2).repeat(1, inp.shape[1], 1)
    emb = emb + inp
    return _resnet_trunk(sd, p, x, emb, groups)


# ----------------------------------------------------------------------------- grasp_vae.py
def decoder_forward(sd, p, z_h, cond, groups=4):
    """ConditionalGraspPoseDecoder.forward, R/models/grasp_vae.py:401-436 -> (tmrp [B,6], logits [B,1])."""
    h = F.linear(z_h, sd[p + "in_layer.weight"], sd[p + "in_layer.bias"]).unsqueeze(-2)
    h = resnet1d_forward(sd, p + "net.", h, cond, groups).squeeze(-2)
    return (F.linear(h, sd[p + "tmrp.weight"], sd[p + "tmrp.bias"]),
            F.linear(h, sd[p + "class_logits.weight"], sd[p + "class_logits.bias"]))


# ----------------------------------------------------------------------------- PVCNN encoder
def voxelize_coords(coords, r):
    """Voxelization.forward, R/.../pvcnn/modules/voxelization.py:16-35, normalize=False branch.
    coords f32[B,3,N] -> (int32 voxel coords [B,3,N], float clamped coords [B,3,N])."""
    nc = coords - coords.mean(2, keepdim=True)
    nc = (nc + 1) / 2.0
    nc = torch.clamp(nc * r, 0, r - 1)
    return torch.round(nc).to(torch.int32), nc


def _swish(x):
    return x * torch.sigmoid(x)


def _shared_mlp(sd, p, x):
    """SharedMLP (Conv1d k1 + BatchNorm1d eval + ReLU), R/.../pvcnn/modules/shared_mlp.py:6-35."""
    x = F.conv1d(x, sd[p + "layers.0.weight"], sd[p + "layers.0.bias"])
    x = F.batch_norm(x, sd[p + "layers.1.running_mean"], sd[p + "layers.1.running_var"],
                     sd[p + "layers.1.weight"], sd[p + "layers.1.bias"], False, 0.0, 1e-5)
    return F.relu(x)


def _pvconv(sd, p, feats, coords, r):
    """PVConv.forward, R/.../pvcnn/modules/pvconv.py:76-84 (with_se=True, eval)."""
    B, C, N = feats.shape
    vc, nc = voxelize_coords(coords, r)
    grid, _, _ = ops_np.avg_voxelize_forward(feats.numpy(), vc.numpy(), r)
    v = torch.from_numpy(grid).view(B, C, r, r, r)
    v = F.conv3d(v, sd[p + "voxel_layers.0.weight"], sd[p + "voxel_layers.0.bias"], padding=1)
    v = _swish(F.group_norm(v, 8, sd[p + "voxel_layers.1.weight"], sd[p + "voxel_layers.1.bias"], 1e-5))
    v = F.conv3d(v, sd[p + "voxel_layers.4.weight"], sd[p + "voxel_layers.4.bias"], padding=1)
    v = _swish(F.group_norm(v, 8, sd[p + "voxel_layers.5.weight"], sd[p + "voxel_layers.5.bias"], 1e-5))
    # SE3d, R/.../pvcnn/modules/se.py:12-25
    s = v.mean(-1).mean(-1).mean(-1)
    s = torch.sigmoid(F.linear(_swish(F.linear(s, sd[p + "voxel_layers.7.fc.0.weight"])),
                               sd[p + "voxel_layers.7.fc.2.weight"]))
    v = v * s.view(B, -1, 1, 1, 1)
    Co = v.shape[1]
    dv, _, _ = ops_np.trilinear_devoxelize_forward(r, False, nc.numpy(), v.reshape(B, Co, -1).numpy())
    return torch.from_numpy(dv) + _shared_mlp(sd, p + "point_features.", feats)


def _shared_mlp_n(sd, p, x):
    """Multi-layer SharedMLP over [B,C,N] or [B,C,M,U] (k=1 conv + BatchNorm eval + ReLU per layer), shared_mlp.py:6-35."""
    i = 0
    while (p + f"layers.{i}.weight") in sd:
        w = sd[p + f"layers.{i}.weight"]
        x = torch.einsum("oc,bc...->bo...", w.reshape(w.shape[0], w.shape[1]), x) + \
            sd[p + f"layers.{i}.bias"].view((1, -1) + (1,) * (x.ndim - 2))
        rm, rv = sd[p + f"layers.{i + 1}.running_mean"], sd[p + f"layers.{i + 1}.running_var"]
        sh = (1, -1) + (1,) * (x.ndim - 2)
        x = (x - rm.view(sh)) / torch.sqrt(rv.view(sh) + 1e-5) * sd[p + f"layers.{i + 1}.weight"].view(sh) + \
            sd[p + f"layers.{i + 1}.bias"].view(sh)
        x = F.relu(x)
        i += 3
    return x


def sa_module_forward(sd, p, features, coords, num_centers, radii, num_neighbors):
    """PointNetSAModule.forward, R/.../pvcnn/modules/pointnet.py:100-111 with BallQuery.forward (ball_query.py:16-34):
    FPS -> per radius (ball query -> group (coords - centre | features) -> SharedMLP(dim=2) -> max over neighbours)."""
    t = lambda a: torch.from_numpy(a)
    c = coords.contiguous().numpy()
    centers = t(ops_np.gather_features_forward(c, ops_np.furthest_point_sampling(c, num_centers)))
    outs = []
    for j, (r, u) in enumerate(zip(radii, num_neighbors)):
        idx = ops_np.ball_query(centers.numpy(), c, r, u)
        g = t(ops_np.grouping_forward(c, idx)) - centers.unsqueeze(-1)
        if features is not None:
            g = torch.cat([g, t(ops_np.grouping_forward(features.contiguous().numpy(), idx))], dim=1)
        outs.append(_shared_mlp_n(sd, p + f"mlps.{j}.", g).max(dim=-1).values)
    return (torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]), centers


def fp_module_forward(sd, p, points_coords, centers_coords, centers_features, points_features=None):
    """PointNetFPModule.forward, pointnet.py:122-135: 3-NN inverse-distance interpolation (+ skip features) -> SharedMLP."""
    out, _, _ = ops_np.three_nearest_neighbors_interpolate_forward(points_coords.contiguous().numpy(),
                                                                   centers_coords.contiguous().numpy(),
                                                                   centers_features.contiguous().numpy())
    x = torch.from_numpy(out)
    if points_features is not None:
        x = torch.cat([x, points_features], dim=1)
    return _shared_mlp_n(sd, p + "mlp.", x)


def pvcnn_encoder_forward(sd, p, xyz, resolutions=(24, 12)):
    """PVCNNEncoder.forward, R/models/modules/pc_encoders.py:87-115 over PVCNN.forward
    (R/.../pvcnn/pvcnn_base.py:114-140): xyz [B,N,3] -> z_pc [B,C_out,out_features]."""
    x = xyz.transpose(1, 2).contiguous()
    coords = x[:, :3, :]
    feats = x
    q = p + "pvcnn_modules.point_features."
    feats = _pvconv(sd, q + "0.", feats, coords, resolutions[0])
    feats = _pvconv(sd, q + "1.", feats, coords, resolutions[1])
    feats = _shared_mlp(sd, q + "2.", feats)
    feats = _shared_mlp(sd, q + "3.", feats)
    out = F.conv1d(feats, sd[p + "conv_downscale.weight"], sd[p + "conv_downscale.bias"])
    out = F.conv1d(out, sd[p + "out_layer.0.weight"], sd[p + "out_layer.0.bias"])
    out = F.linear(out, sd[p + "out_layer.1.weight"], sd[p + "out_layer.1.bias"])
    return out.squeeze(1) if out.shape[-2] == 1 else out


# ----------------------------------------------------------------------------- sampling loop
def ldm_sample(sd_ddm, z_cond, x_T, *, noise=None, num_inference_steps=None, kind="ddpm",
               scheduler_kwargs=None, groups=4, return_all=False, prefix="diffusion_model.model.", cls_cond=None):
    """GaussianDiffusion1D.sample, R/models/diffusion/gaussian_diffusion.py:232-277.
    x_T [B,1,D]; noise: optional [n_steps,B,1,D] consumed in loop order (entry i for the i-th
    executed step; the t == 0 entry is ignored as in DDPMScheduler.step)."""
    kw = dict(num_train_timesteps=1000, beta_start=5e-5, beta_end=1e-3, beta_schedule="linear",
              variance_type="fixed_large", clip_sample=True)
    kw.update(scheduler_kwargs or {})
    sch = SchedulerOracle(kind=kind, **kw)
    if num_inference_steps:
        sch.set_timesteps(num_inference_steps)
    x = x_T
    allx = [x_T] if return_all else []
    for i, t in enumerate(timestep_list(sch.T, sch.num_inference_steps)):
        tb = torch.full((x.shape[0],), t, dtype=torch.long)
        eps = denoiser_forward(sd_ddm, prefix, x, tb, z_cond, groups, cls_cond=cls_cond)
        x = sch.step(eps, t, x, None if noise is None else noise[i])
        if return_all:
            allx.append(x)
    return x, allx


# ----------------------------------------------------------------------------- post-processing
def tmrp_to_H(tmrp):
    """R/utils/rotations.py:298-302 -> mrp_to_quat :218-252, quat_to_rotmat :171-215, Rt_to_H :255-274."""
    t, m = tmrp[..., :3], tmrp[..., 3:6]
    magsq = (m * m).sum(-1, keepdim=True)
    qv = (2 * m) / (1 + magsq)
    w = ((1 - magsq) / (1 + magsq))[..., 0]
    x, y, z = qv[..., 0], qv[..., 1], qv[..., 2]
    H = torch.zeros(tmrp.shape[:-1] + (4, 4), dtype=tmrp.dtype)
    H[..., 0, 0] = x * x - y * y - z * z + w * w
    H[..., 1, 0] = 2 * (x * y + z * w)
    H[..., 2, 0] = 2 * (x * z - y * w)
    H[..., 0, 1] = 2 * (x * y - z * w)
    H[..., 1, 1] = -(x * x) + y * y - z * z + w * w
    H[..., 2, 1] = 2 * (y * z + x * w)
    H[..., 0, 2] = 2 * (x * z + y * w)
    H[..., 1, 2] = 2 * (y * z - x * w)
    H[..., 2, 2] = -(x * x) - y * y + z * z + w * w
    H[..., :3, 3] = t
    H[..., 3, 3] = 1
    return H


def generate_grasps_ldm(sd_vae, sd_ddm, xyz, num_grasps, x_T, **sample_kw):
    """GraspLatentDDM.generate_grasps, R/models/grasp_ldm.py:189-233."""
    z_pc = pvcnn_encoder_forward(sd_vae, "encoder.pc_encoder.", xyz)
    z_pc = z_pc.repeat_interleave(num_grasps, dim=0)
    x0, _ = ldm_sample(sd_ddm, z_pc, x_T, **sample_kw)
    return decoder_forward(sd_vae, "decoder.", x0.squeeze(-2), z_pc)


def generate_grasps_vae(sd_vae, xyz, num_grasps, z_h):
    """GraspCVAE.generate_grasps, R/models/grasp_vae.py:226-255 (z_h drawn by the caller)."""
    z_pc = pvcnn_encoder_forward(sd_vae, "encoder.pc_encoder.", xyz).repeat_interleave(num_grasps, dim=0)
    return decoder_forward(sd_vae, "decoder.", z_h, z_pc)


def postprocess(tmrp, logits, xyz, metas, num_pcs, num_grasps):
    """InferenceLDM.generate_grasps tail, R/../tools/inference.py:627-656."""
    tmrp = tmrp.view(num_pcs, num_grasps, 6)
    g = tmrp * metas["grasp_std"].unsqueeze(-2) + metas["grasp_mean"].unsqueeze(-2)
    return dict(grasps=tmrp_to_H(g), grasp_tmrp=g,
                confidence=torch.sigmoid(logits.view(num_pcs, num_grasps, 1)),
                pc=xyz * metas["pc_std"].unsqueeze(-2) + metas["pc_mean"].unsqueeze(-2))


def normalize_input(pc, pc_shift, pc_scale, grasp_shift, grasp_scale):
    """Inference.normalize_input, R/grasp_ldm/inference/inference_base.py:182-212 (metas layout of
    R/tools/inference.py:581-589 for batches): centre on the cloud mean, dataset shift / scale, and the statistics that
    undo it.  The argument is left untouched and `grasp_shift` is not accumulated (the reference does both in place)."""
    assert pc.ndim in (2, 3)
    pc_mean = torch.mean(pc, dim=-2)
    c = pc - (pc_mean.unsqueeze(1) if pc.ndim == 3 else pc_mean)
    c = (c - pc_shift) / pc_scale
    grasp_mean = grasp_shift.clone() if pc.ndim == 2 else grasp_shift.unsqueeze(0).repeat(pc.shape[0], 1)
    grasp_mean[..., :3] += pc_mean
    batched = pc.ndim == 3
    metas = dict(pc_mean=pc_shift + pc_mean, pc_std=pc_scale.unsqueeze(0) if batched else pc_scale,
                 grasp_mean=grasp_mean, grasp_std=grasp_scale.unsqueeze(0) if batched else grasp_scale,
                 use_dataset_statistics=False)
    return c, metas


# ----------------------------------------------------------------------------- elucidated sampler
def edm_sigmas(n, sigma_min=0.002, sigma_max=80.0, rho=7.0):
    """Karras schedule, R/grasp_ldm/models/diffusion/elucidated_diffusion.py:155-168 (eq. 5), with the trailing 0."""
    i = torch.arange(n, dtype=torch.float32)
    s = (sigma_max ** (1 / rho) + i / (n - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat((s, torch.zeros(1)))


def edm_denoise(sd, x, sigma, z_cond, sigma_data=0.5, p="diffusion_model.net."):
    """Preconditioned network (eq. 7), elucidated_diffusion.py:111-153: c_skip x + c_out F(c_in x; log(sigma)/4)."""
    sig = torch.full((x.shape[0],), float(sigma))
    s3 = sig.view(-1, 1, 1)
    c_in = 1 * (s3 ** 2 + sigma_data ** 2) ** -0.5
    c_skip = (sigma_data ** 2) / (s3 ** 2 + sigma_data ** 2)
    c_out = s3 * sigma_data * (sigma_data ** 2 + s3 ** 2) ** -0.5
    f = denoiser_forward(sd, p, c_in * x, torch.log(sig.clamp(min=1e-20)) * 0.25, z_cond)
    return c_skip * x + c_out * f


def edm_sample_heun(sd, z_cond, x_init, noise, n, p="diffusion_model.net.", S_churn=80, S_tmin=0.05, S_tmax=50, S_noise=1.003):
    """Stochastic second-order sampler (Algorithm 2 of Karras et al. 2022), elucidated_diffusion.py:179-258."""
    sig = edm_sigmas(n)
    x = sig[0] * x_init
    for i in range(n):
        s, s_next = float(sig[i]), float(sig[i + 1])
        gamma = min(S_churn / n, math.sqrt(2) - 1) if S_tmin <= s <= S_tmax else 0.0
        s_hat = s + gamma * s
        x_hat = x + math.sqrt(s_hat ** 2 - s ** 2) * (S_noise * noise[i])
        d = (x_hat - edm_denoise(sd, x_hat, s_hat, z_cond, p=p)) / s_hat
        x = x_hat + (s_next - s_hat) * d
        if s_next != 0:
            d2 = (x - edm_denoise(sd, x, s_next, z_cond, p=p)) / s_next
            x = x_hat + 0.5 * (s_next - s_hat) * (d + d2)
    return x


def edm_sample_dpmpp(sd, z_cond, x_init, n, p="diffusion_model.net."):
    """DPM-Solver++(2M), elucidated_diffusion.py:260-315."""
    sig = edm_sigmas(n)
    x = sig[0] * x_init
    prev = None
    for i in range(n):
        den = edm_denoise(sd, x, float(sig[i]), z_cond, p=p)
        t, t_next = -sig[i].log(), -sig[i + 1].log()
        h = t_next - t
        if prev is None or sig[i + 1] == 0:
            dd = den
        else:
            r = (t - (-sig[i - 1].log())) / h
            g = -1 / (2 * r)
            dd = (1 - g) * den + g * prev
        x = ((-t_next).exp() / (-t).exp()) * x - (-h).expm1() * dd
        prev = den
    return x

