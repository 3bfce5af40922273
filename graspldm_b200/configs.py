"""Model kwargs of the reference's two generation configs as plain dicts
(R/configs/generation/fpc/fpc_1a_latentc3_z4_pc64_180k.py:23-157 and
R/configs/generation/partial_pc/ppc_1a_partial_63cat8k_filtered_latentc3_z16_pc256_180k.py) and the builder
that reproduces tools/inference.py:514-516 (DDM first, then the VAE) so that `torch.manual_seed(s)` followed by
`build_ldm(name)` yields the same random-init weights as the reference."""
import copy

from .grasp_ldm import GraspLatentDDM
from .grasp_vae import GraspCVAE
from .resnets import TimeConditionedResNet1D


def model_config(name="fpc"):
    pc_latent_dims, grasp_latent_dims = {"fpc": (64, 4), "ppc": (256, 16)}[name]
    dropout = 0.1
    resnet_args = dict(block_channels=(32, 64, 128, 256), input_conditioning_dims=pc_latent_dims,
                       resnet_block_groups=4, dropout=dropout)
    vae = dict(
        grasp_latent_size=grasp_latent_dims, pc_latent_size=pc_latent_dims,
        pc_encoder_config=dict(type="PVCNNEncoder", args=dict(in_features=3, n_points=1024, scale_channels=0.75,
                                                              scale_voxel_resolution=0.75, num_blocks=(1, 1, 1, 1),
                                                              out_channels=3, use_global_attention=False)),
        grasp_encoder_config=dict(type="ResNet1D", args=dict(in_features=7, **resnet_args)),
        decoder_config=dict(type="ResNet1D", args=dict(**resnet_args)),
        loss_config=None, num_output_qualities=0, intermediate_feature_resolution=16)
    denoiser = dict(dim=grasp_latent_dims, channels=1, block_channels=(32, 64, 128, 256),
                    input_conditioning_dims=pc_latent_dims, resnet_block_groups=4, dropout=dropout,
                    is_time_conditioned=True, learned_variance=False, learned_sinusoidal_cond=False,
                    random_fourier_features=True)
    ddm = dict(latent_in_features=grasp_latent_dims, diffusion_timesteps=1000, noise_scheduler_type="ddpm",
               diffusion_loss="l2", beta_schedule="linear", is_conditioned=True, joint_training=False,
               denoising_loss_weight=1, variance_type="fixed_large", elucidated_diffusion=False,
               beta_start=0.00005, beta_end=0.001)
    return dict(vae=vae, denoiser=denoiser, ddm=ddm)


def build_ldm(name="fpc", noise_scheduler_type="ddpm"):
    cfg = copy.deepcopy(model_config(name))
    cfg["ddm"]["noise_scheduler_type"] = noise_scheduler_type
    denoiser = TimeConditionedResNet1D(**cfg["denoiser"])
    model = GraspLatentDDM(model=denoiser, **cfg["ddm"])
    model.set_vae_model(GraspCVAE(**cfg["vae"]))
    return model.eval()


def build_vae(name="fpc"):
    return GraspCVAE(**copy.deepcopy(model_config(name))["vae"]).eval()
