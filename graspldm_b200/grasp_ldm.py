"""GraspLatentDDM (R/grasp_ldm/models/grasp_ldm.py) - generation side."""
import torch
from torch import nn

from .edm import ElucidatedDiffusion
from .gaussian_diffusion import GaussianDiffusion1D


class GraspLatentDDM(nn.Module):
    def __init__(self, model, latent_in_features, diffusion_timesteps, diffusion_loss, beta_schedule="linear",
                 noise_scheduler_type: str = "ddpm", is_conditioned=True, joint_training=False, denoising_loss_weight=1,
                 variance_type="fixed_small", elucidated_diffusion=False, beta_start=5e-5, beta_end=5e-2) -> None:
        super().__init__()
        self.vae_model = None
        self.is_elucidated_diffusion = elucidated_diffusion
        if elucidated_diffusion:            # grasp_ldm.py:59-62
            self.diffusion_model = ElucidatedDiffusion(net=model, seq_length=latent_in_features)
        else:
            self.diffusion_model = GaussianDiffusion1D(model=model, n_dims=latent_in_features, num_steps=diffusion_timesteps,
                                                       loss_type=diffusion_loss, beta_schedule=beta_schedule,
                                                       beta_start=beta_start, beta_end=beta_end,
                                                       noise_scheduler_type=noise_scheduler_type, variance_type=variance_type)
        self.is_conditioned, self.joint_training, self.loss_weight = is_conditioned, joint_training, denoising_loss_weight
        self.is_vae_frozen = False

    @property
    def _type(self) -> str:
        return self.__class__.__name__

    @property
    def use_grasp_qualities(self):
        return self.vae_model.use_grasp_qualities

    @property
    def scheduler_type(self):
        return self.diffusion_model._noise_scheduler_type

    def set_vae_model(self, vae_model):
        self.vae_model = vae_model

    def load_vae_weights(self, state_dict):
        self.vae_model.load_state_dict(state_dict, strict=True)

    def set_inference_timesteps(self, num_inference_steps):
        self.diffusion_model.set_inference_timesteps(num_inference_steps)

    def forward(self, *a, **k):
        raise NotImplementedError("GraspLatentDDM.forward (training) is outside the generation path")

    @torch.no_grad()
    def generate_grasps(self, xyz, num_grasps=10, return_intermediate=False, **kwargs):
        """grasp_ldm.py:189-233: encode_pc -> T-step latent denoising -> decode.
        Returns ((tmrp [B*G,6], logits [B*G,1]), steps) with steps == [] unless return_intermediate.
        The per-object latent is indexed by sample // num_grasps inside the kernels instead of being
        repeat_interleaved in HBM (:207).  Sampler extensions (x_T=, noise=, seed=) pass through kwargs."""
        z_pc = self.vae_model.encode_pc(xyz)
        n = z_pc.shape[0] * num_grasps
        if self.is_elucidated_diffusion:
            # the reference forwards **kwargs (use_dpmpp, num_sample_steps, clamp) to ElucidatedDiffusion.sample (:214-219);
            # the object latent is indexed per grasp inside the kernel instead of being repeated in HBM
            kw = {k: kwargs[k] for k in ("use_dpmpp", "num_sample_steps", "clamp", "x_init", "noise", "precision", "seed", "cls_cond")
                  if k in kwargs}
            out, all_outs = self.diffusion_model.sample(z_cond=z_pc, batch_size=n, return_all=return_intermediate,
                                                        grasps_per_object=num_grasps, **kw)
            res = self.vae_model.decoder(out.squeeze(-2), z_pc, grasps_per_object=num_grasps)
            if not return_intermediate:
                return (res, [])
            step_outs = []
            for idx in torch.linspace(0, len(all_outs) - 1, steps=50, dtype=torch.int):
                _out = self.vae_model.decoder(all_outs[idx].squeeze(-2), z_pc, grasps_per_object=num_grasps)
                step_outs.append([t.detach().cpu() for t in _out])
            return res, step_outs
        sample_kw = {k: kwargs[k] for k in ("x_T", "noise", "seed", "device", "cls_cond", "metas") if k in kwargs}
        out, all_outs = self.diffusion_model.sample(z_cond=z_pc, batch_size=n, return_all=return_intermediate,
                                                    grasps_per_object=num_grasps, **sample_kw)
        res = self.vae_model.decoder(out.squeeze(-2), z_pc, grasps_per_object=num_grasps)
        if not return_intermediate:
            return (res, [])
        step_outs = []
        for idx in torch.linspace(0, len(all_outs) - 1, steps=50, dtype=torch.int):
            _out = self.vae_model.decoder(all_outs[idx].squeeze(-2), z_pc, grasps_per_object=num_grasps)
            step_outs.append([t.detach().cpu() for t in _out])
        return res, step_outs
