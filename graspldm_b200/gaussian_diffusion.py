"""GaussianDiffusion1D (R/grasp_ldm/models/diffusion/gaussian_diffusion.py) - sampling side.

`sample` keeps the reference signature; the whole T-step loop (denoiser, scheduler update, noise) runs as
one persistent CUDA kernel (csrc/resnet1d_*.cu) instead of one Python iteration and ~300 launches per step.
"""
from typing import Tuple

import torch
from torch import Tensor, nn

from . import engine
from .schedulers import DDIM, DDPM, NoiseSchedule


class GaussianDiffusion1D(nn.Module):
    NOISE_SCHEDULERS = ["ddpm", "ddim"]
    BETA_SCHEDULES = ["linear", "scaled_linear", "squaredcos_cap_v2", "cosine"]

    def __init__(self, model: nn.Module, n_dims: int, noise_scheduler_type: str = "ddpm", beta_schedule: str = "linear",
                 variance_type: str = "fixed_small", pred_type: str = "epsilon", beta_start=0.0001, beta_end=0.02,
                 num_steps: int = 1000, loss_type: str = "l1", clip_sample=True) -> None:
        super().__init__()
        assert noise_scheduler_type in self.NOISE_SCHEDULERS, f"{noise_scheduler_type} Not supported"
        assert beta_schedule in self.BETA_SCHEDULES, f"{beta_schedule} not supported"
        self.num_train_timesteps = self.num_steps = num_steps
        self.beta_start, self.beta_end = beta_start, beta_end
        self.beta_schedule = beta_schedule if beta_schedule != "cosine" else "squaredcos_cap_v2"
        self.variance_type, self.pred_type, self.clip_sample = variance_type, pred_type, clip_sample
        self.model = model
        self.n_dims = n_dims
        self.channels = 1
        self.loss_type = loss_type
        self._noise_scheduler_type = noise_scheduler_type
        self.noise_scheduler = NoiseSchedule(noise_scheduler_type, num_steps, beta_start, beta_end, self.beta_schedule,
                                             variance_type, pred_type, clip_sample)
        assert self.model.out_channels == 1, "fixed-variance samplers need a single eps channel"
        self.precision = "fp32"       # "bf16": tcgen05 tensor-core sampler (bf16 operands, fp32 accumulation)
        self.rng_mode = "reference"   # "reference": per-step torch.randn on the device generator, as diffusers;
                                      # "fused": Philox4x32-10 + Box-Muller inside the sampler kernel

    @property
    def num_inference_steps(self):
        inf_t = self.noise_scheduler.num_inference_steps
        return inf_t if inf_t is not None else self.num_steps

    def set_inference_timesteps(self, num_steps):
        self.noise_scheduler.set_timesteps(num_steps)

    def forward(self, *a, **k):
        raise NotImplementedError("training (noise-prediction loss) is outside the generation path")

    @torch.no_grad()
    def sample(self, z_cond: Tensor = None, batch_size: int = 1, return_all: bool = False,
               device: torch.device = "cuda:0", *, x_T: Tensor = None, noise: Tensor = None,
               grasps_per_object: int = 1, seed: int = None, **kwargs) -> Tuple[Tensor, list]:
        """Reverse diffusion.  Reference arguments: z_cond [B,C,Dc] (already repeated per grasp), batch_size,
        return_all, device.  Extensions (keyword-only): x_T and per-step `noise` [n_steps,B,1,D] for parity runs,
        `grasps_per_object` when z_cond holds one row per OBJECT (avoids materialising the repeat),
        `seed` for the fused RNG, `cls_cond` for a ClassTimeConditionedResNet1D (default: kwargs["metas"]["mode_cls"], as
        the reference's denoiser reads it).  Other extra kwargs (e.g. metas=) are ignored, as the reference's denoiser does."""
        device = z_cond.device if z_cond is not None and z_cond.is_cuda else torch.device(device)
        if x_T is None:
            # gaussian_diffusion.py:253 - drawn on the CPU generator, then moved
            # same values as the reference (CPU generator); drawn into pinned memory so the copy does not make the
            # launching thread wait for the encoder kernels that are still in flight
            x_T = torch.randn((batch_size, self.channels, self.n_dims), pin_memory=True).to(device, non_blocking=True)
        assert x_T.shape[0] == batch_size == z_cond.shape[0] * grasps_per_object
        ts, coef = self.noise_scheduler.table()
        kind = DDPM if self._noise_scheduler_type == "ddpm" else DDIM
        if noise is None and kind == DDPM and self.rng_mode == "reference":
            # diffusers draws randn(model_output.shape, device=cuda) once per step with t > 0, in loop order
            with torch.cuda.device(device):
                noise = torch.stack([torch.randn((batch_size, self.channels, self.n_dims), device=device)
                                     if t > 0 else torch.zeros((batch_size, self.channels, self.n_dims), device=device)
                                     for t in ts])
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if noise is None else 0
        cls_cond = None
        if hasattr(self.model, "cls_embed"):
            cls_cond = self.model.class_condition(kwargs.get("cls_cond"), kwargs)
        x0, x_all = engine.sampler_run(self.model, x_T, z_cond, grasps_per_object, ts, coef, kind, self.clip_sample,
                                       noise=noise, seed=seed, return_all=return_all,
                                       precision=kwargs.get("precision", self.precision), cls_cond=cls_cond)
        return x0, (list(x_all) if return_all else [])
