"""ElucidatedDiffusion (R/grasp_ldm/models/diffusion/elucidated_diffusion.py) - sampling side.

Same constructor arguments and the same `sample(use_dpmpp=..., batch_size=..., z_cond=..., num_sample_steps=..., clamp=...,
return_all=...)` / `sample_normal` / `sample_using_dpmpp` / `preconditioned_network_forward` surface.  A sampler is compiled
on the host into an EVALUATION PROGRAM - one 16-float row per network evaluation holding the preconditioning constants
(c_in, c_skip, c_out), the update rule (stochastic Heun first / second evaluation, DPM-Solver++(2M) step) and its scalar
coefficients - and the persistent sampler kernel executes the whole program in ONE launch: network evaluations, churn noise,
Heun / multistep updates (csrc/resnet_layout.cuh::eval_update; GLDM_SCHED_EDM in include/graspldm_b200.h).  No Python loop
and no element-wise launches between evaluations.

Keyword-only extensions: `x_init` (the N(0,1) draw that is scaled by sigma_0), `noise` ([num_sample_steps, B, C, L] N(0,1)
draws of the stochastic sampler) for parity runs, `seed` for the in-kernel Philox stream, `grasps_per_object` when z_cond
holds one row per object, `precision`."""
import math

import torch
from torch import nn

from . import engine

HEUN_FIRST, HEUN_SECOND, DPMPP_2M = 0.0, 1.0, 2.0


class ElucidatedDiffusion(nn.Module):
    def __init__(self, net, *, seq_length, channels=1, num_sample_steps=32, sigma_min=0.002, sigma_max=80, sigma_data=0.5,
                 rho=7, P_mean=-1.2, P_std=1.2, S_churn=80, S_tmin=0.05, S_tmax=50, S_noise=1.003):
        super().__init__()
        assert net.random_or_learned_sinusoidal_cond
        self.self_condition = False
        self.net = net
        self.channels, self.seq_length = channels, seq_length
        self.sigma_min, self.sigma_max, self.sigma_data, self.rho = sigma_min, sigma_max, sigma_data, rho
        self.P_mean, self.P_std = P_mean, P_std
        self.num_sample_steps = num_sample_steps
        self.S_churn, self.S_tmin, self.S_tmax, self.S_noise = S_churn, S_tmin, S_tmax, S_noise
        self.precision = "fp32"

    @property
    def device(self):
        return next(self.net.parameters()).device

    def forward(self, *a, **k):
        raise NotImplementedError("ElucidatedDiffusion.forward (training loss) is outside the generation path")

    # ------------------------------------------------------------------ host-side scalar tables (fp32, as the reference)
    def sample_schedule(self, num_sample_steps=None):
        """sigma_i of Karras et al. eq. 5 with a trailing 0 (:152-168)"""
        n = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        i = torch.arange(n, dtype=torch.float32)
        lo, hi = self.sigma_max ** (1 / self.rho), self.sigma_min ** (1 / self.rho)
        return torch.cat([(lo + i / (n - 1) * (hi - lo)) ** self.rho, torch.zeros(1)]).to(self.device)

    def _precondition(self, sigma):
        """(c_in, c_skip, c_out, c_noise) of Table 1 for a float sigma, rounded as fp32 tensor arithmetic would (:111-124)"""
        s = torch.tensor(float(sigma), dtype=torch.float32)
        sd = self.sigma_data
        var = s ** 2 + sd ** 2
        return (float(var ** -0.5), float(sd ** 2 / var), float(s * sd * var ** -0.5), float(torch.log(s.clamp(min=1e-20)) * 0.25))

    def _row(self, sigma, kind, k, noise_slot=-1, out_slot=-1):
        c_in, c_skip, c_out, c_noise = self._precondition(sigma)
        return [c_in, c_skip, c_out, kind, *k, float(noise_slot), float(out_slot)] + [0.0] * 6, c_noise

    def heun_program(self, n):
        """stochastic sampler of Algorithm 2 (:179-258): per step one evaluation at sigma_hat, and - unless the step lands on
        sigma = 0 - the second-order correction at sigma_next"""
        sig = self.sample_schedule(n).cpu()
        churn = min(self.S_churn / n, math.sqrt(2) - 1)
        rows, times = [], []
        for i in range(n):
            s, s_next = float(sig[i]), float(sig[i + 1])
            s_hat = s + (churn if self.S_tmin <= s <= self.S_tmax else 0.0) * s
            last = s_next == 0
            r, t = self._row(s_hat, HEUN_FIRST, [math.sqrt(s_hat ** 2 - s ** 2), s_hat, s_next - s_hat, self.S_noise],
                             noise_slot=i, out_slot=i + 1 if last else -1)
            rows.append(r), times.append(t)
            if not last:
                r, t = self._row(s_next, HEUN_SECOND, [0.0, s_next, 0.5 * (s_next - s_hat), 0.0], out_slot=i + 1)
                rows.append(r), times.append(t)
        return torch.tensor(times, dtype=torch.float32), torch.tensor(rows, dtype=torch.float32), float(sig[0])

    def dpmpp_program(self, n):
        """DPM-Solver++(2M) in the sigma parametrisation (:260-315): t = -log sigma, h = t_next - t, multistep weights from
        the previous step size - all as 0-dim fp32 tensors, as there"""
        sig = self.sample_schedule(n).cpu()
        t_of = lambda s: s.log().neg()
        rows, times = [], []
        for i in range(n):
            t, t_next = t_of(sig[i]), t_of(sig[i + 1])
            h = t_next - t
            if i == 0 or float(sig[i + 1]) == 0:
                w_new, w_old = 1.0, 0.0
            else:
                gamma = -1 / (2 * ((t - t_of(sig[i - 1])) / h))
                w_new, w_old = float(1 - gamma), float(gamma)
            r, tm = self._row(float(sig[i]), DPMPP_2M, [w_new, w_old, float(t_next.neg().exp() / t.neg().exp()), float((-h).expm1())],
                              out_slot=i + 1)
            rows.append(r), times.append(tm)
        return torch.tensor(times, dtype=torch.float32), torch.tensor(rows, dtype=torch.float32), float(sig[0])

    # ------------------------------------------------------------------ sampling
    def _run(self, program, batch_size, z_cond, clamp, return_all, x_init, noise, precision, seed, grasps_per_object, n_slots,
             cls_cond=None):
        times, rows, sigma0 = program
        dev = self.device
        shape = (batch_size, self.channels, self.seq_length)
        x0 = x_init.to(dev).view(shape) if x_init is not None else torch.randn(shape, device=dev)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if noise is None else 0
        if hasattr(self.net, "cls_embed") and cls_cond is None:
            raise RuntimeError("the class-conditioned denoiser needs cls_cond=")
        x, x_all = engine.sampler_program(self.net, sigma0 * x0, z_cond, grasps_per_object, times, rows, n_slots, clip_sample=clamp,
                                          noise=noise, seed=seed, return_all=return_all, precision=precision or self.precision,
                                          cls_cond=cls_cond)
        return x, (list(x_all) if return_all else [])

    def sample(self, **kwargs):
        if kwargs.pop("use_dpmpp"):
            return self.sample_using_dpmpp(**kwargs)
        return self.sample_normal(**kwargs)

    @torch.no_grad()
    def sample_normal(self, batch_size=16, z_cond=None, num_sample_steps=None, clamp=False, return_all=False, *, x_init=None,
                      noise=None, precision=None, seed=None, grasps_per_object=1, cls_cond=None, **kwargs):
        n = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        x, xs = self._run(self.heun_program(n), batch_size, z_cond, clamp, return_all, x_init, noise, precision, seed,
                          grasps_per_object, n + 1, cls_cond)
        return x, xs

    @torch.no_grad()
    def sample_using_dpmpp(self, batch_size=16, z_cond=None, num_sample_steps=20, clamp=False, return_all=False, *, x_init=None,
                           precision=None, seed=None, grasps_per_object=1, cls_cond=None, **kwargs):
        n = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        x, xs = self._run(self.dpmpp_program(n), batch_size, z_cond, clamp, return_all, x_init, None, precision, seed,
                          grasps_per_object, n + 1, cls_cond)
        return x, xs

    @torch.no_grad()
    def preconditioned_network_forward(self, noised_x, sigma, *, z_cond=None, self_cond=None, clamp=False, precision=None,
                                       grasps_per_object=1, cls_cond=None):
        """D(x; sigma) = c_skip x + c_out F(c_in x; c_noise(sigma))  (eq. 7, :126-150) as a one-evaluation program."""
        if torch.is_tensor(sigma):
            if sigma.numel() > 1 and not bool((sigma == sigma.reshape(-1)[0]).all()):
                raise NotImplementedError("per-sample noise levels occur only in the training loss")
            sigma = float(sigma.reshape(-1)[0])
        row, t = self._row(sigma, DPMPP_2M, [1.0, 0.0, 0.0, -1.0])               # x <- 0 * x + 1 * D
        x, _ = engine.sampler_program(self.net, noised_x, z_cond, grasps_per_object, torch.tensor([t]), torch.tensor([row]), 1,
                                      clip_sample=clamp, precision=precision or self.precision, cls_cond=cls_cond)
        return x
