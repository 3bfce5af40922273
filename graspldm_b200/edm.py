"""ElucidatedDiffusion (R/grasp_ldm/models/diffusion/elucidated_diffusion.py) - sampling side.

Same constructor arguments, `sample(use_dpmpp=..., batch_size=..., z_cond=..., num_sample_steps=..., clamp=...,
return_all=...)`, `sample_normal` (stochastic Heun sampler of Karras et al. 2022, :179-258) and `sample_using_dpmpp`
(DPM-Solver++ 2M, :260-315).  Every network evaluation is one launch of the denoiser kernel (tensor-core or strict fp32,
continuous time c_noise(sigma) = log(sigma) / 4 through `gldm_denoiser_forward_*_ftime`); the update rule between two
evaluations is O(16 bytes) per sample and stays in a handful of element-wise operations on the device.  Unlike the
DDPM / DDIM path the loop is not fused into one persistent launch yet.

Keyword-only extensions for parity runs: `x_init` (the N(0,1) draw that is scaled by sigma_0) and `noise`
([num_sample_steps, B, C, L] N(0,1) draws of the stochastic sampler), `precision`."""
from math import sqrt

import torch
from torch import nn


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

    # preconditioning (Table 1 of the paper; reference :111-124)
    def c_skip(self, sigma):
        return (self.sigma_data ** 2) / (sigma ** 2 + self.sigma_data ** 2)

    def c_out(self, sigma):
        return sigma * self.sigma_data * (self.sigma_data ** 2 + sigma ** 2) ** -0.5

    def c_in(self, sigma):
        return 1 * (sigma ** 2 + self.sigma_data ** 2) ** -0.5

    def c_noise(self, sigma):
        return torch.log(sigma.clamp(min=1e-20)) * 0.25

    def preconditioned_network_forward(self, noised_x, sigma, *, z_cond=None, self_cond=None, clamp=False, precision=None):
        batch, device = noised_x.shape[0], noised_x.device
        if isinstance(sigma, float):
            sigma = torch.full((batch,), sigma, device=device)
        padded = sigma.view(-1, 1, 1)
        net_out = self.net(self.c_in(padded) * noised_x, time=self.c_noise(sigma), z_cond=z_cond,
                           precision=precision or self.precision)
        out = self.c_skip(padded) * noised_x + self.c_out(padded) * net_out
        return out.clamp(-1.0, 1.0) if clamp else out

    def sample_schedule(self, num_sample_steps=None):
        n = num_sample_steps if num_sample_steps is not None else self.num_sample_steps
        inv_rho = 1 / self.rho
        steps = torch.arange(n, device=self.device, dtype=torch.float32)
        sigmas = (self.sigma_max ** inv_rho + steps / (n - 1) * (self.sigma_min ** inv_rho - self.sigma_max ** inv_rho)) ** self.rho
        return torch.nn.functional.pad(sigmas, (0, 1), value=0.0)

    def sample(self, **kwargs):
        if kwargs.pop("use_dpmpp"):
            return self.sample_using_dpmpp(**kwargs)
        return self.sample_normal(**kwargs)

    def _draw(self, given, shape):
        return given.to(self.device).view(shape) if given is not None else torch.randn(shape, device=self.device)

    @torch.no_grad()
    def sample_normal(self, batch_size=16, z_cond=None, num_sample_steps=None, clamp=False, return_all=False, *, x_init=None,
                      noise=None, precision=None):
        n = num_sample_steps if num_sample_steps is not None else self.num_sample_steps
        shape = (batch_size, self.channels, self.seq_length)
        sigmas = self.sample_schedule(n)
        gammas = torch.where((sigmas >= self.S_tmin) & (sigmas <= self.S_tmax), min(self.S_churn / n, sqrt(2) - 1), 0.0)
        x = sigmas[0] * self._draw(x_init, shape)
        all_x = [x]
        for i, (sigma, sigma_next, gamma) in enumerate(zip(sigmas[:-1].tolist(), sigmas[1:].tolist(), gammas[:-1].tolist())):
            eps = self.S_noise * self._draw(None if noise is None else noise[i], shape)
            sigma_hat = sigma + gamma * sigma
            x_hat = x + sqrt(sigma_hat ** 2 - sigma ** 2) * eps
            out = self.preconditioned_network_forward(x_hat, sigma_hat, z_cond=z_cond, clamp=clamp, precision=precision)
            d = (x_hat - out) / sigma_hat
            x_next = x_hat + (sigma_next - sigma_hat) * d
            if sigma_next != 0:           # second-order (Heun) correction
                out_next = self.preconditioned_network_forward(x_next, sigma_next, z_cond=z_cond, clamp=clamp,
                                                               precision=precision)
                d_prime = (x_next - out_next) / sigma_next
                x_next = x_hat + 0.5 * (sigma_next - sigma_hat) * (d + d_prime)
            x = x_next
            all_x += [x] if return_all else []
        return x, all_x

    @torch.no_grad()
    def sample_using_dpmpp(self, batch_size=16, z_cond=None, num_sample_steps=20, clamp=False, return_all=False, *,
                           x_init=None, precision=None):
        n = num_sample_steps if num_sample_steps is not None else self.num_sample_steps
        sigmas = self.sample_schedule(n)
        shape = (batch_size, self.channels, self.seq_length)
        x = sigmas[0] * self._draw(x_init, shape)
        all_x = [x]
        sigma_fn = lambda t: t.neg().exp()
        t_fn = lambda sigma: sigma.log().neg()
        old_denoised = None
        for i in range(len(sigmas) - 1):
            denoised = self.preconditioned_network_forward(x, sigmas[i].item(), z_cond=z_cond, clamp=clamp, precision=precision)
            t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
            h = t_next - t
            if old_denoised is None or sigmas[i + 1] == 0:
                denoised_d = denoised
            else:
                h_last = t - t_fn(sigmas[i - 1])
                r = h_last / h
                gamma = -1 / (2 * r)
                denoised_d = (1 - gamma) * denoised + gamma * old_denoised
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * denoised_d
            all_x += [x] if return_all else []
            old_denoised = denoised
        return x, all_x

    def forward(self, *a, **k):
        raise NotImplementedError("ElucidatedDiffusion.forward (training loss) is outside the generation path")
