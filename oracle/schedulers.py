"""CPU oracle for the noise-scheduler arithmetic.  TEST INFRASTRUCTURE ONLY.

The reference delegates the reverse-diffusion update to the third-party package
`diffusers` (requirements.txt:4, `diffusers[torch]`, UNPINNED; call sites
grasp_ldm/models/diffusion/gaussian_diffusion.py:146-160 (constructor kwargs),
:122 (set_timesteps), :272 (step(...).prev_sample), :112 (num_inference_steps)).
`diffusers` is not present in this image and cannot be fetched, so this file
restates the published algorithm of DDPMScheduler.step / DDIMScheduler.step as
implemented by diffusers >= 0.15 (Ho et al. 2020 eq. 7 / Song et al. 2021 eq. 12):

  * betas      = linspace(beta_start, beta_end, T, float32)     ("linear")
  * alphas_cumprod = cumprod(1 - betas)                          (float32)
  * prev_t     = t - T // num_inference_steps   (num_inference_steps -> T when unset)
  * DDPM: current_alpha_t = a_t / a_prev, x0 = (x - sqrt(1-a_t) eps) / sqrt(a_t),
          clamp(x0, -1, 1) when clip_sample, mu = c0 x0 + c1 x,
          x_prev = mu + sqrt(var) z for t > 0, var = current_beta_t for "fixed_large",
          clamp((1-a_prev)/(1-a_t) current_beta_t, 1e-20) for "fixed_small"
  * DDIM (eta=0, set_alpha_to_one=True): x_prev = sqrt(a_prev) x0 + sqrt(1-a_prev) eps

PARITY UNPINNED for this file: the reference has no test, golden vector or pinned
version at this boundary.  All coefficients are 0-dim float32
torch tensors combined in the same operator order as diffusers so that the tables
are reproducible bit for bit.  What is pinned offline: tests/test_scheduler_equations_cpu.py
checks these coefficients (and the product's tables) against the papers' closed forms
evaluated independently in float64.
"""
import torch


def make_betas(num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2, beta_schedule="linear"):
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "squaredcos_cap_v2":
        import math
        betas = []
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        for i in range(num_train_timesteps):
            t1, t2 = i / num_train_timesteps, (i + 1) / num_train_timesteps
            betas.append(min(1 - f(t2) / f(t1), 0.999))
        return torch.tensor(betas, dtype=torch.float32)
    raise NotImplementedError(beta_schedule)


class SchedulerOracle:
    """Minimal DDPM/DDIM scheduler with the attribute/method surface the reference uses."""

    def __init__(self, kind="ddpm", num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2,
                 beta_schedule="linear", variance_type="fixed_small", prediction_type="epsilon",
                 clip_sample=True, clip_sample_range=1.0):
        assert kind in ("ddpm", "ddim")
        assert prediction_type == "epsilon"
        self.kind = kind
        self.T = num_train_timesteps
        self.variance_type = variance_type
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range
        self.betas = make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None

    def set_timesteps(self, n):
        if n > self.T:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = n

    def _prev(self, t):
        n = self.num_inference_steps if self.num_inference_steps else self.T
        return t - self.T // n

    def coefficients(self, t):
        """0-dim float32 tensors used by one step at integer timestep t."""
        prev_t = self._prev(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - a_t
        beta_prod_prev = 1 - a_prev
        c = dict(sqrt_beta_prod_t=beta_prod_t ** 0.5, sqrt_alpha_prod_t=a_t ** 0.5)
        if self.kind == "ddpm":
            cur_alpha = a_t / a_prev
            cur_beta = 1 - cur_alpha
            c["x0_coeff"] = (a_prev ** 0.5 * cur_beta) / beta_prod_t
            c["xt_coeff"] = cur_alpha ** 0.5 * beta_prod_prev / beta_prod_t
            var = torch.clamp(beta_prod_prev / beta_prod_t * cur_beta, min=1e-20)
            if self.variance_type == "fixed_large":
                var = cur_beta
                # diffusers special-cases t == 1 for fixed_large only when num_inference_steps == T
                # (variance = beta_1 ... ) - irrelevant here: noise is added only for t > 0 with var as is.
            elif self.variance_type != "fixed_small":
                raise NotImplementedError(self.variance_type)
            c["sigma"] = var ** 0.5 if t > 0 else torch.tensor(0.0)
        else:
            c["x0_coeff"] = a_prev ** 0.5
            c["eps_coeff"] = (1 - a_prev - torch.tensor(0.0) ** 2) ** 0.5
        return c

    def step(self, eps, t, x, noise=None):
        c = self.coefficients(int(t))
        x0 = (x - c["sqrt_beta_prod_t"] * eps) / c["sqrt_alpha_prod_t"]
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        if self.kind == "ddpm":
            prev = c["x0_coeff"] * x0 + c["xt_coeff"] * x
            if int(t) > 0:
                if noise is None:
                    noise = torch.randn(eps.shape, dtype=eps.dtype)
                prev = prev + c["sigma"] * noise
            return prev
        return c["x0_coeff"] * x0 + c["eps_coeff"] * eps


def timestep_list(T, num_inference_steps):
    """gaussian_diffusion.py:258-266: reversed(range(0, T, T // n_inf)), n_inf -> T when unset."""
    n = num_inference_steps if num_inference_steps else T
    return list(reversed(range(0, T, int(T // n))))
