"""Host-side noise-schedule tables for the fused sampler kernel.

The reference delegates the update rule to diffusers' DDPMScheduler / DDIMScheduler
(R/grasp_ldm/models/diffusion/gaussian_diffusion.py:146-160, :272).  The kernel applies the update itself;
this module only produces, per executed step, the integer timestep and the float32 coefficients, computed
with 0-dim float32 torch tensors in diffusers' operator order (>= 0.15: prev_t = t - T // n_inference).
"""
import torch

DDPM, DDIM = 0, 1


def _betas(T, beta_start, beta_end, schedule):
    if schedule == "linear":
        return torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2
    if schedule == "squaredcos_cap_v2":
        import math
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return torch.tensor([min(1 - f((i + 1) / T) / f(i / T), 0.999) for i in range(T)], dtype=torch.float32)
    raise NotImplementedError(schedule)


class NoiseSchedule:
    """Attribute surface the reference touches on a diffusers scheduler: num_inference_steps, set_timesteps."""

    def __init__(self, kind, num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2, beta_schedule="linear",
                 variance_type="fixed_small", prediction_type="epsilon", clip_sample=True):
        if prediction_type != "epsilon":
            raise NotImplementedError("epsilon prediction only")
        if kind == "ddpm" and variance_type not in ("fixed_small", "fixed_large"):
            raise NotImplementedError(f"variance_type={variance_type}")
        self.kind = kind
        self.T = int(num_train_timesteps)
        self.variance_type = variance_type
        self.clip_sample = bool(clip_sample)
        self.alphas_cumprod = torch.cumprod(1.0 - _betas(self.T, beta_start, beta_end, beta_schedule), dim=0)
        self.num_inference_steps = None
        self._tables = {}

    def set_timesteps(self, n):
        if n > self.T:
            raise ValueError(f"num_inference_steps {n} cannot exceed num_train_timesteps {self.T}")
        self.num_inference_steps = int(n)

    def timesteps(self):
        """gaussian_diffusion.py:258-266: reversed(range(0, T, T // n_inf))"""
        n = self.num_inference_steps if self.num_inference_steps else self.T
        return list(reversed(range(0, self.T, int(self.T // n))))

    def table(self):
        """-> (timesteps list[int], coef float32 [n_steps, 8]):
        [sqrt(1-abar_t), sqrt(abar_t), c_x0, c_xt (DDPM) | c_eps (DDIM), sigma, 0, 0, 0]"""
        key = self.num_inference_steps
        if key not in self._tables:
            n = self.num_inference_steps if self.num_inference_steps else self.T
            stride = self.T // n
            ts = self.timesteps()
            one = torch.tensor(1.0)
            rows = []
            for t in ts:
                prev_t = t - stride
                a_t = self.alphas_cumprod[t]
                a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else one
                bp_t, bp_prev = 1 - a_t, 1 - a_prev
                row = [bp_t ** 0.5, a_t ** 0.5]
                if self.kind == "ddpm":
                    cur_alpha = a_t / a_prev
                    cur_beta = 1 - cur_alpha
                    row += [(a_prev ** 0.5 * cur_beta) / bp_t, cur_alpha ** 0.5 * bp_prev / bp_t]
                    var = cur_beta if self.variance_type == "fixed_large" else torch.clamp(bp_prev / bp_t * cur_beta, min=1e-20)
                    row.append(var ** 0.5 if t > 0 else torch.tensor(0.0))
                else:
                    row += [a_prev ** 0.5, (1 - a_prev) ** 0.5, torch.tensor(0.0)]
                rows.append(torch.stack([r.float() for r in row] + [torch.tensor(0.0)] * 3))
            self._tables[key] = (ts, torch.stack(rows).contiguous())
        return self._tables[key]
