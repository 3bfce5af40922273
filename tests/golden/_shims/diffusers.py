"""Stand-in for the third-party `diffusers` package (absent from this image, unpinned in the
reference).  Routes DDPMScheduler / DDIMScheduler to oracle/schedulers.py so the reference's own
sampling loop (gaussian_diffusion.py:232-277) can run in the build container.  Used ONLY by
tests/golden/make_golden.py."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))
from oracle.schedulers import SchedulerOracle  # noqa: E402


class _Out:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


class _Base:
    KIND = None

    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2, beta_schedule="linear",
                 variance_type="fixed_small", prediction_type="epsilon", clip_sample=True):
        self._o = SchedulerOracle(self.KIND, num_train_timesteps, beta_start, beta_end, beta_schedule,
                                  variance_type, prediction_type, clip_sample)
        self.injected_noise = None   # list of tensors consumed in call order (golden generation)
        self._calls = 0

    @property
    def num_inference_steps(self):
        return self._o.num_inference_steps

    def set_timesteps(self, n):
        self._o.set_timesteps(n)

    def step(self, model_output, timestep, sample):
        noise = None
        if self.injected_noise is not None:
            noise = self.injected_noise[self._calls]
        self._calls += 1
        return _Out(self._o.step(model_output, timestep, sample, noise))


class DDPMScheduler(_Base):
    KIND = "ddpm"


class DDIMScheduler(_Base):
    KIND = "ddim"

    def __init__(self, **kw):
        kw.setdefault("variance_type", "fixed_small")
        super().__init__(**kw)
