"""Scheduler arithmetic against the PUBLISHED equations (no GPU).

`diffusers` is absent from the image and from /root/reference, so `oracle/schedulers.py` cannot be pinned by running the
third-party code ("parity unpinned" for this piece).  What can be pinned offline is that the oracle -
and the product's coefficient tables, which tests/test_oracle_golden.py ties to it bit for bit - satisfy the closed forms of
the papers, evaluated here independently in float64 from the beta schedule alone:

  Ho et al. 2020, eq. 6-7 :  q(x_{t-1} | x_t, x_0) = N(mu~_t, beta~_t I),
        mu~_t = sqrt(abar_{t-1}) beta_t / (1 - abar_t) x_0 + sqrt(alpha_t) (1 - abar_{t-1}) / (1 - abar_t) x_t,
        beta~_t = (1 - abar_{t-1}) / (1 - abar_t) beta_t          ("fixed_small"; "fixed_large" is beta_t)
  Song et al. 2021, eq. 12 (eta = 0):  x_{t-1} = sqrt(abar_{t-1}) x0_hat + sqrt(1 - abar_{t-1}) eps
  strided sampling (n < T steps): the same with alpha_t := abar_t / abar_prev, prev = t - T // n, abar_{-1} := 1

plus the identities any correct posterior must satisfy (they involve no implementation detail at all):
  c_x0 + c_xt sqrt(abar_t) = sqrt(abar_prev)                       (the posterior mean keeps the forward marginal's mean)
  c_xt^2 (1 - abar_t) + beta~_t = 1 - abar_prev                    (... and its variance)
"""
import math

import numpy as np
import pytest
import torch

from graspldm_b200.schedulers import NoiseSchedule
from oracle import schedulers as S


def _abar64(T, beta_start=1e-4, beta_end=2e-2):
    return np.cumprod(1.0 - np.linspace(beta_start, beta_end, T, dtype=np.float64))


def _rtol(small):
    """float32 tolerance of a quantity that carries the cancellation 1 - x with 1 - x = `small` (half an ulp of 1 is 6e-8, and
    the float32 cumprod behind abar has drifted by a few ulps after hundreds of factors)"""
    return max(5e-5, 1e-6 / small)


# constructor defaults (gaussian_diffusion.py:30-31) and the betas of the reference's fpc / ppc configs
BETAS = [(1e-4, 2e-2), (5e-5, 1e-3)]


@pytest.mark.parametrize("betas", BETAS)
@pytest.mark.parametrize("n_steps", [1000, 100, 10])
@pytest.mark.parametrize("variance_type", ["fixed_small", "fixed_large"])
def test_ddpm_coefficients_are_the_posterior_of_ho_et_al(n_steps, variance_type, betas):
    T = 1000
    abar = _abar64(T, *betas)
    sch = S.SchedulerOracle("ddpm", num_train_timesteps=T, variance_type=variance_type, beta_start=betas[0], beta_end=betas[1])
    sch.set_timesteps(n_steps)
    ts = S.timestep_list(T, n_steps)
    assert ts[0] == T - T // n_steps and ts[-1] == 0 and len(ts) == n_steps
    for t in ts:
        prev = t - T // n_steps
        a_t, a_p = abar[t], (abar[prev] if prev >= 0 else 1.0)
        alpha = a_t / a_p
        beta = 1.0 - alpha
        c = {k: float(v) for k, v in sch.coefficients(t).items()}
        want_x0 = math.sqrt(a_p) * beta / (1 - a_t)
        want_xt = math.sqrt(alpha) * (1 - a_p) / (1 - a_t)
        tol = max(_rtol(beta), _rtol(1 - a_t))
        np.testing.assert_allclose(c["x0_coeff"], want_x0, rtol=tol, atol=1e-7)
        np.testing.assert_allclose(c["xt_coeff"], want_xt, rtol=max(tol, _rtol(1 - a_p) if prev >= 0 else 0), atol=2e-6)
        var = beta if variance_type == "fixed_large" else max((1 - a_p) / (1 - a_t) * beta, 1e-20)
        np.testing.assert_allclose(c["sigma"], math.sqrt(var) if t > 0 else 0.0, rtol=max(tol, _rtol(1 - a_p) if prev >= 0 else 0), atol=2e-6)
        # the identities, on the oracle's own float32 numbers
        tol_p = max(tol, _rtol(1 - a_p) if prev >= 0 else 0)
        np.testing.assert_allclose(c["x0_coeff"] + c["xt_coeff"] * math.sqrt(a_t), math.sqrt(a_p), rtol=tol_p, atol=2e-6)
        if variance_type == "fixed_small" and t > 0:
            np.testing.assert_allclose(c["xt_coeff"] ** 2 * (1 - a_t) + c["sigma"] ** 2, 1 - a_p, rtol=3 * tol_p, atol=1e-7)


def test_true_noise_recovers_x0_and_the_posterior_mean():
    """With eps = the noise that produced x_t, step() must return exactly the posterior mean of eq. 7 evaluated at the TRUE
    x_0 (clipping inactive: |x_0| < 1), for a full-length and a strided schedule."""
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(64, 4, generator=g) * 1.6 - 0.8
    eps = torch.randn(64, 4, generator=g)
    T = 1000
    abar = _abar64(T)
    for n_steps in (1000, 100):
        sch = S.SchedulerOracle("ddpm", num_train_timesteps=T, variance_type="fixed_large")
        sch.set_timesteps(n_steps)
        for t in (999 if n_steps == 1000 else 990, 500, 10 * (T // n_steps) if n_steps == 100 else 37, 0):
            prev = t - T // n_steps
            a_t, a_p = abar[t], (abar[prev] if prev >= 0 else 1.0)
            x_t = math.sqrt(a_t) * x0.double() + math.sqrt(1 - a_t) * eps.double()
            got = sch.step(eps, t, x_t.float(), noise=torch.zeros_like(eps))
            alpha = a_t / a_p
            mu = math.sqrt(a_p) * (1 - alpha) / (1 - a_t) * x0.double() + math.sqrt(alpha) * (1 - a_p) / (1 - a_t) * x_t
            torch.testing.assert_close(got.double(), mu, rtol=1e-4, atol=2e-4)
            if t == 0:                                   # the last step returns x_0 itself and adds no noise
                torch.testing.assert_close(got.double(), x0.double(), rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("n_steps", [5, 10, 50])
def test_ddim_chain_with_the_true_noise_returns_x0(n_steps):
    """Song et al. eq. 12 with eta = 0: along the deterministic trajectory through (x_0, eps) every step lands on
    sqrt(abar_prev) x_0 + sqrt(1 - abar_prev) eps, and the last one (abar_prev := 1, set_alpha_to_one) on x_0."""
    g = torch.Generator().manual_seed(1)
    x0 = torch.rand(32, 4, generator=g) * 1.6 - 0.8
    eps = torch.randn(32, 4, generator=g)
    T = 1000
    abar = _abar64(T)
    sch = S.SchedulerOracle("ddim", num_train_timesteps=T)
    sch.set_timesteps(n_steps)
    ts = S.timestep_list(T, n_steps)
    x = (math.sqrt(abar[ts[0]]) * x0.double() + math.sqrt(1 - abar[ts[0]]) * eps.double()).float()
    for t in ts:
        x = sch.step(eps, t, x)
        prev = t - T // n_steps
        a_p = abar[prev] if prev >= 0 else 1.0
        torch.testing.assert_close(x.double(), math.sqrt(a_p) * x0.double() + math.sqrt(1 - a_p) * eps.double(), rtol=1e-4, atol=3e-4)
    torch.testing.assert_close(x, x0, rtol=1e-4, atol=3e-4)


def test_clip_sample_clamps_the_predicted_x0_only():
    """clip_sample (reference constructor default, gaussian_diffusion.py:146-160) clamps x0_hat to [-1, 1] before the
    posterior mean; x_t itself is not clamped."""
    sch = S.SchedulerOracle("ddpm", variance_type="fixed_large", clip_sample=True)
    t = 500
    c = {k: float(v) for k, v in sch.coefficients(t).items()}
    x = torch.full((1, 4), 5.0)
    eps = torch.zeros(1, 4)                                # x0_hat = 5 / sqrt(abar) > 1 -> clamped to 1
    got = sch.step(eps, t, x, noise=torch.zeros(1, 4))
    torch.testing.assert_close(got, torch.full((1, 4), c["x0_coeff"] * 1.0 + c["xt_coeff"] * 5.0))
    free = S.SchedulerOracle("ddpm", variance_type="fixed_large", clip_sample=False).step(eps, t, x, noise=torch.zeros(1, 4))
    assert float(free[0, 0]) > float(got[0, 0])


@pytest.mark.parametrize("betas", BETAS)
@pytest.mark.parametrize("kind,n_steps", [("ddpm", 1000), ("ddpm", 100), ("ddim", 10), ("ddim", 50)])
def test_product_tables_satisfy_the_same_closed_forms(kind, n_steps, betas):
    """The kernel's coefficient table (graspldm_b200/schedulers.py, what gldm_sampler_* consumes) against float64 closed
    forms directly - independent of the oracle."""
    T = 1000
    abar = _abar64(T, *betas)
    sch = NoiseSchedule(kind, num_train_timesteps=T, variance_type="fixed_large", beta_start=betas[0], beta_end=betas[1])
    sch.set_timesteps(n_steps)
    ts, tab = sch.table()
    tab = tab.double().numpy()
    assert ts == list(reversed(range(0, T, T // n_steps)))
    for row, t in zip(tab, ts):
        prev = t - T // n_steps
        a_t, a_p = abar[t], (abar[prev] if prev >= 0 else 1.0)
        np.testing.assert_allclose(row[0], math.sqrt(1 - a_t), rtol=_rtol(1 - a_t))
        np.testing.assert_allclose(row[1], math.sqrt(a_t), rtol=5e-5)
        if kind == "ddpm":
            alpha = a_t / a_p
            tol = max(_rtol(1 - alpha), _rtol(1 - a_t), _rtol(1 - a_p) if prev >= 0 else 0)
            np.testing.assert_allclose(row[2], math.sqrt(a_p) * (1 - alpha) / (1 - a_t), rtol=tol, atol=1e-7)
            np.testing.assert_allclose(row[3], math.sqrt(alpha) * (1 - a_p) / (1 - a_t), rtol=tol, atol=2e-6)
            np.testing.assert_allclose(row[4], math.sqrt(1 - alpha) if t > 0 else 0.0, rtol=tol, atol=2e-6)
        else:
            np.testing.assert_allclose(row[2], math.sqrt(a_p), rtol=5e-5)
            np.testing.assert_allclose(row[3], math.sqrt(1 - a_p), rtol=_rtol(1 - a_p) if prev >= 0 else 0, atol=2e-6)
            assert row[4] == 0.0
