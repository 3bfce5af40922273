"""GPU parity of the tensor-core (tcgen05, bf16 operands / fp32 accumulation) sampler against the CPU oracle and
the strict-fp32 CUDA path.  Tolerances are the stated bf16 tolerances of the path."""
import os

import numpy as np
import pytest
import torch

import _data
import _models
from oracle import model_torch as M

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def fpc(cuda):
    m = _models.build("fpc")
    vae, ddm = _models.split_state_dicts(m)
    return m.to(cuda), vae, ddm


@pytest.mark.parametrize("B", [1, 16, 37])
def test_denoiser_forward_bf16(fpc, cuda, B):
    m, vae, ddm = fpc
    gen = torch.Generator().manual_seed(11 + B)
    x, zc = torch.randn(B, 1, 4, generator=gen), torch.randn(B, 3, 64, generator=gen)
    tt = torch.randint(0, 1000, (B,), generator=gen)
    with torch.no_grad():
        want = M.denoiser_forward(ddm, "diffusion_model.model.", x, tt, zc)
    got = m.diffusion_model.model(x.to(cuda), time=tt.to(cuda), z_cond=zc.to(cuda), precision="bf16").cpu()
    f32 = m.diffusion_model.model(x.to(cuda), time=tt.to(cuda), z_cond=zc.to(cuda)).cpu()
    err = (got - want).abs().max().item()
    print(f"[B={B}] bf16 tensor-core denoiser: max|err| vs oracle {err:.3e}, vs fp32 kernel {(got - f32).abs().max().item():.3e}, "
          f"max|eps| {want.abs().max().item():.3f}")
    # bf16 operands (8-bit mantissa) through ~40 GEMM layers with fp32 accumulation and fp32 normalisations
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=5e-2, atol=5e-2)


@pytest.mark.parametrize("kind,steps", [("ddpm", 10), ("ddim", 5), ("ddpm", 100)])
def test_sampler_bf16_tracks_fp32(cuda, kind, steps):
    m = _models.build("fpc", scheduler=kind).to(cuda)
    m.set_inference_timesteps(steps)
    gen = torch.Generator().manual_seed(5)
    n_obj, G_ = 3, 7
    z = torch.randn(n_obj, 3, 64, generator=gen).to(cuda)
    x_T = torch.randn(n_obj * G_, 1, 4, generator=gen).to(cuda)
    noise = torch.randn(steps, n_obj * G_, 1, 4, generator=gen).to(cuda)
    a, alla = m.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, return_all=True, x_T=x_T, noise=noise,
                                       grasps_per_object=G_)
    b, allb = m.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, return_all=True, x_T=x_T, noise=noise,
                                       grasps_per_object=G_, precision="bf16")
    assert len(allb) == steps + 1 and torch.equal(allb[0], x_T) and torch.equal(allb[-1], b)
    err = (a - b).abs().max().item()
    print(f"[{kind}{steps}] bf16 vs fp32 latents after {steps} steps: max|diff| {err:.3e}")
    # the recurrence damps eps errors (x0 coefficient of the posterior mean is O(1e-2) per step)
    np.testing.assert_allclose(b.cpu().numpy(), a.cpu().numpy(), rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("name", ["fpc", "ppc"])
def test_encoder_bf16_pointwise_layers(cuda, name):
    """Point-wise tail (96->768->1536->768) on the tensor cores vs the committed reference fixture."""
    m = _models.build(name).to(cuda)
    enc = m.vae_model.encoder.pc_encoder
    xyz = torch.cat([_data.synthetic_clouds(2, seed=1234, dist="S"), _data.synthetic_clouds(1, seed=99, dist="G")]).to(cuda)
    z32 = m.vae_model.encode_pc(xyz)
    enc.precision = "bf16"
    try:
        z16 = m.vae_model.encode_pc(xyz)
    finally:
        enc.precision = "fp32"
    want = np.load(os.path.join(G, f"encoder_{name}.npz"))["z_pc"]
    err = np.abs(z16.cpu().numpy() - want).max()
    print(f"[{name}] bf16 encoder tail: max|err| vs reference {err:.3e}, vs fp32 path {(z16 - z32).abs().max().item():.3e}, "
          f"max|z| {np.abs(want).max():.3f}")
    # three bf16 GEMM layers (K up to 1536) with fp32 accumulation, then a 1024-term fp32 reduction
    np.testing.assert_allclose(z16.cpu().numpy(), want, rtol=2e-2, atol=2e-2)
