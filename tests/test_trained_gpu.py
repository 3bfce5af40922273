"""GPU parity on CHECKPOINT-LIKE state.  A seeded random init leaves BatchNorm running statistics at
0 / 1 and every BatchNorm / GroupNorm / LayerNorm affine parameter at 1 / 0, so a kernel that ignored them - or a wrong
BatchNorm fold - would pass every other fixture.  Here every norm parameter and buffer is randomised
(tests/_models.py::trained_like_), the state is written as a Lightning-style checkpoint, loaded through
`inference.load_checkpoint` (tools/inference.py:514-566) into a freshly built model, and every kernel family
(strict fp32, tcgen05 channel-major, tcgen05 row-major) is compared with fixtures produced by the UNMODIFIED reference
classes holding the same state (tests/golden/make_golden.py::trained_golden)."""
import os

import numpy as np
import pytest
import torch

import _data
import _models
from oracle import model_torch as M

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _rot_angle_deg(Ra, Rb):
    R = Ra.transpose(-1, -2) @ Rb
    c = ((R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]) - 1.0) / 2.0
    return torch.rad2deg(torch.acos(c.clamp(-1.0, 1.0)))


def _set_precision(model, prec):
    model.diffusion_model.precision = prec
    model.vae_model.encoder.pc_encoder.precision = prec
    model.vae_model.decoder.precision = prec


def _from_checkpoint(tmp_path_factory, name, scheduler, cuda):
    """trained-like state -> Lightning .ckpt (EMA copy under `ema_model.online_model.`, a poisoned raw copy under
    `model.`) -> load_checkpoint into a default-initialised model."""
    from graspldm_b200.inference import load_checkpoint
    src = _models.build_trained_like(name, scheduler)
    sd = {}
    for k, v in src.state_dict().items():
        sd["ema_model.online_model." + k] = v.clone()
        sd["model." + k] = torch.zeros_like(v)
    path = os.path.join(str(tmp_path_factory.mktemp("ckpt")), f"{name}_{scheduler}.ckpt")
    torch.save({"state_dict": sd, "epoch": 3}, path)
    dst = _models.build(name, scheduler, seed=17)
    load_checkpoint(dst, path, use_ema_model=True)
    return dst.to(cuda)


@pytest.fixture(scope="module", params=["fpc", "ppc"])
def trained(request, tmp_path_factory, cuda):
    return request.param, _from_checkpoint(tmp_path_factory, request.param, "ddpm", cuda)


def _kernel_modes(name):
    """(label, precision, rows flag): rows = 1 forces the row-major tcgen05 sampler (fpc latent only), 0 the channel-major."""
    modes = [("fp32", "fp32", -1), ("bf16 channel-major", "bf16", 0)]
    if name == "fpc":
        modes.append(("bf16 row-major", "bf16", 1))
    return modes


def test_single_evaluations_on_trained_like_state(trained, cuda):
    """denoiser (resnets.py:558-616: GroupNorm weight / bias, LayerNorm.g), decoder (grasp_vae.py:401-436) and encoder
    (shared_mlp.py:18-28 BatchNorm fold with running statistics; pvconv.py:48-73 GroupNorm affine) vs the reference."""
    from graspldm_b200 import _lib
    name, m = trained
    g0 = np.load(os.path.join(G, f"dense_{name}.npz"))
    g = np.load(os.path.join(G, f"dense_{name}_trained.npz"))
    t = lambda k: torch.from_numpy(g0[k]).to(cuda)
    xyz = torch.cat([_data.synthetic_clouds(2, seed=1234, dist="S"), _data.synthetic_clouds(1, seed=99, dist="G")]).to(cuda)
    zw = np.load(os.path.join(G, f"encoder_{name}_trained.npz"))["z_pc"]
    for label, prec, rows in _kernel_modes(name):
        _lib.call("gldm_sampler_tc_set_rows", rows)
        try:
            _set_precision(m, prec)
            eps = m.diffusion_model.model(t("x"), time=t("t"), z_cond=t("z_cond"), precision=prec).cpu().numpy()
            tm, lg = m.vae_model.decoder(t("z_h"), t("z_cond"))
            z = m.vae_model.encode_pc(xyz).cpu().numpy()
        finally:
            _lib.call("gldm_sampler_tc_set_rows", -1)
            _set_precision(m, "fp32")
        print(f"[{name} trained-like, {label}] denoiser max|err| {np.abs(eps - g['eps']).max():.2e} (max|eps| {np.abs(g['eps']).max():.2f}), "
              f"decoder tmrp {np.abs(tm.cpu().numpy() - g['tmrp']).max():.2e}, encoder {np.abs(z - zw).max():.2e} (max|z| {np.abs(zw).max():.2f})")
        if prec == "fp32":
            tol, tol_z = dict(rtol=1e-4, atol=3e-5), dict(rtol=2e-4, atol=1e-4)
        else:       # stated bf16 tolerances of the tensor-core path
            tol, tol_z = dict(rtol=5e-2, atol=5e-2), dict(rtol=1e-2, atol=4e-3)
        np.testing.assert_allclose(eps, g["eps"], **tol)
        np.testing.assert_allclose(tm.cpu().numpy(), g["tmrp"], **tol)
        np.testing.assert_allclose(lg.cpu().numpy(), g["logit"], **tol)
        np.testing.assert_allclose(z, zw, **tol_z)


def test_vae_mode_on_trained_like_state(trained, cuda):
    name, m = trained
    g = np.load(os.path.join(G, f"vae_{name}_trained.npz"))
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S").to(cuda)
    for prec, tol in (("fp32", dict(rtol=1e-3, atol=2e-4)), ("bf16", dict(rtol=5e-2, atol=5e-2))):
        _set_precision(m, prec)
        try:
            tm, lg = m.vae_model.generate_grasps(xyz, 3, z_h=torch.from_numpy(g["z_h"]).to(cuda))
        finally:
            _set_precision(m, "fp32")
        np.testing.assert_allclose(tm.cpu().numpy(), g["tmrp"], **tol)
        np.testing.assert_allclose(lg.cpu().numpy(), g["logit"], **tol)


@pytest.mark.parametrize("name,kind,steps", [("fpc", "ddpm", 100), ("fpc", "ddim", 10), ("ppc", "ddpm", 100), ("ppc", "ddim", 10)])
def test_ldm_generation_on_trained_like_state(tmp_path_factory, cuda, name, kind, steps):
    """encode_pc -> T-step sampler -> decoder through GraspLatentDDM.generate_grasps on every kernel family."""
    from graspldm_b200 import _lib
    from graspldm_b200.inference import InferenceLDM, default_metas
    g = np.load(os.path.join(G, f"ldm_{name}_trained_{kind}{steps}.npz"))
    m = _from_checkpoint(tmp_path_factory, name, kind, cuda)
    m.set_inference_timesteps(steps)
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S")
    want = M.postprocess(torch.from_numpy(g["tmrp"]), torch.from_numpy(g["logit"]), xyz, default_metas(2), 2, 3)
    inf = InferenceLDM(m, device=cuda)
    for label, prec, rows in _kernel_modes(name):
        _lib.call("gldm_sampler_tc_set_rows", rows)
        try:
            _set_precision(m, prec)
            out = inf.generate_grasps(xyz, default_metas(2), num_grasps=3, x_T=torch.from_numpy(g["x_T"]).to(cuda),
                                      noise=torch.from_numpy(g["noise"]).to(cuda))
        finally:
            _lib.call("gldm_sampler_tc_set_rows", -1)
        H, Hw = out["grasps"].cpu(), want["grasps"]
        dt = (H[..., :3, 3] - Hw[..., :3, 3]).norm(dim=-1).max().item()
        da = _rot_angle_deg(H[..., :3, :3], Hw[..., :3, :3]).max().item()
        dm = (out["grasp_tmrp"].cpu() - want["grasp_tmrp"]).abs().max().item()
        print(f"[{name} {kind}{steps} trained-like, {label}] translation {dt * 1e3:.3f} mm, rotation {da:.3f} deg, max|tmrp err| {dm:.2e}")
        if prec == "fp32":
            np.testing.assert_allclose(out["grasp_tmrp"].cpu().numpy().reshape(-1, 6) / np.array([[.05] * 3 + [.5] * 3]),
                                       g["tmrp"], rtol=1e-3, atol=1e-3)
            assert dt < 3e-4 and da < 0.3
        else:       # stated bf16 tolerance of the path: 1 mm / 2 degrees after un-normalisation
            assert dt < 1e-3 and da < 2.0
            np.testing.assert_allclose(out["confidence"].cpu().numpy(), want["confidence"].numpy(), atol=3e-2)


def test_config2_size_rows_kernel_four_streams_vs_reference(tmp_path_factory, cuda):
    """BASELINE config 2 at full size the way bench.py runs it: 64 objects x 20 grasps, 100 DDPM steps, the row-major
    tcgen05 sampler forced, four generation calls in flight on four streams - every one of them against the fixture of
    the unmodified reference classes (pre-drawn x_T / noise re-drawn here from the same seeded CPU generators)."""
    from graspldm_b200 import _lib
    from graspldm_b200.inference import InferenceLDM, default_metas
    g = np.load(os.path.join(G, "ldm_fpc_trained_config2.npz"))
    m = _from_checkpoint(tmp_path_factory, "fpc", "ddpm", cuda)
    m.set_inference_timesteps(100)
    n_obj, G_ = 64, 20
    noise = torch.randn(100, n_obj * G_, 1, 4, generator=torch.Generator().manual_seed(42)).to(cuda)
    torch.manual_seed(42)
    x_T = torch.randn((n_obj * G_, 1, 4)).to(cuda)
    xyz = _data.synthetic_clouds(n_obj, seed=1234, dist="S")
    want = M.postprocess(torch.from_numpy(g["tmrp"]), torch.from_numpy(g["logit"]), xyz, default_metas(n_obj), n_obj, G_)
    inf = InferenceLDM(m, device=cuda)
    _set_precision(m, "bf16")
    xyz_dev = xyz.to(cuda)
    metas = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in default_metas(n_obj).items()}
    inf.generate_grasps(xyz_dev, metas, num_grasps=G_, x_T=x_T, noise=noise)        # weight packing outside the streams
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=cuda) for _ in range(4)]
    outs = []
    _lib.call("gldm_sampler_tc_set_rows", 1)
    try:
        for rep in range(8):
            with torch.cuda.stream(streams[rep % 4]):
                outs.append(inf.generate_grasps(xyz_dev, metas, num_grasps=G_, x_T=x_T, noise=noise))
        torch.cuda.synchronize()
    finally:
        _lib.call("gldm_sampler_tc_set_rows", -1)
    Hw = want["grasps"]
    for i, out in enumerate(outs):
        H = out["grasps"].cpu()
        dt = (H[..., :3, 3] - Hw[..., :3, 3]).norm(dim=-1)
        da = _rot_angle_deg(H[..., :3, :3], Hw[..., :3, :3])
        if i == 0:
            print(f"[config 2, rows kernel, 4 streams vs reference] translation max {dt.max().item() * 1e3:.3f} mm "
                  f"(mean {dt.mean().item() * 1e3:.3f}), rotation max {da.max().item():.3f} deg (mean {da.mean().item():.3f})")
        assert dt.max().item() < 1e-3 and da.max().item() < 2.0
        assert torch.equal(out["grasps"], outs[0]["grasps"])                           # concurrent calls do not interfere
    # the strict-fp32 path on the same inputs, for the record of what bf16 costs
    _set_precision(m, "fp32")
    f = inf.generate_grasps(xyz_dev, metas, num_grasps=G_, x_T=x_T, noise=noise)["grasps"].cpu()
    dt = (f[..., :3, 3] - Hw[..., :3, 3]).norm(dim=-1).max().item()
    da = _rot_angle_deg(f[..., :3, :3], Hw[..., :3, :3]).max().item()
    print(f"[config 2, fp32 path vs reference] translation max {dt * 1e3:.4f} mm, rotation max {da:.4f} deg")
    assert dt < 3e-4 and da < 0.3
