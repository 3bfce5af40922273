"""GPU parity of the fused generation path (boundary A) in strict-fp32 mode, through the model classes and the
C ABI, against the CPU oracle on the same seeded inputs and against the committed reference fixtures.
Tolerances (fp32 SIMT vs fp32 CPU, different summation orders) are written next to each check."""
import os

import numpy as np
import pytest
import torch

import _data
import _models
from oracle import model_torch as M

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def maxerr(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


@pytest.fixture(scope="module")
def fpc(cuda):
    m = _models.build("fpc")
    vae, ddm = _models.split_state_dicts(m)
    return m.to(cuda), vae, ddm


@pytest.mark.parametrize("name", ["fpc", "ppc"])
def test_denoiser_and_decoder_forward(cuda, name):
    m = _models.build(name)
    vae, ddm = _models.split_state_dicts(m)
    m = m.to(cuda)
    g = np.load(os.path.join(G, f"dense_{name}.npz"))
    t = lambda k: torch.from_numpy(g[k])
    eps = m.diffusion_model.model(t("x").to(cuda), time=t("t").to(cuda), z_cond=t("z_cond").to(cuda))
    tm, lg = m.vae_model.decoder(t("z_h").to(cuda), t("z_cond").to(cuda))
    print(f"[{name}] denoiser max|err| {maxerr(eps.cpu(), g['eps']):.2e}  decoder tmrp {maxerr(tm.cpu(), g['tmrp']):.2e} "
          f"logit {maxerr(lg.cpu(), g['logit']):.2e}")
    np.testing.assert_allclose(eps.cpu().numpy(), g["eps"], rtol=1e-4, atol=2e-5)      # vs reference module (golden)
    np.testing.assert_allclose(tm.cpu().numpy(), g["tmrp"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(lg.cpu().numpy(), g["logit"], rtol=1e-4, atol=2e-5)
    # a batch that does not fill the last tile, every timestep class, vs the oracle live
    gen = torch.Generator().manual_seed(11)
    D, Dc = g["x"].shape[-1], g["z_cond"].shape[-1]
    B = 37
    x, zc = torch.randn(B, 1, D, generator=gen), torch.randn(B, 3, Dc, generator=gen)
    tt = torch.randint(0, 1000, (B,), generator=gen)
    with torch.no_grad():
        want = M.denoiser_forward(ddm, "diffusion_model.model.", x, tt, zc)
        wt, wl = M.decoder_forward(vae, "decoder.", x[:, 0], zc)
    got = m.diffusion_model.model(x.to(cuda), time=tt.to(cuda), z_cond=zc.to(cuda))
    gt, gl = m.vae_model.decoder(x[:, 0].to(cuda), zc.to(cuda))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gt.cpu().numpy(), wt.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gl.cpu().numpy(), wl.numpy(), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("name", ["fpc", "ppc"])
def test_encoder_forward(cuda, name):
    m = _models.build(name)
    vae, _ = _models.split_state_dicts(m)
    m = m.to(cuda)
    xyz = torch.cat([_data.synthetic_clouds(2, seed=1234, dist="S"), _data.synthetic_clouds(1, seed=99, dist="G")])
    z = m.vae_model.encode_pc(xyz.to(cuda))
    want = np.load(os.path.join(G, f"encoder_{name}.npz"))["z_pc"]
    print(f"[{name}] encoder max|err| {maxerr(z.cpu(), want):.2e} (|z| max {np.abs(want).max():.3f})")
    # 8 GFLOP of fp32 accumulation per cloud in a different order than the CPU path
    np.testing.assert_allclose(z.cpu().numpy(), want, rtol=2e-4, atol=1e-4)


@pytest.mark.parametrize("tag,kind,steps", [("ddpm10", "ddpm", 10), ("ddim5", "ddim", 5), ("ddpm100", "ddpm", 100),
                                            ("ddpmfull", "ddpm", None)])
def test_ldm_generation_matches_reference_fixture(cuda, tag, kind, steps):
    g = np.load(os.path.join(G, f"ldm_fpc_{tag}.npz"))
    m = _models.build("fpc", scheduler=kind).to(cuda)
    if steps:
        m.set_inference_timesteps(steps)
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S").to(cuda)
    (tm, lg), steps_out = m.generate_grasps(xyz, num_grasps=3, x_T=torch.from_numpy(g["x_T"]).to(cuda),
                                            noise=torch.from_numpy(g["noise"]).to(cuda))
    assert steps_out == []
    print(f"[{tag}] tmrp max|err| {maxerr(tm.cpu(), g['tmrp']):.2e} logit {maxerr(lg.cpu(), g['logit']):.2e}")
    # fp32 end to end; the recurrence damps eps errors (x0 coefficient ~2e-2 per step), clamp flips aside
    np.testing.assert_allclose(tm.cpu().numpy(), g["tmrp"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(lg.cpu().numpy(), g["logit"], rtol=1e-3, atol=1e-3)


def test_sampler_trajectory_and_return_all(fpc, cuda):
    m, vae, ddm = fpc
    m.set_inference_timesteps(20)
    gen = torch.Generator().manual_seed(3)
    n_obj, G_ = 3, 5
    z = torch.randn(n_obj, 3, 64, generator=gen)
    x_T = torch.randn(n_obj * G_, 1, 4, generator=gen)
    noise = torch.randn(20, n_obj * G_, 1, 4, generator=gen)
    x0, allx = m.diffusion_model.sample(z_cond=z.to(cuda), batch_size=n_obj * G_, return_all=True, x_T=x_T.to(cuda),
                                        noise=noise.to(cuda), grasps_per_object=G_)
    assert len(allx) == 21 and torch.equal(allx[0].cpu(), x_T) and torch.equal(allx[-1], x0)
    with torch.no_grad():
        want, wall = M.ldm_sample(ddm, z.repeat_interleave(G_, 0), x_T, noise=noise, num_inference_steps=20, return_all=True)
    for i in (1, 5, 20):
        np.testing.assert_allclose(allx[i].cpu().numpy(), wall[i].numpy(), rtol=1e-3, atol=5e-4)
    # reference call convention: z_cond already repeated per grasp, grasps_per_object defaulted
    x0b, _ = m.diffusion_model.sample(z_cond=z.repeat_interleave(G_, 0).to(cuda), batch_size=n_obj * G_,
                                      x_T=x_T.to(cuda), noise=noise.to(cuda))
    assert torch.equal(x0, x0b)
    m.diffusion_model.noise_scheduler.num_inference_steps = None


def test_vae_mode_and_pose_postprocessing(fpc, cuda):
    m, _, _ = fpc
    g = np.load(os.path.join(G, "vae_fpc.npz"))
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S")
    from graspldm_b200.inference import InferenceVAE, default_metas
    inf = InferenceVAE(m.vae_model, device=cuda)
    res = inf.generate_grasps(xyz, default_metas(2), num_grasps=3, z_h=torch.from_numpy(g["z_h"]).to(cuda))
    np.testing.assert_allclose(res["grasp_tmrp"].cpu().numpy(), g["grasp_tmrp"], rtol=1e-3, atol=1e-4)
    assert res["grasps"].shape == (2, 3, 4, 4) and res["confidence"].shape == (2, 3, 1) and res["pc"].shape == (2, 1024, 3)
    # pose kernel against the reference's tmrp_to_H on the reference's own poses (pure fp32, same op order)
    from graspldm_b200.rotations import tmrp_to_H
    H = tmrp_to_H(torch.from_numpy(g["grasp_tmrp"]).to(cuda))
    np.testing.assert_allclose(H.cpu().numpy(), g["H"], rtol=1e-6, atol=1e-7)
    R = H[..., :3, :3]
    eye = torch.eye(3, device=cuda).expand_as(R)
    torch.testing.assert_close(R @ R.transpose(-1, -2), eye, rtol=0, atol=1e-5)     # proper rotations
    # same seed -> same grasps through the reference-style RNG (z_h drawn on the CPU generator)
    torch.manual_seed(5)
    a = m.vae_model.generate_grasps(xyz.to(cuda), 3)[0]
    np.testing.assert_allclose(a.cpu().numpy(), g["tmrp"], rtol=1e-3, atol=1e-4)


def test_full_size_properties_config2(fpc, cuda):
    """BASELINE config 2 (64 objects x 20 grasps, 100 DDPM steps): size-independent properties."""
    from graspldm_b200.inference import InferenceLDM, default_metas
    m, _, _ = fpc
    m.set_inference_timesteps(100)
    m.diffusion_model.rng_mode = "fused"
    inf = InferenceLDM(m, device=cuda)
    pcs = _data.synthetic_clouds(64, seed=1234, dist="S")
    torch.manual_seed(0)      # x_T comes from the global CPU generator, as in the reference
    a = inf.generate_grasps(pcs, default_metas(64), num_grasps=20, seed=7)
    torch.manual_seed(0)
    b = inf.generate_grasps(pcs, default_metas(64), num_grasps=20, seed=7)
    assert all(torch.equal(a[k], b[k]) for k in ("grasps", "grasp_tmrp", "confidence"))     # deterministic
    assert torch.isfinite(a["grasps"]).all() and a["grasps"].shape == (64, 20, 4, 4)
    assert (a["confidence"] > 0).all() and (a["confidence"] < 1).all()
    # objects are independent: generating a slice alone gives the same grasps for those objects
    x_T = torch.randn(64 * 20, 1, 4, generator=torch.Generator().manual_seed(1)).to(cuda)
    full = inf.generate_grasps(pcs, default_metas(64), num_grasps=20, seed=7, x_T=x_T)
    part = inf.generate_grasps(pcs[16:24], default_metas(8), num_grasps=20, seed=7, x_T=x_T[16 * 20:24 * 20])
    # (the fused RNG is keyed by the sample index within the call, so compare a DDIM-free quantity: the encoder)
    z_full = m.vae_model.encode_pc(pcs.to(cuda))
    z_part = m.vae_model.encode_pc(pcs[16:24].to(cuda))
    assert torch.equal(z_full[16:24], z_part)
    assert full["grasps"].shape[0] == 64 and part["grasps"].shape[0] == 8
    torch.manual_seed(0)
    c = inf.generate_grasps(pcs, default_metas(64), num_grasps=20, seed=8)
    assert not torch.equal(a["grasp_tmrp"], c["grasp_tmrp"])                                # the seed matters
    m.diffusion_model.rng_mode = "reference"
    m.diffusion_model.noise_scheduler.num_inference_steps = None


def test_fused_rng_is_standard_normal(fpc, cuda):
    """DDPM with eps weights zeroed is not available; instead sample the in-kernel Philox stream through a
    1-step run: x_prev - mu = sigma * z  ->  z recovered from two runs that differ only in the noise source."""
    m, _, _ = fpc
    m.set_inference_timesteps(2)          # t = 500, 0
    n = 20000
    z = torch.zeros(1, 3, 64, device=cuda)
    x_T = torch.zeros(n, 1, 4, device=cuda)
    zero = torch.zeros(2, n, 1, 4, device=cuda)
    x_a, all_a = m.diffusion_model.sample(z_cond=z, batch_size=n, x_T=x_T, noise=zero, grasps_per_object=n, return_all=True)
    m.diffusion_model.rng_mode = "fused"
    x_b, all_b = m.diffusion_model.sample(z_cond=z, batch_size=n, x_T=x_T, grasps_per_object=n, seed=123, return_all=True)
    _, coef = m.diffusion_model.noise_scheduler.table()
    zz = ((all_b[1] - all_a[1]) / coef[0, 4].item()).flatten().double().cpu()
    assert abs(zz.mean().item()) < 0.02 and abs(zz.std().item() - 1.0) < 0.02
    assert abs((zz ** 3).mean().item()) < 0.05 and abs((zz ** 4).mean().item() - 3.0) < 0.15
    m.diffusion_model.rng_mode = "reference"
    m.diffusion_model.noise_scheduler.num_inference_steps = None


def test_normalize_input_and_per_object_postprocessing(fpc, cuda):
    """raw clouds in, world-frame grasps out (inference_base.py:161-212, tools/inference.py:570-666)."""
    from graspldm_b200.inference import InferenceLDM, InferenceVAE
    from graspldm_b200 import engine
    m, vae_sd, _ = fpc
    g = np.load(os.path.join(G, "normalize_input.npz"))
    norm = dict(pc_shift=g["pc_shift"].tolist(), grasp_shift=g["grasp_shift"].tolist(),
                translation_scale=float(g["translation_scale"]), rotation_scale=float(g["rotation_scale"]))
    inf = InferenceVAE(m.vae_model, device=cuda, norm_config=norm)
    raw = torch.from_numpy(g["raw"])
    for tag, pc in (("single", raw[1]), ("batch", raw)):
        before = pc.clone()
        pcn, metas = inf.normalize_input(pc)
        assert torch.equal(pc, before)                                        # caller's tensor untouched
        # the cloud mean is an fp64 reduction here, an fp32 cascade sum in torch.mean: 1 ulp of the mean (~1e-7 m)
        np.testing.assert_allclose(pcn.cpu().numpy(), g[f"{tag}_pc"], rtol=0, atol=2e-5)
        for k in ("pc_mean", "pc_std", "grasp_mean", "grasp_std"):
            assert tuple(metas[k].shape) == g[f"{tag}_{k}"].shape, k
            np.testing.assert_allclose(metas[k].cpu().numpy(), g[f"{tag}_{k}"], rtol=0, atol=3e-7)
    # per-object un-normalisation + pose kernel on the reference's statistics: exact fp32
    tm = torch.from_numpy(g["tmrp"]).to(cuda).reshape(12, 6)
    gt, H, _ = engine.pose_postprocess(tm, torch.zeros(12, 1, device=cuda), torch.from_numpy(g["batch_grasp_mean"]),
                                       torch.from_numpy(g["batch_grasp_std"]), grasps_per_obj=4)
    np.testing.assert_array_equal(gt.cpu().numpy().reshape(3, 4, 6), g["batch_grasp_tmrp"])
    np.testing.assert_allclose(H.cpu().numpy().reshape(3, 4, 4, 4), g["batch_H"], rtol=1e-6, atol=1e-7)
    with pytest.raises(RuntimeError):
        engine.pose_postprocess(tm, torch.zeros(12, 1, device=cuda), torch.zeros(2, 6), torch.ones(1, 6), grasps_per_obj=4)
    # raw clouds end to end: same grasps as normalise-by-hand + generate_grasps, translated by each cloud's mean
    z_h = torch.randn(12, 4, generator=torch.Generator().manual_seed(3)).to(cuda)
    res = inf.generate_on_pointcloud(raw, num_grasps=4, z_h=z_h)
    pcn, metas = M.normalize_input(raw, torch.from_numpy(g["pc_shift"]), torch.ones(3) * norm["translation_scale"],
                                   torch.from_numpy(g["grasp_shift"]),
                                   torch.cat((torch.ones(3) * norm["translation_scale"], torch.ones(3) * norm["rotation_scale"])))
    with torch.no_grad():
        tmw, lgw = M.generate_grasps_vae(vae_sd, pcn, 4, z_h.cpu())
    want = M.postprocess(tmw, lgw, pcn, metas, 3, 4)
    np.testing.assert_allclose(res["grasp_tmrp"].cpu().numpy(), want["grasp_tmrp"].numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(res["grasps"].cpu().numpy(), want["grasps"].numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(res["pc"].cpu().numpy(), g["raw"], rtol=0, atol=1e-6)
    # LDM wrapper: same entry point, per-object statistics, intermediate steps for a single cloud
    m.set_inference_timesteps(10)
    ldm = InferenceLDM(m, device=cuda, norm_config=norm)
    out = ldm.infer_on_pointcloud(raw[0], num_grasps=4, return_intermediate=True)
    assert out["grasps"].shape == (1, 4, 4, 4) and len(out["all_steps_grasps"]) == 50
    torch.testing.assert_close(out["all_steps_grasps"][-1], out["grasps"][0], rtol=0, atol=0)
    centre = out["grasps"][0, :, :3, 3].mean(0).cpu()
    assert (centre - raw[0].mean(0)).abs().max() < 0.5                        # grasps live around the raw cloud



def test_elucidated_samplers(fpc, cuda):
    """ElucidatedDiffusion.sample (stochastic Heun and DPM-Solver++ 2M) around the denoiser kernels,
    against the fixture produced by the reference's own class (elucidated_diffusion.py:126-315)."""
    from graspldm_b200.edm import ElucidatedDiffusion
    from graspldm_b200.grasp_ldm import GraspLatentDDM
    m, vae_sd, _ = fpc
    g = np.load(os.path.join(G, "edm_fpc.npz"))
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    edm = ElucidatedDiffusion(net=m.diffusion_model.model, seq_length=4)
    torch.testing.assert_close(edm.sample_schedule(8).cpu()[[0, 7, 8]], torch.tensor([80.0, 0.002, 0.0]), rtol=1e-5, atol=0)
    # the whole sampler (all evaluations, churn noise, Heun / multistep updates) is ONE launch of the persistent kernel.
    # bf16: a single preconditioned evaluation is within the usual 5e-2; over the 15 evaluations of the 8-step Heun sampler the
    # errors add up (the elucidated update has no contraction like the DDPM posterior mean), hence the wider bound there
    for prec, tol, tol_s in (("fp32", dict(rtol=1e-3, atol=2e-4), dict(rtol=1e-3, atol=2e-4)),
                             ("bf16", dict(rtol=5e-2, atol=5e-2), dict(rtol=1.5e-1, atol=1.5e-1))):
        for k, sg in enumerate((80.0, 2.5, 0.05)):
            got = edm.preconditioned_network_forward(t("denoise_x") * sg, sg, z_cond=t("z_cond"), precision=prec)
            np.testing.assert_allclose(got.cpu().numpy(), g[f"denoise_{k}"], **tol)
        x, allx = edm.sample(use_dpmpp=False, batch_size=6, z_cond=t("z_cond"), num_sample_steps=int(g["heun_steps"]),
                             return_all=True, x_init=t("heun_x_init"), noise=t("heun_noise"), precision=prec)
        assert len(allx) == int(g["heun_steps"]) + 1
        print(f"[edm heun {prec}] max|err| {np.abs(x.cpu().numpy() - g['heun_x']).max():.3e}")
        np.testing.assert_allclose(x.cpu().numpy(), g["heun_x"], **tol_s)
        x, _ = edm.sample(use_dpmpp=True, batch_size=6, z_cond=t("z_cond"), num_sample_steps=int(g["dpmpp_steps"]),
                          x_init=t("dpmpp_x_init"), precision=prec)
        print(f"[edm dpm++ {prec}] max|err| {np.abs(x.cpu().numpy() - g['dpmpp_x']).max():.3e}")
        np.testing.assert_allclose(x.cpu().numpy(), g["dpmpp_x"], **tol_s)
    # the model-level switch of the reference (grasp_ldm.py:59-62, 214-219): elucidated_diffusion=True + use_dpmpp kwarg
    ldm = GraspLatentDDM(model=m.diffusion_model.model, latent_in_features=4, diffusion_timesteps=1000, diffusion_loss="l2",
                         elucidated_diffusion=True)
    ldm.set_vae_model(m.vae_model)
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S").to(cuda)
    (tm, lg), steps = ldm.generate_grasps(xyz, num_grasps=3, use_dpmpp=True, num_sample_steps=10)
    assert tm.shape == (6, 6) and lg.shape == (6, 1) and steps == [] and torch.isfinite(tm).all()
    with pytest.raises(KeyError):
        ldm.generate_grasps(xyz, num_grasps=3)        # the reference pops `use_dpmpp` unconditionally (:171)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ppc_ldm_generation_matches_reference_fixture(cuda, precision):
    """The partial-point-cloud model family end to end (encoder -> 10 DDPM steps on the 16-position latent -> decoder)
    against the fixture of the unmodified reference classes; bf16 = every tensor-core kernel switched on."""
    g = np.load(os.path.join(G, "ldm_ppc_ddpm10.npz"))
    m = _models.build("ppc").to(cuda)
    m.set_inference_timesteps(10)
    m.diffusion_model.precision = precision
    m.vae_model.encoder.pc_encoder.precision = precision
    m.vae_model.decoder.precision = precision
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S").to(cuda)
    (tm, lg), _ = m.generate_grasps(xyz, num_grasps=3, x_T=torch.from_numpy(g["x_T"]).to(cuda),
                                    noise=torch.from_numpy(g["noise"]).to(cuda))
    print(f"[ppc {precision}] tmrp max|err| {maxerr(tm.cpu(), g['tmrp']):.2e} logit {maxerr(lg.cpu(), g['logit']):.2e}")
    tol = dict(rtol=1e-3, atol=1e-3) if precision == "fp32" else dict(rtol=3e-2, atol=3e-2)
    np.testing.assert_allclose(tm.cpu().numpy(), g["tmrp"], **tol)
    np.testing.assert_allclose(lg.cpu().numpy(), g["logit"], **tol)


def _class_conditioned(cuda):
    from graspldm_b200 import configs
    from graspldm_b200.gaussian_diffusion import GaussianDiffusion1D
    from graspldm_b200.resnets import ClassTimeConditionedResNet1D
    torch.manual_seed(0)
    den = _models.trained_like_(ClassTimeConditionedResNet1D(**configs.model_config("fpc")["denoiser"]), 4).eval()
    gd = GaussianDiffusion1D(model=den, n_dims=4, num_steps=1000, loss_type="l2", beta_schedule="linear", beta_start=5e-5,
                             beta_end=1e-3, noise_scheduler_type="ddpm", variance_type="fixed_large").eval()
    return den, gd


def test_class_conditioned_denoiser(cuda):
    """ClassTimeConditionedResNet1D (class_conditioned_resnet.py:9-122) - single evaluations and a 10-step
    DDPM run with metas["mode_cls"] - against the fixture of the reference class; the class embedding is added to the time
    embedding inside all three sampler kernels."""
    from graspldm_b200 import _lib
    g = np.load(os.path.join(G, "cls_fpc.npz"))
    den, gd = _class_conditioned(cuda)
    np.testing.assert_array_equal(den.cls_embed[0].weight.detach().numpy(), g["cls_w"])        # same seeded weights as the reference
    den, gd = den.to(cuda), gd.to(cuda)
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    sd = {"m." + k: v.detach().cpu() for k, v in den.state_dict().items()}
    with torch.no_grad():
        want = M.denoiser_forward(sd, "m.", t("x").cpu(), t("t").cpu(), t("z_cond").cpu(), cls_cond=t("cls").cpu())
    np.testing.assert_allclose(want.numpy(), g["eps"], rtol=1e-5, atol=2e-6)                   # oracle pinned by the reference
    gd.set_inference_timesteps(10)
    for label, prec, rows, tol in (("fp32", "fp32", -1, dict(rtol=1e-4, atol=3e-5)), ("bf16 channel-major", "bf16", 0, dict(rtol=5e-2, atol=5e-2)),
                                   ("bf16 row-major", "bf16", 1, dict(rtol=5e-2, atol=5e-2))):
        _lib.call("gldm_sampler_tc_set_rows", rows)
        try:
            eps = den(t("x"), time=t("t"), z_cond=t("z_cond"), cls_cond=t("cls"), precision=prec)
            eps_m = den(t("x"), time=t("t"), z_cond=t("z_cond"), metas=dict(mode_cls=t("cls").view(-1)), precision=prec)
            eps_0 = den(t("x"), time=t("t"), z_cond=t("z_cond"), cls_cond=torch.zeros_like(t("cls")), precision=prec)
            x0, _ = gd.sample(z_cond=t("z_cond"), batch_size=6, x_T=t("x_T"), noise=t("noise"), metas=dict(mode_cls=t("cls").view(-1)),
                              precision=prec)
        finally:
            _lib.call("gldm_sampler_tc_set_rows", -1)
        print(f"[class-conditioned, {label}] eps max|err| {np.abs(eps.cpu().numpy() - g['eps']).max():.2e}, "
              f"x0 after 10 DDPM steps {np.abs(x0.cpu().numpy() - g['x0']).max():.2e}")
        np.testing.assert_allclose(eps.cpu().numpy(), g["eps"], **tol)
        assert torch.equal(eps, eps_m)
        assert (eps - eps_0).abs().max() > 1e-3                                     # the class does matter
        np.testing.assert_allclose(x0.cpu().numpy(), g["x0"], rtol=max(tol["rtol"], 1e-3), atol=max(tol["atol"], 1e-3))
    with pytest.raises(AssertionError):
        den(t("x"), time=t("t"), z_cond=t("z_cond"))                               # the reference asserts on a missing class
