"""CPU: pins the oracle (oracle/*.py) to fixtures produced by the reference itself.

  * tests/golden/{dense,encoder,ldm,vae}_*.npz - reference PyTorch modules run unmodified on CPU
    (tests/golden/make_golden.py); the functional restatement oracle/model_torch.py must reproduce them.
  * tests/golden/ref_ops_gpu.npz - the reference's own CUDA kernels run on a B200
    (tests/golden/make_golden_gpu.py); the numpy restatement oracle/ops_np.py must reproduce them.
  * tests/golden/state_dict_manifest.json - key names, shapes and value hashes of the reference models built
    from torch.manual_seed(0); the product's module tree must reproduce them exactly.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import _data
import _models
from oracle import model_torch as M
from oracle import ops_np
from oracle.schedulers import SchedulerOracle

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def fpc():
    m = _models.build("fpc")
    return (m,) + _models.split_state_dicts(m)


@pytest.mark.parametrize("name", ["fpc", "ppc"])
def test_state_dict_manifest(name):
    man = json.load(open(os.path.join(G, "state_dict_manifest.json")))[name]
    sd = _models.build(name).state_dict()
    assert set(sd) == set(man)
    for k, v in sd.items():
        a = v.detach().contiguous().numpy()
        assert list(a.shape) == man[k]["shape"], k
        assert hashlib.sha256(a.tobytes()).hexdigest()[:16] == man[k]["sha"], f"{k}: seeded init differs from the reference"


@pytest.mark.parametrize("name", ["fpc", "ppc"])
def test_oracle_denoiser_decoder_vs_reference(name):
    vae, ddm = _models.split_state_dicts(_models.build(name))
    g = np.load(os.path.join(G, f"dense_{name}.npz"))
    t = lambda k: torch.from_numpy(g[k])
    with torch.no_grad():
        eps = M.denoiser_forward(ddm, "diffusion_model.model.", t("x"), t("t"), t("z_cond"))
        tm, lg = M.decoder_forward(vae, "decoder.", t("z_h"), t("z_cond"))
    np.testing.assert_allclose(eps.numpy(), g["eps"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tm.numpy(), g["tmrp"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lg.numpy(), g["logit"], rtol=1e-5, atol=1e-6)


def test_oracle_encoder_vs_reference(fpc):
    _, vae, _ = fpc
    xyz = torch.cat([_data.synthetic_clouds(2, seed=1234, dist="S"), _data.synthetic_clouds(1, seed=99, dist="G")])
    with torch.no_grad():
        z = M.pvcnn_encoder_forward(vae, "encoder.pc_encoder.", xyz)
    np.testing.assert_allclose(z.numpy(), np.load(os.path.join(G, "encoder_fpc.npz"))["z_pc"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag,kind,steps", [("ddpm10", "ddpm", 10), ("ddim5", "ddim", 5), ("ddpm100", "ddpm", 100)])
def test_oracle_ldm_generation_vs_reference_loop(fpc, tag, kind, steps):
    _, vae, ddm = fpc
    g = np.load(os.path.join(G, f"ldm_fpc_{tag}.npz"))
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S")
    with torch.no_grad():
        tm, lg = M.generate_grasps_ldm(vae, ddm, xyz, 3, torch.from_numpy(g["x_T"]), noise=torch.from_numpy(g["noise"]),
                                       num_inference_steps=steps, kind=kind)
    np.testing.assert_allclose(tm.numpy(), g["tmrp"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(lg.numpy(), g["logit"], rtol=1e-4, atol=2e-5)


def test_oracle_vae_and_pose_postprocessing(fpc):
    _, vae, _ = fpc
    g = np.load(os.path.join(G, "vae_fpc.npz"))
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S")
    with torch.no_grad():
        tm, lg = M.generate_grasps_vae(vae, xyz, 3, torch.from_numpy(g["z_h"]))
    np.testing.assert_allclose(tm.numpy(), g["tmrp"], rtol=1e-4, atol=2e-5)
    metas = dict(grasp_std=torch.tensor([[.05, .05, .05, .5, .5, .5]]), grasp_mean=torch.zeros(1, 6),
                 pc_std=torch.full((2, 3), 0.05), pc_mean=torch.zeros(2, 3))
    out = M.postprocess(torch.from_numpy(g["tmrp"]), torch.from_numpy(g["logit"]), xyz, metas, 2, 3)
    np.testing.assert_allclose(out["grasp_tmrp"].numpy(), g["grasp_tmrp"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(out["grasps"].numpy(), g["H"], rtol=1e-5, atol=1e-6)


def test_product_schedule_tables_equal_oracle_scheduler():
    from graspldm_b200.schedulers import NoiseSchedule
    kw = dict(num_train_timesteps=1000, beta_start=5e-5, beta_end=1e-3, beta_schedule="linear", variance_type="fixed_large")
    for kind, n in (("ddpm", None), ("ddpm", 100), ("ddpm", 30), ("ddim", 10), ("ddim", 50)):
        ns, so = NoiseSchedule(kind, **kw), SchedulerOracle(kind, **kw)
        if n:
            ns.set_timesteps(n)
            so.set_timesteps(n)
        ts, coef = ns.table()
        from oracle.schedulers import timestep_list
        assert ts == timestep_list(1000, n)
        for i, t in enumerate(ts):
            c = so.coefficients(t)
            want = [c["sqrt_beta_prod_t"], c["sqrt_alpha_prod_t"], c["x0_coeff"],
                    c["xt_coeff"] if kind == "ddpm" else c["eps_coeff"], c["sigma"] if kind == "ddpm" else torch.tensor(0.)]
            assert torch.equal(coef[i, :5], torch.stack([w.float() for w in want])), (kind, n, t)


# ------------------------------------------------------------------ operator oracle vs the reference's CUDA kernels
@pytest.fixture(scope="module")
def gold_ops():
    p = os.path.join(G, "ref_ops_gpu.npz")
    if not os.path.exists(p):
        pytest.skip("ref_ops_gpu.npz not generated yet")
    return np.load(p)


@pytest.mark.parametrize("name,coords", _data.op_cases(), ids=[c[0] for c in _data.op_cases()])
def test_ops_oracle_vs_reference_kernels(gold_ops, name, coords):
    c = coords.numpy()
    for m in _data.FPS_M[name]:
        if m * c.shape[0] * c.shape[2] > 3_000_000:      # keep the CPU suite short
            continue
        assert np.array_equal(ops_np.furthest_point_sampling(c, m), gold_ops[f"{name}/fps{m}"]), f"fps {m}"
    centers = gold_ops[f"{name}/centers"]
    for r, u in _data.BQ:
        assert np.array_equal(ops_np.ball_query(centers, c, r, u), gold_ops[f"{name}/bq{r}_{u}"].astype(np.int32))
    nb = ops_np.ball_query(centers, c, 0.4, 8)
    f = _data.features_for(coords, 5, 11).numpy()
    assert np.array_equal(ops_np.grouping_forward(f, nb), gold_ops[f"{name}/group"])
    cf = _data.features_for(torch.from_numpy(centers), 4, 12).numpy()
    o, i3, w3 = ops_np.three_nearest_neighbors_interpolate_forward(c, centers, cf)
    assert np.array_equal(i3, gold_ops[f"{name}/nn_idx"].astype(np.int32))
    np.testing.assert_allclose(w3, gold_ops[f"{name}/nn_w"], rtol=2e-6, atol=0)
    np.testing.assert_allclose(o, gold_ops[f"{name}/nn_out"], rtol=1e-5, atol=1e-6)
    for r, ch in ((24, 3), (12, 6)):
        vc, nc = _data.vox_coords(coords, r)
        feats = (coords if ch == 3 else _data.features_for(coords, ch, 13)).numpy()
        g, ind, cnt = ops_np.avg_voxelize_forward(feats, vc.numpy(), r)
        assert np.array_equal(ind, gold_ops[f"{name}/vox{r}_ind"].astype(np.int32))
        nz = np.nonzero(cnt)
        assert np.array_equal(np.stack(nz, 0), gold_ops[f"{name}/vox{r}_cnt_nz"].astype(np.int64))
        assert np.array_equal(cnt[nz], gold_ops[f"{name}/vox{r}_cnt_v"].astype(np.int32))
        np.testing.assert_allclose(g.astype(np.float64).sum((0, 2)), gold_ops[f"{name}/vox{r}_sum"], rtol=1e-4, atol=1e-4)
        dv, di, dw = ops_np.trilinear_devoxelize_forward(r, True, nc.numpy(), g)
        np.testing.assert_allclose(dv, gold_ops[f"{name}/devox{r}"], rtol=1e-4, atol=1e-5)   # reference grid came from float atomics
        if name == "gauss100":
            assert np.array_equal(di, gold_ops[f"{name}/devox{r}_inds"])
            assert np.array_equal(dw, gold_ops[f"{name}/devox{r}_wgts"])


def test_oracle_normalize_input_and_per_object_statistics():
    """normalize_input / unnormalize_grasps / unnormalize_pc of the reference (inference_base.py:61-84, 182-212 for one
    cloud; tools/inference.py:570-591 for a batch) on raw, un-centred clouds."""
    g = np.load(os.path.join(G, "normalize_input.npz"))
    shift, gshift = torch.from_numpy(g["pc_shift"]), torch.from_numpy(g["grasp_shift"])
    scale = torch.ones(3) * float(g["translation_scale"])
    gscale = torch.cat((scale, torch.ones(3) * float(g["rotation_scale"])))
    raw = torch.from_numpy(g["raw"])
    for tag, pc, tm in (("single", raw[1], torch.from_numpy(g["tmrp"][1])), ("batch", raw, torch.from_numpy(g["tmrp"]))):
        before = pc.clone()
        pcn, metas = M.normalize_input(pc, shift, scale, gshift, gscale)
        assert torch.equal(pc, before)
        np.testing.assert_array_equal(pcn.numpy(), g[f"{tag}_pc"])
        for k in ("pc_mean", "pc_std", "grasp_mean", "grasp_std"):
            assert metas[k].shape == g[f"{tag}_{k}"].shape, k
            np.testing.assert_array_equal(metas[k].numpy(), g[f"{tag}_{k}"])
        if tag == "batch":
            out = M.postprocess(tm, torch.zeros(12, 1), pcn, metas, 3, 4)
            np.testing.assert_array_equal(out["grasp_tmrp"].numpy(), g["batch_grasp_tmrp"])
            np.testing.assert_allclose(out["grasps"].numpy(), g["batch_H"], rtol=1e-6, atol=1e-7)
            np.testing.assert_array_equal(out["pc"].numpy(), g["batch_pc_unnorm"])
            np.testing.assert_allclose(out["pc"].numpy(), g["raw"], rtol=0, atol=5e-7)     # round trip to the raw cloud



def test_oracle_elucidated_samplers(fpc):
    """Oracle restatement of the elucidated sampler (preconditioning, stochastic Heun, DPM-Solver++ 2M) against the
    outputs of the reference's own ElucidatedDiffusion around the reference denoiser (tests/golden/make_golden.py)."""
    _, _, ddm = fpc
    g = np.load(os.path.join(G, "edm_fpc.npz"))
    t = lambda k: torch.from_numpy(g[k])
    P = "diffusion_model.model."
    with torch.no_grad():
        for k, sg in enumerate((80.0, 2.5, 0.05)):
            got = M.edm_denoise(ddm, t("denoise_x") * sg, sg, t("z_cond"), p=P)
            np.testing.assert_allclose(got.numpy(), g[f"denoise_{k}"], rtol=1e-4, atol=2e-5)
        x = M.edm_sample_heun(ddm, t("z_cond"), t("heun_x_init"), t("heun_noise"), int(g["heun_steps"]), p=P)
        np.testing.assert_allclose(x.numpy(), g["heun_x"], rtol=1e-3, atol=1e-4)
        x = M.edm_sample_dpmpp(ddm, t("z_cond"), t("dpmpp_x_init"), int(g["dpmpp_steps"]), p=P)
        np.testing.assert_allclose(x.numpy(), g["dpmpp_x"], rtol=1e-3, atol=1e-4)


def test_oracle_ppc_ldm_generation_vs_reference_loop():
    """Second model family (partial point clouds: latent 16, conditioning width 256), 10 DDPM steps end to end."""
    m = _models.build("ppc")
    vae, ddm = _models.split_state_dicts(m)
    g = np.load(os.path.join(G, "ldm_ppc_ddpm10.npz"))
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S")
    with torch.no_grad():
        tm, lg = M.generate_grasps_ldm(vae, ddm, xyz, 3, torch.from_numpy(g["x_T"]), noise=torch.from_numpy(g["noise"]),
                                       num_inference_steps=10, kind="ddpm")
    np.testing.assert_allclose(tm.numpy(), g["tmrp"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(lg.numpy(), g["logit"], rtol=1e-4, atol=2e-5)


# --------------------------------------------------------------------------------------------------------------------
# checkpoint-like state: BatchNorm running statistics / every norm's affine parameters away from their init values
# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["fpc", "ppc"])
def test_trained_like_state_matches_the_reference_tree(name):
    """tests/_models.py::trained_like_ gives the product's module tree the same values it gave the reference's tree when
    the *_trained fixtures were generated (hashes of every tensor it changed)."""
    man = json.load(open(os.path.join(G, "state_dict_manifest.json")))[name + "_trained"]
    sd = _models.build_trained_like(name).state_dict()
    kinds = {k.rsplit(".", 1)[-1] for k in man}
    assert {"running_mean", "running_var", "weight", "bias", "g"} <= kinds
    assert any(".norm.weight" in k and "diffusion_model" in k for k in man) and any("voxel_layers" in k for k in man)
    base = _models.build(name).state_dict()
    for k, v in sd.items():
        a = v.detach().contiguous().numpy()
        if k in man:
            assert hashlib.sha256(a.tobytes()).hexdigest()[:16] == man[k]["sha"], k
            assert not np.array_equal(a, base[k].numpy()), k
        else:
            assert np.array_equal(a, base[k].numpy()), k


@pytest.mark.parametrize("name", ["fpc", "ppc"])
def test_oracle_vs_reference_on_trained_like_state(name):
    vae, ddm = _models.split_state_dicts(_models.build_trained_like(name))
    g0 = np.load(os.path.join(G, f"dense_{name}.npz"))
    g = np.load(os.path.join(G, f"dense_{name}_trained.npz"))
    t = lambda k: torch.from_numpy(g0[k])
    with torch.no_grad():
        eps = M.denoiser_forward(ddm, "diffusion_model.model.", t("x"), t("t"), t("z_cond"))
        tm, lg = M.decoder_forward(vae, "decoder.", t("z_h"), t("z_cond"))
    assert np.abs(g["eps"] - g0["eps"]).max() > 1e-2          # the fixture does depend on the norm parameters
    np.testing.assert_allclose(eps.numpy(), g["eps"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(tm.numpy(), g["tmrp"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(lg.numpy(), g["logit"], rtol=1e-5, atol=2e-6)
    xyz = torch.cat([_data.synthetic_clouds(2, seed=1234, dist="S"), _data.synthetic_clouds(1, seed=99, dist="G")])
    with torch.no_grad():
        z = M.pvcnn_encoder_forward(vae, "encoder.pc_encoder.", xyz)
    want = np.load(os.path.join(G, f"encoder_{name}_trained.npz"))["z_pc"]
    assert np.abs(want - np.load(os.path.join(G, f"encoder_{name}.npz"))["z_pc"]).max() > 1e-2
    np.testing.assert_allclose(z.numpy(), want, rtol=1e-4, atol=2e-5)
    for kind, steps in (("ddpm", 100), ("ddim", 10)):
        f = np.load(os.path.join(G, f"ldm_{name}_trained_{kind}{steps}.npz"))
        with torch.no_grad():
            tm, lg = M.generate_grasps_ldm(vae, ddm, xyz[:2], 3, torch.from_numpy(f["x_T"]), noise=torch.from_numpy(f["noise"]),
                                           num_inference_steps=steps, kind=kind)
        np.testing.assert_allclose(tm.numpy(), f["tmrp"], rtol=1e-4, atol=5e-5)
        np.testing.assert_allclose(lg.numpy(), f["logit"], rtol=1e-4, atol=5e-5)


def test_oracle_class_conditioned_denoiser_vs_reference():
    from graspldm_b200 import configs
    from graspldm_b200.resnets import ClassTimeConditionedResNet1D
    g = np.load(os.path.join(G, "cls_fpc.npz"))
    torch.manual_seed(0)
    den = _models.trained_like_(ClassTimeConditionedResNet1D(**configs.model_config("fpc")["denoiser"]), 4).eval()
    np.testing.assert_array_equal(den.cls_embed[0].weight.detach().numpy(), g["cls_w"])
    np.testing.assert_array_equal(den.cls_embed[0].bias.detach().numpy(), g["cls_b"])
    sd = {"diffusion_model.model." + k: v.detach() for k, v in den.state_dict().items()}
    t = lambda k: torch.from_numpy(g[k])
    with torch.no_grad():
        eps = M.denoiser_forward(sd, "diffusion_model.model.", t("x"), t("t"), t("z_cond"), cls_cond=t("cls"))
        x0, _ = M.ldm_sample(sd, t("z_cond"), t("x_T"), noise=t("noise"), num_inference_steps=10, cls_cond=t("cls"))
    np.testing.assert_allclose(eps.numpy(), g["eps"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(x0.numpy(), g["x0"], rtol=1e-4, atol=2e-5)
