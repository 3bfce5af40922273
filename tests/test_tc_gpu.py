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


@pytest.mark.parametrize("name,kind,steps", [("fpc", "ddpm", 10), ("fpc", "ddim", 5), ("fpc", "ddpm", 100),
                                             ("ppc", "ddpm", 10), ("ppc", "ddim", 5)])
def test_sampler_bf16_tracks_fp32(cuda, name, kind, steps):
    m = _models.build(name, scheduler=kind).to(cuda)
    m.set_inference_timesteps(steps)
    gen = torch.Generator().manual_seed(5)
    n_obj, G_ = 3, 7
    D, Dc = (4, 64) if name == "fpc" else (16, 256)
    z = torch.randn(n_obj, 3, Dc, generator=gen).to(cuda)
    x_T = torch.randn(n_obj * G_, 1, D, generator=gen).to(cuda)
    noise = torch.randn(steps, n_obj * G_, 1, D, generator=gen).to(cuda)
    a, alla = m.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, return_all=True, x_T=x_T, noise=noise,
                                       grasps_per_object=G_)
    b, allb = m.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, return_all=True, x_T=x_T, noise=noise,
                                       grasps_per_object=G_, precision="bf16")
    assert len(allb) == steps + 1 and torch.equal(allb[0], x_T) and torch.equal(allb[-1], b)
    err = (a - b).abs().max().item()
    print(f"[{name} {kind}{steps}] bf16 vs fp32 latents after {steps} steps: max|diff| {err:.3e}")
    # the recurrence damps eps errors (x0 coefficient of the posterior mean is O(1e-2) per step)
    np.testing.assert_allclose(b.cpu().numpy(), a.cpu().numpy(), rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("fold", ["1", "0"])
@pytest.mark.parametrize("name", ["fpc", "ppc"])
def test_encoder_bf16_pointwise_layers(cuda, name, fold, monkeypatch):
    """Conv3d stacks and the point-wise tail (96->768->1536->768->3) on the tensor cores vs the committed reference
    fixture: with conv_downscale and out_layer.0 composed into the epilogue of the 768->1536 GEMM (default) and layer by
    layer (GLDM_FOLD_DOWNSCALE=0)."""
    monkeypatch.setenv("GLDM_FOLD_DOWNSCALE", fold)
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
    np.testing.assert_allclose(z16.cpu().numpy(), want, rtol=1e-2, atol=2e-3)


@pytest.mark.parametrize("fold", ["1", "0"])
def test_encoder_gemm_variants_agree(cuda, fold, monkeypatch):
    """The point-wise GEMM kernels - one 128 x 128 tile per CTA, one 128 x 256 tile per CTA, persistent 128 x 256 tiles
    with double-buffered accumulators (several tiles per CTA at this size) - on the same encoder pass."""
    monkeypatch.setenv("GLDM_FOLD_DOWNSCALE", fold)
    m = _models.build("fpc").to(cuda)
    enc = m.vae_model.encoder.pc_encoder
    enc.precision = "bf16"
    xyz = _data.synthetic_clouds(9, seed=77, dist="S").to(cuda)
    outs = {}
    for name, env in (("tile128", {"GLDM_GEMM_PERSISTENT": "0", "GLDM_GEMM_NT": "1"}),
                      ("tile256", {"GLDM_GEMM_PERSISTENT": "0", "GLDM_GEMM_NT": "2"}),
                      ("persistent", {"GLDM_GEMM_PERSISTENT": "2", "GLDM_GEMM_NT": "2"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        outs[name] = m.vae_model.encode_pc(xyz).cpu().numpy()
    enc.precision = "fp32"
    z32 = m.vae_model.encode_pc(xyz).cpu().numpy()
    # same bf16 operands and K order everywhere; only the order of the fp32 partial sums of the fused projection differs
    np.testing.assert_allclose(outs["tile256"], outs["tile128"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(outs["persistent"], outs["tile128"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(outs["persistent"], z32, rtol=1e-2, atol=2e-3)


def _rot_angle_deg(Ra, Rb):
    """geodesic angle between rotation matrices [..., 3, 3]"""
    R = Ra.transpose(-1, -2) @ Rb
    c = ((R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]) - 1.0) / 2.0
    return torch.rad2deg(torch.acos(c.clamp(-1.0, 1.0)))


def _set_precision(model, prec):
    model.diffusion_model.precision = prec
    model.vae_model.encoder.pc_encoder.precision = prec
    model.vae_model.decoder.precision = prec


@pytest.fixture(params=["row_major", "channel_major"])
def l16_kernel(request, cuda):
    """The L = 16 networks (decoder trunk, ppc denoiser) run on the row-major kernel by default (8 samples per CTA);
    gldm_sampler_tc_set_rows(0) forces the channel-major one (4 samples per CTA).  Both stay covered."""
    from graspldm_b200 import _lib
    _lib.call("gldm_sampler_tc_set_rows", 0 if request.param == "channel_major" else -1)
    yield request.param
    _lib.call("gldm_sampler_tc_set_rows", -1)


@pytest.mark.parametrize("B,gpo", [(1, 1), (4, 1), (37, 1), (40, 20), (300, 20)])
def test_decoder_forward_bf16(fpc, cuda, B, gpo, l16_kernel):
    """Grasp decoder (in_layer -> ResNet1D trunk L = 16 -> heads) on the tensor cores vs the oracle."""
    m, vae, _ = fpc
    gen = torch.Generator().manual_seed(21 + B)
    n_obj = B // gpo
    z_h, zc = torch.randn(B, 4, generator=gen), torch.randn(n_obj, 3, 64, generator=gen)
    with torch.no_grad():
        wt, wl = M.decoder_forward(vae, "decoder.", z_h, zc.repeat_interleave(gpo, 0))
    dec = m.vae_model.decoder
    dec.precision = "bf16"
    try:
        gt, gl = dec(z_h.to(cuda), zc.to(cuda), grasps_per_object=gpo)
    finally:
        dec.precision = "fp32"
    et, el = (gt.cpu() - wt).abs().max().item(), (gl.cpu() - wl).abs().max().item()
    print(f"[B={B}] bf16 tensor-core decoder: max|tmrp err| {et:.3e} (max|tmrp| {wt.abs().max().item():.3f}), max|logit err| {el:.3e}")
    np.testing.assert_allclose(gt.cpu().numpy(), wt.numpy(), rtol=5e-2, atol=5e-2)
    np.testing.assert_allclose(gl.cpu().numpy(), wl.numpy(), rtol=5e-2, atol=5e-2)


def test_ldm_generation_bf16_vs_reference_fixture(cuda):
    """End to end (encoder -> 100 DDPM steps -> decoder -> poses) with every tensor-core kernel switched on, against
    the fixture produced by the unmodified reference modules.  Stated bf16 tolerance of the path:
    latents / tmrp atol 3e-2, translation < 1 mm, rotation < 2 degrees after un-normalisation."""
    from graspldm_b200.inference import InferenceLDM, default_metas
    g = np.load(os.path.join(G, "ldm_fpc_ddpm100.npz"))
    m = _models.build("fpc").to(cuda)
    m.set_inference_timesteps(100)
    _set_precision(m, "bf16")
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S")
    inf = InferenceLDM(m, device=cuda)
    out = inf.generate_grasps(xyz, default_metas(2), num_grasps=3, x_T=torch.from_numpy(g["x_T"]).to(cuda),
                              noise=torch.from_numpy(g["noise"]).to(cuda))
    want = M.postprocess(torch.from_numpy(g["tmrp"]), torch.from_numpy(g["logit"]), xyz, default_metas(2), 2, 3)
    got_H, want_H = out["grasps"].cpu(), want["grasps"]
    dt = (got_H[..., :3, 3] - want_H[..., :3, 3]).norm(dim=-1).max().item()
    da = _rot_angle_deg(got_H[..., :3, :3], want_H[..., :3, :3]).max().item()
    dm = (out["grasp_tmrp"].cpu() - want["grasp_tmrp"]).abs().max().item()
    print(f"[bf16 e2e] max translation error {dt * 1e3:.3f} mm, max rotation error {da:.3f} deg, max|tmrp err| {dm:.3e}")
    assert dt < 1e-3 and da < 2.0
    np.testing.assert_allclose(out["confidence"].cpu().numpy(), want["confidence"].numpy(), atol=2e-2)


def test_full_size_properties_config2_bf16(cuda):
    """BASELINE config 2 on the tensor-core path: determinism, finiteness, proper rotations, object independence."""
    from graspldm_b200.inference import InferenceLDM, default_metas
    m = _models.build("fpc").to(cuda)
    m.set_inference_timesteps(100)
    m.diffusion_model.rng_mode = "fused"
    _set_precision(m, "bf16")
    inf = InferenceLDM(m, device=cuda)
    pcs = _data.synthetic_clouds(64, seed=1234, dist="S")
    x_T = torch.randn(64 * 20, 1, 4, generator=torch.Generator().manual_seed(1)).to(cuda)
    a = inf.generate_grasps(pcs, default_metas(64), num_grasps=20, seed=7, x_T=x_T)
    b = inf.generate_grasps(pcs, default_metas(64), num_grasps=20, seed=7, x_T=x_T)
    assert all(torch.equal(a[k], b[k]) for k in ("grasps", "grasp_tmrp", "confidence"))      # run-to-run deterministic
    assert torch.isfinite(a["grasps"]).all() and a["grasps"].shape == (64, 20, 4, 4)
    R = a["grasps"][..., :3, :3]
    torch.testing.assert_close(R @ R.transpose(-1, -2), torch.eye(3, device=cuda).expand_as(R), rtol=0, atol=1e-5)
    assert (a["confidence"] > 0).all() and (a["confidence"] < 1).all()
    # objects are independent: a slice generated alone (same per-sample noise keys need the same sample indices, so
    # compare the deterministic part, the encoder) and a tile-aligned slice of the sampler
    z_full = m.vae_model.encode_pc(pcs.to(cuda))
    z_part = m.vae_model.encode_pc(pcs[16:24].to(cuda))
    # (GroupNorm statistics are summed per 128-row tile of the whole batch, so the summation order - not the result to
    # fp32 accuracy - depends on where a cloud's rows fall; run to run the encoder is bit-reproducible, see above)
    torch.testing.assert_close(z_full[16:24], z_part, rtol=0, atol=1e-5)
    assert torch.equal(m.vae_model.encode_pc(pcs.to(cuda)), z_full)
    c = inf.generate_grasps(pcs, default_metas(64), num_grasps=20, seed=8, x_T=x_T)
    assert not torch.equal(a["grasp_tmrp"], c["grasp_tmrp"])                                  # the noise seed matters
    # bf16 vs strict-fp32 path on the same inputs and the same in-kernel noise stream
    _set_precision(m, "fp32")
    f = inf.generate_grasps(pcs, default_metas(64), num_grasps=20, seed=7, x_T=x_T)
    dt = (a["grasps"][..., :3, 3] - f["grasps"][..., :3, 3]).norm(dim=-1).max().item()
    da = _rot_angle_deg(a["grasps"][..., :3, :3], f["grasps"][..., :3, :3]).max().item()
    print(f"[config 2, bf16 vs fp32 path] max translation diff {dt * 1e3:.3f} mm, max rotation diff {da:.3f} deg")
    assert dt < 1e-3 and da < 2.0


def test_sampler_two_sets_per_cta_matches_one_set(cuda):
    """32 samples per CTA in two interleaved sets (tensor-core phase of one set under the epilogue of the other) must
    reproduce the single-set kernel bit for bit: the per-sample arithmetic is identical."""
    from graspldm_b200 import _lib
    m = _models.build("fpc").to(cuda)
    m.set_inference_timesteps(10)
    gen = torch.Generator().manual_seed(9)
    n_obj, G_ = 5, 21                    # 105 samples: 4 CTAs of 32 with a ragged tail / 7 CTAs of 16
    z = torch.randn(n_obj, 3, 64, generator=gen).to(cuda)
    x_T = torch.randn(n_obj * G_, 1, 4, generator=gen).to(cuda)
    noise = torch.randn(10, n_obj * G_, 1, 4, generator=gen).to(cuda)
    outs = []
    try:
        _lib.call("gldm_sampler_tc_set_rows", 0)          # the channel-major kernel (default for this model: row-major)
        for sets in (1, 2):
            _lib.call("gldm_sampler_tc_set_sets", sets)
            x0, allx = m.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, return_all=True, x_T=x_T, noise=noise,
                                                grasps_per_object=G_, precision="bf16")
            outs.append((x0.clone(), allx[5].clone()))
    finally:
        _lib.call("gldm_sampler_tc_set_sets", 0)
        _lib.call("gldm_sampler_tc_set_rows", 1)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    try:
        # the row-major kernel (activations as the M operand, 32 samples per CTA) against the channel-major one: same
        # network, different summation orders and bf16 rounding points -> bf16-level agreement; and it is deterministic
        r0, rall = m.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, return_all=True, x_T=x_T, noise=noise,
                                            grasps_per_object=G_, precision="bf16")
        r1, _ = m.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, x_T=x_T, noise=noise, grasps_per_object=G_,
                                         precision="bf16")
    finally:
        _lib.call("gldm_sampler_tc_set_rows", -1)         # back to the automatic choice
    assert torch.equal(r0, r1) and len(rall) == 11
    print(f"row-major vs channel-major sampler after 10 steps: max|diff| {(r0 - outs[0][0]).abs().max().item():.3e}")
    np.testing.assert_allclose(r0.cpu().numpy(), outs[0][0].cpu().numpy(), rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("B,ci,co,r,multi", [(3, 48, 48, 24, "1"), (2, 48, 96, 12, "0"), (5, 96, 96, 12, "0"),
                                             (2, 48, 96, 12, "2"), (5, 96, 96, 12, "2"), (3, 96, 96, 12, "2")])
def test_fused_voxel_branch_ops(cuda, B, ci, co, r, multi, monkeypatch):
    """Channels-last Conv3d (+ GroupNorm statistics) -> GroupNorm + Swish in place (+ SE squeeze) -> SE gate ->
    devoxelize, op by op against torch on the same bf16-rounded operands (pvconv.py:48-67, se.py:10-21).  multi = "2"
    forces the persistent two-tiles-per-weight-stage Conv3d kernel the large batches use (odd and even tile counts)."""
    monkeypatch.setenv("GLDM_CONV3D_MULTI", multi)
    import torch.nn.functional as F
    from graspldm_b200 import _lib
    from graspldm_b200.engine import _aligned_bytes, _stream
    gen = torch.Generator().manual_seed(100 * ci + co + r)
    P, r3, n = (r + 2) ** 3, r ** 3, 500
    x = torch.randn(B, ci, r3, generator=gen).to(cuda)
    w = (torch.randn(co, ci, 3, 3, 3, generator=gen) / (27 * ci) ** 0.5).to(cuda)
    bias, gamma, beta = (torch.randn(co, generator=gen).to(cuda) * s + o for s, o in ((0.3, 0.1), (0.2, 1.0), (0.2, 0.0)))
    st = _stream(cuda)
    img = _aligned_bytes(_lib.lib().gldm_conv3d_tc_weight_bytes(ci), cuda)
    _lib.call("gldm_conv3d_tc_pack_weight", w.contiguous().data_ptr(), co, ci, img.data_ptr(), st)
    cpad_i, cpad_o = -(-ci // 64) * 64, -(-co // 64) * 64
    x_cl = torch.zeros((B * P, cpad_i), device=cuda, dtype=torch.bfloat16)
    _lib.call("gldm_cl_pad", x.data_ptr(), B, ci, r, x_cl.data_ptr(), st)
    xb, wb = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
    want = F.conv3d(xb.view(B, ci, r, r, r), wb, bias, padding=1)                       # fp32 on bf16-rounded operands
    interior = lambda t: t.view(B, r + 2, r + 2, r + 2, -1)[:, 1:-1, 1:-1, 1:-1]        # [B,r,r,r,C] view of a padded grid
    cl = lambda t: t.permute(0, 2, 3, 4, 1)                                             # [B,C,r,r,r] -> [B,r,r,r,C]
    for fp32_out in (0, 1):
        stride = co if fp32_out else cpad_o
        y = torch.zeros((B * P, stride), device=cuda, dtype=torch.float32 if fp32_out else torch.bfloat16)
        stats = torch.full((B, 8, 2), float("nan"), device=cuda, dtype=torch.float64)      # fully overwritten
        ws = torch.empty(_lib.lib().gldm_voxel_ws_bytes(B, max(ci, co), r) // 8 + 1, device=cuda, dtype=torch.float64)
        _lib.call("gldm_conv3d_tc_cl", x_cl.data_ptr(), img.data_ptr(), bias.data_ptr(), B, ci, co, r, y.data_ptr(), fp32_out,
                  stride, stats.data_ptr(), ws.data_ptr(), st)
        got = interior(y)[..., :co].float()
        tol = dict(rtol=2e-3, atol=2e-3) if fp32_out else dict(rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(got, cl(want), **tol)
        full = y.view(B, r + 2, r + 2, r + 2, stride).float()
        assert full[:, 0].abs().max() == 0 and full[:, :, :, -1].abs().max() == 0       # halo rows untouched
        assert stride == co or full[..., co:].abs().max() == 0                          # padding channels written as zero
        g = want.view(B, 8, -1).double()
        torch.testing.assert_close(stats[..., 0], g.sum(-1), rtol=1e-3, atol=0.5)
        torch.testing.assert_close(stats[..., 1], (g * g).sum(-1), rtol=2e-3, atol=0.5)
        # GroupNorm + Swish in place, from the kernel's own statistics, against torch on the kernel's own conv output
        raw = got.permute(0, 4, 1, 2, 3).contiguous()
        se_sum = torch.full((B, co), float("nan"), device=cuda, dtype=torch.float64)
        _lib.call("gldm_gn_swish_cl", y.data_ptr(), fp32_out, stride, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, co,
                  r, 1e-5, se_sum.data_ptr(), ws.data_ptr(), st)
        act = F.silu(F.group_norm(want if fp32_out else raw, 8, gamma, beta, 1e-5))
        got_act = interior(y)[..., :co].float()
        torch.testing.assert_close(got_act, cl(act), **(dict(rtol=5e-3, atol=5e-3) if fp32_out else dict(rtol=2e-2, atol=2e-2)))
        # the SE sums are taken before the bf16 rounding of the stored grid
        torch.testing.assert_close(se_sum.float() / r3, got_act.mean((1, 2, 3)), rtol=1e-4, atol=1e-5 if fp32_out else 5e-4)
        full = y.view(B, r + 2, r + 2, r + 2, stride).float()
        assert full[:, 0].abs().max() == 0 and full[:, :, 0].abs().max() == 0
    # fp32 grid (y, activated) -> SE gate + devoxelize: same operation order as the fp32 kernels => bit-equal
    cr = max(co // 8, 1)
    w1, w2 = torch.randn(cr, co, generator=gen).to(cuda) * 0.2, torch.randn(co, cr, generator=gen).to(cuda) * 0.2
    gate = torch.empty((B, co), device=cuda)
    _lib.call("gldm_se_gate_sum", se_sum.data_ptr(), r3, w1.data_ptr(), w2.data_ptr(), B, co, cr, gate.data_ptr(), st)
    mean = (se_sum / r3).float()
    torch.testing.assert_close(gate, torch.sigmoid(F.silu(mean @ w1.t()) @ w2.t()), rtol=1e-5, atol=1e-6)
    coords = (torch.rand(B, 3, n, generator=gen) * (r - 1)).to(cuda)
    coords[0, :, :4] = torch.tensor([[0.0, r - 1.0, 3.0, 0.5], [0.0, r - 1.0, 2.0, 7.0], [0.0, r - 1.0, 1.0, r - 1.0]], device=cuda)
    point = torch.randn(B, co, n, generator=gen).to(cuda)
    out = torch.empty((B, co, n), device=cuda)
    _lib.call("gldm_devox_cl", coords.data_ptr(), y.data_ptr(), 1, co, gate.data_ptr(), point.data_ptr(), B, co, n, r,
              out.data_ptr(), st)
    grid = interior(y).permute(0, 4, 1, 2, 3).reshape(B, co, r3).contiguous()
    ref = torch.empty((B, co, n), device=cuda)
    _lib.call("gldm_devox_gate_add_f32", coords.data_ptr(), grid.data_ptr(), gate.data_ptr(), point.data_ptr(), B, co, n, r,
              ref.data_ptr(), st)
    assert torch.equal(out, ref)


def test_first_conv3d_channels_last(cuda):
    """3 -> 48 SIMT Conv3d writing the padded channels-last bf16 grid + GroupNorm statistics (pvconv.py:50-52)."""
    import torch.nn.functional as F
    from graspldm_b200 import _lib
    from graspldm_b200.engine import _stream
    gen = torch.Generator().manual_seed(5)
    B, ci, co, r = 3, 3, 48, 24
    P, r3 = (r + 2) ** 3, r ** 3
    x = torch.randn(B, ci, r3, generator=gen).to(cuda)
    w = (torch.randn(co, ci, 3, 3, 3, generator=gen) / 9).to(cuda)
    bias = torch.randn(co, generator=gen).to(cuda) * 0.2
    wp = w.permute(1, 2, 3, 4, 0).reshape(ci, 27, co).contiguous()
    y = torch.zeros((B * P, 64), device=cuda, dtype=torch.bfloat16)
    stats = torch.full((B, 8, 2), float("nan"), device=cuda, dtype=torch.float64)
    ws = torch.empty(_lib.lib().gldm_voxel_ws_bytes(B, co, r) // 8 + 1, device=cuda, dtype=torch.float64)
    _lib.call("gldm_conv3d_k3_f32_cl", x.data_ptr(), wp.data_ptr(), bias.data_ptr(), B, ci, r, y.data_ptr(), 64,
              stats.data_ptr(), ws.data_ptr(), _stream(cuda))
    want = F.conv3d(x.view(B, ci, r, r, r), w, bias, padding=1)
    full = y.view(B, r + 2, r + 2, r + 2, 64).float()
    got = full[:, 1:-1, 1:-1, 1:-1, :co]
    torch.testing.assert_close(got, want.permute(0, 2, 3, 4, 1), rtol=8e-3, atol=8e-3)      # bf16 storage
    assert full[..., co:].abs().max() == 0 and full[:, 0].abs().max() == 0 and full[:, :, -1].abs().max() == 0
    g = want.view(B, 8, -1).double()
    torch.testing.assert_close(stats[..., 0], g.sum(-1), rtol=1e-4, atol=0.05)
    torch.testing.assert_close(stats[..., 1], (g * g).sum(-1), rtol=1e-4, atol=0.05)


def test_l16_row_major_matches_channel_major(cuda):
    """The two tcgen05 kernels of the L = 16 networks compute the same function with different operand roles (different
    summation orders, the attention on mma.sync instead of SIMT): 10 DDPM steps of the ppc sampler and the decoder must
    agree within the bf16 noise floor, ragged sizes included (sample counts that are not multiples of 8 or 4)."""
    from graspldm_b200 import _lib
    m = _models.build("ppc").to(cuda)
    m.set_inference_timesteps(10)
    gen = torch.Generator().manual_seed(77)
    for n_obj, G_ in ((3, 7), (1, 1), (5, 16)):
        n = n_obj * G_
        z = torch.randn(n_obj, 3, 256, generator=gen).to(cuda)
        x_T = torch.randn(n, 1, 16, generator=gen).to(cuda)
        noise = torch.randn(10, n, 1, 16, generator=gen).to(cuda)
        zh = torch.randn(n, 16, generator=gen).to(cuda)
        out = {}
        for kern, flag in (("rows", -1), ("chan", 0)):
            _lib.call("gldm_sampler_tc_set_rows", flag)
            try:
                x, _ = m.diffusion_model.sample(z_cond=z, batch_size=n, x_T=x_T, noise=noise, grasps_per_object=G_, precision="bf16")
                dec = m.vae_model.decoder
                dec.precision = "bf16"
                t, l = dec(zh, z, grasps_per_object=G_)
                dec.precision = "fp32"
            finally:
                _lib.call("gldm_sampler_tc_set_rows", -1)
            out[kern] = (x, t, l)
        for a, b, what in zip(out["rows"], out["chan"], ("latents", "tmrp", "logits")):
            err = (a - b).abs().max().item()
            print(f"[{n_obj}x{G_}] row-major vs channel-major {what}: max|diff| {err:.3e}")
            assert torch.isfinite(a).all()
            assert err < 3e-2, (what, err)


def test_ppc_denoiser_forward_bf16_vs_reference_fixture(cuda, l16_kernel):
    """ppc latent denoiser (16 positions, time conditioned, embedding width 64) on the tensor-core kernels, against the
    fixture of the reference module (resnets.py:558-616)."""
    g = np.load(os.path.join(G, "dense_ppc.npz"))
    m = _models.build("ppc").to(cuda)
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    got = m.diffusion_model.model(t("x"), time=t("t"), z_cond=t("z_cond"), precision="bf16").cpu().numpy()
    print(f"[ppc] bf16 denoiser vs reference fixture: max|err| {np.abs(got - g['eps']).max():.3e}, max|eps| {np.abs(g['eps']).max():.3f}")
    np.testing.assert_allclose(got, g["eps"], rtol=5e-2, atol=5e-2)


@pytest.fixture
def row_major(cuda):
    """Force the row-major sampler kernel (the automatic choice takes it only above one wave of 16-sample CTAs)."""
    from graspldm_b200 import _lib
    _lib.call("gldm_sampler_tc_set_rows", 1)
    yield
    _lib.call("gldm_sampler_tc_set_rows", -1)


def test_row_major_kernel_parity(fpc, cuda, row_major):
    """The row-major tcgen05 sampler kernel (csrc/sampler_rows.cuh) on its own: single evaluation against the oracle,
    full DDPM / DDIM trajectories against the fp32 path, end to end against the reference fixture."""
    from graspldm_b200.inference import InferenceLDM, default_metas
    m, vae, ddm = fpc
    for B in (1, 37, 70):                                    # partially filled CTAs of 32 samples
        gen = torch.Generator().manual_seed(11 + B)
        x, zc = torch.randn(B, 1, 4, generator=gen), torch.randn(B, 3, 64, generator=gen)
        tt = torch.randint(0, 1000, (B,), generator=gen)
        with torch.no_grad():
            want = M.denoiser_forward(ddm, "diffusion_model.model.", x, tt, zc)
        got = m.diffusion_model.model(x.to(cuda), time=tt.to(cuda), z_cond=zc.to(cuda), precision="bf16").cpu()
        print(f"[rows B={B}] denoiser max|err| vs oracle {(got - want).abs().max().item():.3e}")
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=5e-2, atol=5e-2)
    for kind, steps in (("ddpm", 100), ("ddim", 5)):
        mm = _models.build("fpc", scheduler=kind).to(cuda)
        mm.set_inference_timesteps(steps)
        gen = torch.Generator().manual_seed(5)
        n_obj, G_ = 3, 15
        z = torch.randn(n_obj, 3, 64, generator=gen).to(cuda)
        x_T = torch.randn(n_obj * G_, 1, 4, generator=gen).to(cuda)
        noise = torch.randn(steps, n_obj * G_, 1, 4, generator=gen).to(cuda)
        a, _ = mm.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, x_T=x_T, noise=noise, grasps_per_object=G_)
        b, allb = mm.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, return_all=True, x_T=x_T, noise=noise,
                                            grasps_per_object=G_, precision="bf16")
        assert len(allb) == steps + 1 and torch.equal(allb[0], x_T) and torch.equal(allb[-1], b)
        print(f"[rows {kind}{steps}] bf16 vs fp32 latents: max|diff| {(a - b).abs().max().item():.3e}")
        np.testing.assert_allclose(b.cpu().numpy(), a.cpu().numpy(), rtol=3e-2, atol=3e-2)
        # a sample's latent does not depend on its CTA neighbours
        b2, _ = mm.diffusion_model.sample(z_cond=z[:1], batch_size=G_, x_T=x_T[:G_], noise=noise[:, :G_].contiguous(),
                                          grasps_per_object=G_, precision="bf16")
        assert torch.equal(b2, b[:G_])
    g = np.load(os.path.join(G, "ldm_fpc_ddpm100.npz"))
    mm = _models.build("fpc").to(cuda)
    mm.set_inference_timesteps(100)
    _set_precision(mm, "bf16")
    xyz = _data.synthetic_clouds(2, seed=1234, dist="S")
    out = InferenceLDM(mm, device=cuda).generate_grasps(xyz, default_metas(2), num_grasps=3,
                                                        x_T=torch.from_numpy(g["x_T"]).to(cuda),
                                                        noise=torch.from_numpy(g["noise"]).to(cuda))
    want = M.postprocess(torch.from_numpy(g["tmrp"]), torch.from_numpy(g["logit"]), xyz, default_metas(2), 2, 3)
    dt = (out["grasps"].cpu()[..., :3, 3] - want["grasps"][..., :3, 3]).norm(dim=-1).max().item()
    da = _rot_angle_deg(out["grasps"].cpu()[..., :3, :3], want["grasps"][..., :3, :3]).max().item()
    print(f"[rows e2e] max translation error {dt * 1e3:.3f} mm, max rotation error {da:.3f} deg")
    assert dt < 1e-3 and da < 2.0


@pytest.mark.parametrize("persistent", ["0", "2"])
def test_first_conv3d_tensor_core(cuda, persistent, monkeypatch):
    """3 -> 48 Conv3d on the tensor cores (K = 16 rows, SWIZZLE_32B) against torch on the same bf16-rounded operands, incl.
    the padded channels-last output and the GroupNorm statistics: the one-tile kernel (one A tile per filter column) and
    the persistent weight-stationary one the large batches use (one A box per dx plane, several tiles per CTA)."""
    monkeypatch.setenv("GLDM_CONV3D_PERSISTENT16", persistent)
    import torch.nn.functional as F
    from graspldm_b200 import _lib
    from graspldm_b200.engine import _aligned_bytes, _stream
    gen = torch.Generator().manual_seed(6)
    B, ci, co, r = 3, 3, 48, 24
    P, r3 = (r + 2) ** 3, r ** 3
    x = torch.randn(B, ci, r3, generator=gen).to(cuda)
    w = (torch.randn(co, ci, 3, 3, 3, generator=gen) / 9).to(cuda)
    bias = torch.randn(co, generator=gen).to(cuda) * 0.2
    st = _stream(cuda)
    img = _aligned_bytes(_lib.lib().gldm_conv3d_tc16_weight_bytes(), cuda)
    _lib.call("gldm_conv3d_tc16_pack_weight", w.contiguous().data_ptr(), co, ci, img.data_ptr(), st)
    scratch = torch.empty((B * P, 16), device=cuda, dtype=torch.bfloat16)
    y = torch.zeros((B * P, 64), device=cuda, dtype=torch.bfloat16)
    stats = torch.full((B, 8, 2), float("nan"), device=cuda, dtype=torch.float64)
    ws = torch.empty(_lib.lib().gldm_voxel_ws_bytes(B, co, r) // 8 + 1, device=cuda, dtype=torch.float64)
    _lib.call("gldm_conv3d_tc16_cl", x.data_ptr(), img.data_ptr(), bias.data_ptr(), B, ci, co, r, scratch.data_ptr(),
              y.data_ptr(), 64, stats.data_ptr(), ws.data_ptr(), st)
    want = F.conv3d(x.to(torch.bfloat16).float().view(B, ci, r, r, r), w.to(torch.bfloat16).float(), bias, padding=1)
    full = y.view(B, r + 2, r + 2, r + 2, 64).float()
    got = full[:, 1:-1, 1:-1, 1:-1, :co]
    torch.testing.assert_close(got, want.permute(0, 2, 3, 4, 1), rtol=8e-3, atol=8e-3)      # bf16 storage
    assert full[..., co:].abs().max() == 0 and full[:, 0].abs().max() == 0 and full[:, :, -1].abs().max() == 0
    g = want.view(B, 8, -1).double()
    torch.testing.assert_close(stats[..., 0], g.sum(-1), rtol=1e-4, atol=0.05)
    torch.testing.assert_close(stats[..., 1], (g * g).sum(-1), rtol=1e-4, atol=0.05)
