"""Edge cases of the path on the GPU: empty inputs, the documented maximum sizes, ragged batch sizes (partially
filled sampler CTAs) and independence of a sample's result from what else is in the batch."""
import numpy as np
import pytest
import torch

import _data
import _models
from oracle import ops_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be(cuda):
    from graspldm_b200 import _pvcnn_backend
    return _pvcnn_backend


def test_empty_inputs(be, cuda):
    """b = 0 / m = 0: every entry point returns correctly shaped empty tensors (the reference launches zero-sized
    grids, sampling.cu:160-172, which CUDA rejects; here they are no-ops)."""
    c0 = torch.zeros((0, 3, 64), device=cuda)
    assert be.furthest_point_sampling(c0, 8).shape == (0, 8)
    c1 = torch.randn(2, 3, 64, device=cuda)
    assert be.furthest_point_sampling(c1, 0).shape == (2, 0)
    f0 = torch.zeros((0, 5, 64), device=cuda)
    out, ind, cnt = be.avg_voxelize_forward(f0, torch.zeros((0, 3, 64), device=cuda, dtype=torch.int32), 4)
    assert out.shape == (0, 5, 64) and ind.shape == (0, 64) and cnt.shape == (0, 64)
    assert be.ball_query(torch.zeros((0, 3, 4), device=cuda), c0, 0.2, 8).shape == (0, 4, 8)
    assert be.gather_features_forward(f0, torch.zeros((0, 7), device=cuda, dtype=torch.int32)).shape == (0, 5, 7)
    from graspldm_b200 import engine
    gt, H, conf = engine.pose_postprocess(torch.zeros((0, 6), device=cuda), torch.zeros((0, 1), device=cuda),
                                          torch.zeros(1, 6), torch.ones(1, 6), grasps_per_obj=3)
    assert gt.shape == (0, 6) and H.shape == (0, 4, 4) and conf.shape == (0, 1)
    pcn, pm, gm = engine.normalize_clouds(torch.zeros((0, 16, 3), device=cuda), torch.zeros(3), torch.ones(3), torch.zeros(6))
    assert pcn.shape == (0, 16, 3) and pm.shape == (0, 3) and gm.shape == (0, 6)
    torch.cuda.synchronize()


def test_maximum_sizes(be, cuda):
    """FPS up to 32768 points per cloud (bit-exact against the oracle), voxelize up to 8192 points / resolution 36;
    one past either limit is a RuntimeError, not a wrong answer."""
    g = torch.Generator().manual_seed(3)
    big = torch.randn(1, 3, 32768, generator=g)
    idx = be.furthest_point_sampling(big.to(cuda), 24).cpu().numpy()
    np.testing.assert_array_equal(idx, ops_np.furthest_point_sampling(big.numpy(), 24))
    with pytest.raises(RuntimeError, match="32768"):
        be.furthest_point_sampling(torch.randn(1, 3, 32769, device=cuda), 4)
    n, r, c = 8192, 32, 2
    coords = torch.randint(0, r, (1, 3, n), generator=g, dtype=torch.int32)
    feats = torch.randn(1, c, n, generator=g)
    out, ind, cnt = be.avg_voxelize_forward(feats.to(cuda), coords.to(cuda), r)
    wo, wi, wc = ops_np.avg_voxelize_forward(feats.numpy(), coords.numpy(), r)
    np.testing.assert_array_equal(ind.cpu().numpy(), wi)
    np.testing.assert_array_equal(cnt.cpu().numpy(), wc)
    np.testing.assert_allclose(out.cpu().numpy(), wo, rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError, match="8192"):
        be.avg_voxelize_forward(torch.randn(1, c, n + 1, device=cuda), torch.zeros((1, 3, n + 1), device=cuda, dtype=torch.int32), r)
    with pytest.raises(RuntimeError, match="36"):
        be.avg_voxelize_forward(feats.to(cuda), coords.to(cuda), 37)


@pytest.mark.parametrize("b,c,n,r,mode", [
    (2, 3, 1024, 24, "one_voxel"),      # every point in the same voxel: one bucket of 1024 (the rank-by-counting worst case)
    (3, 5, 1001, 5, "random"),          # n % 4 != 0 (no 16-byte feature staging), r^3 % 4 != 0 (scalar cells), odd channel count
    (100, 48, 64, 12, "random"),        # 16-channel slices (wide features, large batch)
    (2, 48, 4096, 12, "clustered"),     # ppc-sized clouds, features too large to stage, heavy buckets
    (1, 9, 31, 24, "random"),           # fewer points than a warp; cells split over several CTAs (blockIdx.z)
])
def test_voxelize_bucket_paths(be, cuda, b, c, n, r, mode):
    """avg_voxelize_forward (vox.cu:18-72) on the shapes that take the kernel's other branches; ind / cnt and the grid are
    bit-exact against the oracle (both sum in ascending point index)."""
    g = torch.Generator().manual_seed(b * 1000 + n)
    if mode == "one_voxel":
        coords = torch.full((b, 3, n), r // 2, dtype=torch.int32)
    elif mode == "clustered":
        coords = (torch.randn(b, 3, n, generator=g) * 1.2 + r / 2).round().clamp(0, r - 1).to(torch.int32)
    else:
        coords = torch.randint(0, r, (b, 3, n), generator=g, dtype=torch.int32)
    feats = torch.randn(b, c, n, generator=g)
    out, ind, cnt = be.avg_voxelize_forward(feats.to(cuda), coords.to(cuda), r)
    wo, wi, wc = ops_np.avg_voxelize_forward(feats.numpy(), coords.numpy(), r)
    np.testing.assert_array_equal(ind.cpu().numpy(), wi)
    np.testing.assert_array_equal(cnt.cpu().numpy(), wc)
    np.testing.assert_array_equal(out.cpu().numpy(), wo)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ragged_batches_and_batch_independence(cuda, precision):
    """1 sample, 33 samples (a partially filled second / third CTA), uneven grasps per object: a sample's latent does
    not depend on which other samples share its launch (same x_T / noise rows => bit-identical rows)."""
    m = _models.build("fpc").to(cuda)
    m.set_inference_timesteps(10)
    gen = torch.Generator().manual_seed(17)
    n_obj, G_ = 11, 3                                           # 33 samples
    z = torch.randn(n_obj, 3, 64, generator=gen).to(cuda)
    x_T = torch.randn(n_obj * G_, 1, 4, generator=gen).to(cuda)
    noise = torch.randn(10, n_obj * G_, 1, 4, generator=gen).to(cuda)
    kw = dict(precision=precision) if precision != "fp32" else {}
    full, _ = m.diffusion_model.sample(z_cond=z, batch_size=n_obj * G_, x_T=x_T, noise=noise, grasps_per_object=G_, **kw)
    assert full.shape == (33, 1, 4) and torch.isfinite(full).all()
    # first 5 objects alone (15 samples: one partially filled CTA)
    part, _ = m.diffusion_model.sample(z_cond=z[:5], batch_size=15, x_T=x_T[:15], noise=noise[:, :15].contiguous(),
                                       grasps_per_object=G_, **kw)
    assert torch.equal(part, full[:15])
    # the last object alone, one grasp of it (n = 1)
    one, _ = m.diffusion_model.sample(z_cond=z[10:], batch_size=1, x_T=x_T[30:31], noise=noise[:, 30:31].contiguous(),
                                      grasps_per_object=1, **kw)
    assert torch.equal(one, full[30:31])
    # decoder on the ragged batch
    m.vae_model.decoder.precision = precision
    tm, lg = m.vae_model.decoder(full.squeeze(1), z, grasps_per_object=G_)
    tm1, lg1 = m.vae_model.decoder(full.squeeze(1)[30:31], z[10:], grasps_per_object=1)
    assert tm.shape == (33, 6) and lg.shape == (33, 1) and torch.equal(tm1, tm[30:31]) and torch.equal(lg1, lg[30:31])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_encoder_batch_independence_and_single_cloud(cuda, precision):
    m = _models.build("fpc").to(cuda)
    m.vae_model.encoder.pc_encoder.precision = precision
    xyz = _data.synthetic_clouds(5, seed=4, dist="S").to(cuda)
    z5 = m.vae_model.encode_pc(xyz)
    z1 = m.vae_model.encode_pc(xyz[3:4])
    assert z5.shape == (5, 3, 64) and z1.shape == (1, 3, 64)
    # the statistics of the fused voxel branch are summed per cloud in an order that depends on where the cloud's rows
    # fall in the 128-row tiles of the batch, so a cloud encoded alone can differ in the last bits (and, rarely, by one
    # bf16 rounding flip of a voxel activation) from the same cloud inside a batch; run to run everything is bit-equal
    if precision == "fp32":
        assert torch.equal(z1[0], z5[3])
    else:
        torch.testing.assert_close(z1[0], z5[3], rtol=0, atol=1e-5)
        assert torch.equal(m.vae_model.encode_pc(xyz), z5)
