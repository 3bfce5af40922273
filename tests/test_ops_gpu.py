"""GPU parity of the operator FFI (boundary B): our sm_100a kernels, called through the C ABI via
graspldm_b200._pvcnn_backend, against (1) the numpy oracle, (2) the reference's own kernels compiled
unmodified (oracle/_ref, when present) and (3) the committed golden file produced by those kernels.
Index outputs must be bit-exact; float outputs carry the tolerance written next to each check."""
import os

import numpy as np
import pytest
import torch

import _data
from oracle import ops_np

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_ops_gpu.npz")


@pytest.fixture(scope="module")
def be(cuda):
    from graspldm_b200 import _pvcnn_backend
    return _pvcnn_backend


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD) if os.path.exists(GOLD) else None


CASES = _data.op_cases()


@pytest.mark.parametrize("name,coords", CASES, ids=[c[0] for c in CASES])
def test_fps_bit_exact(be, cuda, ref_backend, gold, name, coords):
    c = coords.to(cuda)
    for m in _data.FPS_M[name]:
        got = be.furthest_point_sampling(c, m).cpu().numpy()
        want = ops_np.furthest_point_sampling(coords.numpy(), m)
        assert np.array_equal(got, want), f"{name} m={m}: differs from oracle"
        if ref_backend is not None:
            assert np.array_equal(got, ref_backend.furthest_point_sampling(c, m).cpu().numpy())
        if gold is not None:
            assert np.array_equal(got, gold[f"{name}/fps{m}"])


@pytest.mark.parametrize("name,coords", CASES, ids=[c[0] for c in CASES])
def test_ball_query_grouping_gather_bit_exact(be, cuda, ref_backend, gold, name, coords):
    c = coords.to(cuda)
    m = min(_data.FPS_M[name][-1], 128)
    idx = be.furthest_point_sampling(c, m)
    centers = be.gather_features_forward(c, idx)
    assert np.array_equal(centers.cpu().numpy(), ops_np.gather_features_forward(coords.numpy(), idx.cpu().numpy()))
    if gold is not None:
        assert np.array_equal(centers.cpu().numpy(), gold[f"{name}/centers"])
    for r, u in _data.BQ:
        got = be.ball_query(centers, c, r, u).cpu().numpy()
        want = ops_np.ball_query(centers.cpu().numpy(), coords.numpy(), r, u)
        assert np.array_equal(got, want), f"{name} r={r} u={u}"
        if ref_backend is not None:
            assert np.array_equal(got, ref_backend.ball_query(centers, c, r, u).cpu().numpy())
        if gold is not None:
            assert np.array_equal(got, gold[f"{name}/bq{r}_{u}"].astype(np.int32))
    nb = be.ball_query(centers, c, 0.4, 8)
    f = _data.features_for(coords, 5, 11).to(cuda)
    got = be.grouping_forward(f, nb).cpu().numpy()
    assert np.array_equal(got, ops_np.grouping_forward(f.cpu().numpy(), nb.cpu().numpy()))
    if gold is not None:
        assert np.array_equal(got, gold[f"{name}/group"])


@pytest.mark.parametrize("name,coords", CASES, ids=[c[0] for c in CASES])
def test_three_nn(be, cuda, ref_backend, gold, name, coords):
    c = coords.to(cuda)
    m = min(_data.FPS_M[name][-1], 128)
    centers = be.gather_features_forward(c, be.furthest_point_sampling(c, m))
    cf = _data.features_for(centers.cpu(), 4, 12).to(cuda)
    out, idx, w = be.three_nearest_neighbors_interpolate_forward(c, centers, cf)
    o2, i2, w2 = ops_np.three_nearest_neighbors_interpolate_forward(coords.numpy(), centers.cpu().numpy(), cf.cpu().numpy())
    assert np.array_equal(idx.cpu().numpy(), i2)
    np.testing.assert_allclose(w.cpu().numpy(), w2, rtol=1e-6, atol=0)      # same op order: expect equality
    np.testing.assert_allclose(out.cpu().numpy(), o2, rtol=1e-6, atol=1e-7)
    if ref_backend is not None:
        o3, i3, w3 = ref_backend.three_nearest_neighbors_interpolate_forward(c, centers, cf)
        assert torch.equal(idx, i3)
        np.testing.assert_allclose(w.cpu().numpy(), w3.cpu().numpy(), rtol=2e-6, atol=0)
        np.testing.assert_allclose(out.cpu().numpy(), o3.cpu().numpy(), rtol=1e-5, atol=1e-6)
    if gold is not None:
        assert np.array_equal(idx.cpu().numpy(), gold[f"{name}/nn_idx"].astype(np.int32))
        np.testing.assert_allclose(out.cpu().numpy(), gold[f"{name}/nn_out"], rtol=1e-5, atol=1e-6)


def test_ball_query_clouds_too_large_for_shared_memory(be, cuda):
    """More than 16384 points per cloud: both ball-query kernels read the cloud through the read-only cache instead of a
    shared-memory copy.  Few centres -> the warp-cooperative kernel; the same eight centres tiled to 131072 -> the
    thread = centre kernel (point count not a multiple of 32: the masked last round)."""
    g = torch.Generator().manual_seed(17)
    n = 16400 + 13
    pts = torch.randn(1, 3, n, generator=g) * 0.7
    ctr = pts[:, :, torch.randint(0, n, (8,), generator=g)].contiguous() + 0.01
    for r, u in ((0.15, 32), (0.4, 5)):
        want = ops_np.ball_query(ctr.numpy(), pts.numpy(), r, u)
        got = be.ball_query(ctr.to(cuda), pts.to(cuda), r, u)
        assert np.array_equal(got.cpu().numpy(), want), (r, u)
        big = be.ball_query(ctr.repeat(1, 1, 16384).to(cuda), pts.to(cuda), r, u)
        assert torch.equal(big, got.repeat(1, 16384, 1)), (r, u)


@pytest.mark.parametrize("m", [1, 2, 3, 5, 131, 4100])
def test_three_nn_centre_counts(be, cuda, ref_backend, m):
    """Centre counts that are not a multiple of four (scalar tail of the packed scan), fewer than three centres (unused slots
    keep index 0 and the clamped 1e10 distance, neighbor_interpolate.cu:38-39,60-62), and more than 4096 centres (the kernel
    without the shared-memory copy); one lattice cloud for exact ties."""
    pts = torch.cat([_data.synthetic_clouds(1, 300, 21, "G"), _data.quantised_clouds(1, 300, 22, 0.5)]).transpose(1, 2).contiguous()
    g = torch.Generator().manual_seed(m)
    ctr = torch.cat([torch.randn(1, 3, m, generator=g) * 0.7, torch.round(torch.randn(1, 3, m, generator=g) * 0.7 / 0.5) * 0.5])
    cf = _data.features_for(ctr, 5, 3)
    out, idx, w = be.three_nearest_neighbors_interpolate_forward(pts.to(cuda), ctr.to(cuda), cf.to(cuda))
    o2, i2, w2 = ops_np.three_nearest_neighbors_interpolate_forward(pts.numpy(), ctr.numpy(), cf.numpy())
    assert np.array_equal(idx.cpu().numpy(), i2)
    np.testing.assert_allclose(w.cpu().numpy(), w2, rtol=1e-6, atol=0)
    np.testing.assert_allclose(out.cpu().numpy(), o2, rtol=1e-6, atol=1e-7)
    if ref_backend is not None:
        o3, i3, w3 = ref_backend.three_nearest_neighbors_interpolate_forward(pts.to(cuda), ctr.to(cuda), cf.to(cuda))
        assert torch.equal(idx, i3)
        np.testing.assert_allclose(w.cpu().numpy(), w3.cpu().numpy(), rtol=2e-6, atol=0)


@pytest.mark.parametrize("name,coords", CASES, ids=[c[0] for c in CASES])
def test_voxelize_devoxelize(be, cuda, ref_backend, gold, name, coords):
    for r, ch in ((24, 3), (12, 6)):
        vc, nc = _data.vox_coords(coords, r)
        feats = (coords if ch == 3 else _data.features_for(coords, ch, 13)).contiguous()
        g, ind, cnt = be.avg_voxelize_forward(feats.to(cuda), vc.to(cuda), r)
        g2, ind2, cnt2 = ops_np.avg_voxelize_forward(feats.numpy(), vc.numpy(), r)
        assert np.array_equal(ind.cpu().numpy(), ind2) and np.array_equal(cnt.cpu().numpy(), cnt2)
        # both sum in ascending point order -> bit-exact against the oracle
        assert np.array_equal(g.cpu().numpy(), g2)
        # run-to-run determinism (the reference's float atomics are not)
        g_again = be.avg_voxelize_forward(feats.to(cuda), vc.to(cuda), r)[0]
        assert torch.equal(g, g_again)
        dv, di, dw = be.trilinear_devoxelize_forward(r, True, nc.to(cuda).contiguous(), g)
        dv2, di2, dw2 = ops_np.trilinear_devoxelize_forward(r, True, nc.numpy(), g2)
        assert np.array_equal(di.cpu().numpy(), di2)
        assert np.array_equal(dw.cpu().numpy(), dw2)
        np.testing.assert_allclose(dv.cpu().numpy(), dv2, rtol=1e-6, atol=1e-7)
        dv_eval, i1, w1 = be.trilinear_devoxelize_forward(r, False, nc.to(cuda).contiguous(), g)
        assert torch.equal(dv_eval, dv) and tuple(i1.shape) == (1,) and tuple(w1.shape) == (1,)
        if ref_backend is not None:
            g3, ind3, cnt3 = ref_backend.avg_voxelize_forward(feats.to(cuda), vc.to(cuda), r)
            assert torch.equal(ind, ind3) and torch.equal(cnt, cnt3)
            # reference sums with float atomics in arbitrary order: tolerance, not equality
            np.testing.assert_allclose(g.cpu().numpy(), g3.cpu().numpy(), rtol=1e-5, atol=1e-6)
            dv3, di3, dw3 = ref_backend.trilinear_devoxelize_forward(r, True, nc.to(cuda).contiguous(), g)
            assert torch.equal(di, di3) and torch.equal(dw, dw3)
            np.testing.assert_allclose(dv.cpu().numpy(), dv3.cpu().numpy(), rtol=1e-6, atol=1e-7)
        if gold is not None:
            assert np.array_equal(ind.cpu().numpy(), gold[f"{name}/vox{r}_ind"].astype(np.int32))
            np.testing.assert_allclose(dv.cpu().numpy(), gold[f"{name}/devox{r}"], rtol=1e-4, atol=1e-5)


def test_fused_voxelize_matches_module_math(cuda):
    """gldm_voxelize_fused == Voxelization.forward + avg_voxelize, except for points whose scaled
    coordinate sits within 2 ulp of a rounding tie (the mean is reduced in a different order)."""
    from graspldm_b200 import functional as F
    for name, coords in CASES[:3]:
        for r in (24, 12):
            vc, nc = _data.vox_coords(coords, r)
            grid, norm, vox = F.voxelize_fused(coords.to(cuda), coords.to(cuda), r, return_vox=True)
            np.testing.assert_allclose(norm.cpu().numpy(), nc.numpy(), rtol=0, atol=r * 2.4e-7)
            frac = nc.numpy() - np.floor(nc.numpy())
            fragile = np.abs(frac - 0.5) < r * 5e-7
            same = vox.cpu().numpy() == vc.numpy()
            assert np.all(same | fragile), f"{name} r={r}: {np.sum(~(same | fragile))} voxel ids differ"
            if same.all():
                g2, _, _ = ops_np.avg_voxelize_forward(coords.numpy(), vc.numpy(), r)
                assert np.array_equal(grid.cpu().numpy().reshape(g2.shape), g2)


def test_preconditions(be, cuda):
    x = torch.zeros(1, 3, 8)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        be.furthest_point_sampling(x, 2)
    xc = x.to(cuda)
    with pytest.raises(RuntimeError, match="contiguous"):
        be.furthest_point_sampling(xc.transpose(1, 2)[:, :3, :3].transpose(1, 2)[:, :, ::2], 2)
    with pytest.raises(RuntimeError, match="int tensor"):
        be.gather_features_forward(xc, torch.zeros(1, 2, device=cuda, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="float tensor"):
        be.ball_query(xc.double(), xc, 0.1, 4)


def test_backward_ops_match_autograd_of_oracle_math(be, cuda):
    """The six backward entry points against torch autograd over the equivalent index/gather math."""
    coords = CASES[3][1].to(cuda)        # gauss100
    B, _, N = coords.shape
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, N, (B, 10, 4), generator=g, dtype=torch.int32).to(cuda)
    gy = torch.randn(B, 5, 10, 4, generator=g).to(cuda)
    gx = be.grouping_backward(gy, idx, N)
    want = torch.zeros(B, 5, N, device=cuda).scatter_add_(2, idx.view(B, 1, -1).expand(B, 5, -1).long(), gy.view(B, 5, -1))
    torch.testing.assert_close(gx, want, rtol=1e-5, atol=1e-5)
    gx = be.gather_features_backward(gy[..., 0].contiguous(), idx[..., 0].contiguous(), N)
    want = torch.zeros(B, 5, N, device=cuda).scatter_add_(2, idx[..., 0].view(B, 1, -1).expand(B, 5, -1).long(), gy[..., 0])
    torch.testing.assert_close(gx, want, rtol=1e-5, atol=1e-5)
    # voxelize backward: grad_x[c,i] = grad_y[c, ind[i]] / cnt[ind[i]]
    r = 6
    vc, nc = _data.vox_coords(coords.cpu(), r)
    f = torch.randn(B, 3, N, generator=g).to(cuda)
    out, ind, cnt = be.avg_voxelize_forward(f, vc.to(cuda), r)
    gy = torch.randn(B, 3, r ** 3, generator=g).to(cuda)
    gx = be.avg_voxelize_backward(gy, ind, cnt)
    want = torch.gather(gy, 2, ind.long().view(B, 1, N).expand(B, 3, N)) / torch.gather(cnt, 1, ind.long()).view(B, 1, N)
    torch.testing.assert_close(gx, want, rtol=1e-6, atol=1e-6)
    # devoxelize backward: scatter of w * g
    dv, di, dw = be.trilinear_devoxelize_forward(r, True, nc.to(cuda).contiguous(), out)
    gy = torch.randn(B, 3, N, generator=g).to(cuda)
    gx = be.trilinear_devoxelize_backward(gy, di, dw, r)
    want = torch.zeros(B, 3, r ** 3, device=cuda)
    for k in range(8):
        want.scatter_add_(2, di[:, k].long().view(B, 1, N).expand(B, 3, N), dw[:, k].view(B, 1, N) * gy)
    torch.testing.assert_close(gx, want, rtol=1e-4, atol=1e-5)
    # 3-NN backward
    centers = coords[:, :, :20].contiguous()
    cf = torch.randn(B, 4, 20, generator=g).to(cuda)
    o, i3, w3 = be.three_nearest_neighbors_interpolate_forward(coords, centers, cf)
    gy = torch.randn(B, 4, N, generator=g).to(cuda)
    gx = be.three_nearest_neighbors_interpolate_backward(gy, i3, w3, 20)
    want = torch.zeros(B, 4, 20, device=cuda)
    for k in range(3):
        want.scatter_add_(2, i3[:, k].long().view(B, 1, N).expand(B, 4, N), w3[:, k].view(B, 1, N) * gy)
    torch.testing.assert_close(gx, want, rtol=1e-4, atol=1e-5)


def test_ball_query_large_batch_kernel(be, cuda, ref_backend):
    """Above 131072 centres per call gldm_ball_query switches from the warp-cooperative kernel to the thread = centre kernel
    (csrc/point_ops.cu): same bit-exact contract - checked against the numpy oracle on distinct clouds (odd point count,
    ties lattice included) tiled up to that size, and against the reference's own kernel when it is available."""
    base = torch.cat([_data.synthetic_clouds(2, 777, 3, "S"), _data.quantised_clouds(2, 777, 8, 0.25)]).transpose(1, 2).contiguous()
    m, reps = 129, 260                                       # 4 * 260 * 129 = 134160 centres
    centers_b = be.gather_features_forward(base.to(cuda), be.furthest_point_sampling(base.to(cuda), m))
    pts = base.repeat(reps, 1, 1).to(cuda)
    ctr = centers_b.repeat(reps, 1, 1).contiguous()
    for r, u in ((0.2, 32), (0.45, 7), (2.5, 64)):
        want = ops_np.ball_query(centers_b.cpu().numpy(), base.numpy(), r, u)
        got = be.ball_query(ctr, pts, r, u)
        assert np.array_equal(got[:4].cpu().numpy(), want) and np.array_equal(got[-4:].cpu().numpy(), want), (r, u)
        assert torch.equal(got, got[:4].repeat(reps, 1, 1))
        if ref_backend is not None:
            assert torch.equal(got, ref_backend.ball_query(ctr, pts, r, u))
