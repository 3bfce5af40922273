"""CPU checks of the oracle itself (no GPU): a literal thread-by-thread simulation of the reference's
FPS kernel pins the tie-break rule the vectorised oracle uses; brute-force loops pin ball query."""
import numpy as np

import _data
from oracle import ops_np

f32 = np.float32


def _fma(a, b, c):
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def _d(x2, y2, z2, x1, y1, z1):
    dx, dy, dz = f32(x2 - x1), f32(y2 - y1), f32(z2 - z1)
    return _fma(dz, dz, _fma(dx, dx, f32(dy * dy)))


def fps_literal(coords, m, block=512):
    """sampling.cu:86-167 executed literally: per-thread strided scan, shared arrays, pairwise tree."""
    x, y, z = coords
    n = x.shape[0]
    dist = np.full(n, 1e38, f32)
    out = np.zeros(m, np.int32)
    old = 0
    for j in range(1, m):
        dists = np.full(block, -1, f32)
        dists_i = np.zeros(block, np.int64)
        for t in range(block):
            best, besti = f32(-1), 0
            for k in range(t, n, block):
                d2 = min(_d(x[k], y[k], z[k], x[old], y[old], z[old]), dist[k])
                dist[k] = d2
                if d2 > best:
                    best, besti = d2, k
            dists[t], dists_i[t] = best, besti
        u = 0
        while (1 << u) < block:
            for t in range(block >> (u + 1)):
                i1, i2 = (t * 2) << u, (t * 2 + 1) << u
                if dists[i1] < dists[i2]:
                    dists[i1], dists_i[i1] = dists[i2], dists_i[i2]
            u += 1
        old = int(dists_i[0])
        out[j] = old
    return out


def test_fps_oracle_equals_literal_kernel_simulation():
    for coords, m in ((_data.quantised_clouds(1, 700, 3)[0].T.numpy(), 40),
                      (_data.synthetic_clouds(1, 300, 4, "G")[0].T.numpy(), 60),
                      (_data.quantised_clouds(1, 1100, 5, 0.5)[0].T.numpy(), 25)):
        coords = np.ascontiguousarray(coords, dtype=f32)
        assert np.array_equal(ops_np.furthest_point_sampling(coords[None], m)[0], fps_literal(coords, m))


def test_ball_query_oracle_equals_loops():
    pts = _data.synthetic_clouds(1, 200, 1, "G")[0].T.numpy().astype(f32)
    ctr = pts[:, ::7].copy()
    r, u = 0.5, 6
    got = ops_np.ball_query(ctr[None], pts[None], r, u)[0]
    r2 = f32(f32(r) * f32(r))
    for j in range(ctr.shape[1]):
        row, cnt = np.zeros(u, np.int32), 0
        for k in range(pts.shape[1]):
            if cnt >= u:
                break
            d2 = _d(ctr[0, j], ctr[1, j], ctr[2, j], pts[0, k], pts[1, k], pts[2, k])
            if d2 < r2:
                if cnt == 0:
                    row[:] = k
                row[cnt] = k
                cnt += 1
        assert np.array_equal(got[j], row)


def test_voxelize_devoxelize_oracle_properties():
    coords = _data.synthetic_clouds(2, 256, 2, "S").transpose(1, 2).contiguous()
    r = 8
    vc, nc = _data.vox_coords(coords, r)
    feats = np.ones((2, 1, 256), f32)
    grid, ind, cnt = ops_np.avg_voxelize_forward(feats, vc.numpy(), r)
    assert cnt.sum() == 2 * 256 and np.all(ind < r ** 3)
    # mean of ones is one wherever a point fell
    np.testing.assert_allclose(grid[0, 0][cnt[0] > 0], 1.0, rtol=1e-6)
    # trilinear weights sum to one
    _, _, w = ops_np.trilinear_devoxelize_forward(r, True, nc.numpy(), grid)
    np.testing.assert_allclose(w.sum(1), 1.0, rtol=1e-6)
    # devoxelising a constant grid returns the constant
    const = np.full((2, 3, r ** 3), 2.5, f32)
    out, _, _ = ops_np.trilinear_devoxelize_forward(r, False, nc.numpy(), const)
    np.testing.assert_allclose(out, 2.5, rtol=1e-6)


def test_three_nn_oracle_weights():
    pts = _data.synthetic_clouds(1, 64, 3, "G").transpose(1, 2).contiguous().numpy()
    ctr = pts[:, :, :9].copy()
    feats = np.random.default_rng(0).standard_normal((1, 2, 9)).astype(f32)
    out, idx, w = ops_np.three_nearest_neighbors_interpolate_forward(pts, ctr, feats)
    np.testing.assert_allclose(w.sum(1), 1.0, rtol=1e-5)
    # a point that coincides with a centre takes (almost) that centre's feature
    np.testing.assert_allclose(out[0, :, :9], feats[0], rtol=1e-3, atol=1e-3)
    assert np.array_equal(idx[0, 0, :9], np.arange(9))
