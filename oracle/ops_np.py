"""CPU oracle for the point-cloud operator extension.  TEST INFRASTRUCTURE ONLY.

A numpy restatement of the eight forward kernels of the reference's
`_pvcnn_backend` extension.  SRC = /root/reference/grasp_ldm/models/modules/ext/
pvcnn/modules/functional/src.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the
product (graspldm_b200/) never does.

Parity status: the reference ships no tests or golden vectors for these ops
(SURVEY.md section 4), so the restatement is pinned against outputs of the
reference's own kernels: tests/golden/ref_ops_gpu.npz is produced on a B200 by
tests/golden/make_golden_gpu.py from oracle/_ref/_pvcnn_backend.so (the
reference sources compiled unmodified, see oracle/build_ref.py), and
tests/test_oracle_golden.py checks every function below against it.

Floating-point contract: nvcc's default -fmad=true contracts `a*a + b*b + c*c`
into FMUL, FFMA, FFMA.  `_fma` emulates one fused step as a float64
product-sum rounded once to float32 (the float32 x float32 product is exact in
float64).
"""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    """float32 fma(a, b, c): exact product in float64, one rounding to float32."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _sqdist(dx, dy, dz):
    """d = fma(dz, dz, fma(dx, dx, dy*dy)).

    This is the order nvcc 12.9 emits for `dx*dx + dy*dy + dz*dz` in all three reference
    kernels (sampling.cu:131-132, ball_query.cu:37, neighbor_interpolate.cu:43), read off the
    sm_100a SASS of oracle/_ref/_pvcnn_backend.so: FMUL t=dy*dy; FFMA t=dx*dx+t; FFMA t=dz*dz+t.
    (SURVEY.md section 8a quotes the other association; the SASS is authoritative.)"""
    return _fma(dz, dz, _fma(dx, dx, (dy * dy).astype(f32)))


# --------------------------------------------------------------------------- voxelization
def avg_voxelize_forward(features, coords, r):
    """SRC/voxelization/vox.cpp:17-43, kernels vox.cu:18-72.

    features f32[B,C,N], coords i32[B,3,N] -> (out f32[B,C,r^3], ind i32[B,N], cnt i32[B,r^3]).
    The reference sums with float atomics (order-free); the oracle sums in point order.
    """
    features = np.ascontiguousarray(features, dtype=f32)
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    B, C, N = features.shape
    r2, r3 = r * r, r * r * r
    ind = (coords[:, 0] * r2 + coords[:, 1] * r + coords[:, 2]).astype(np.int32)
    cnt = np.zeros((B, r3), np.int32)
    out = np.zeros((B, C, r3), f32)
    for b in range(B):
        np.add.at(cnt[b], ind[b], 1)
        div = (1.0 / cnt[b][ind[b]].astype(f32).astype(np.float64)).astype(f32)  # vox.cu:65
        contrib = (features[b] * div[None, :]).astype(f32)
        for c in range(C):
            np.add.at(out[b, c], ind[b], contrib[c])
    return out, ind, cnt


def trilinear_devoxelize_forward(r, is_training, coords, features):
    """SRC/interpolate/trilinear_devox.cpp:18-55, kernel trilinear_devox.cu:21-105.

    coords f32[B,3,N] (already in voxel units, clamped to [0,r-1]), features f32[B,C,r^3]
    -> (outs f32[B,C,N], inds i32[B,8,N], wgts f32[B,8,N]) (inds/wgts are shape-[1] zeros in eval).
    """
    coords = np.ascontiguousarray(coords, dtype=f32)
    features = np.ascontiguousarray(features, dtype=f32)
    B, C, _ = features.shape
    N = coords.shape[2]
    r2 = r * r
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    xl, yl, zl = np.floor(x), np.floor(y), np.floor(z)
    xd1, yd1, zd1 = (x - xl).astype(f32), (y - yl).astype(f32), (z - zl).astype(f32)
    xd0, yd0, zd0 = (f32(1) - xd1).astype(f32), (f32(1) - yd1).astype(f32), (f32(1) - zd1).astype(f32)

    def w(a, b, c):
        return ((a * b).astype(f32) * c).astype(f32)

    wg = [w(xd0, yd0, zd0), w(xd0, yd0, zd1), w(xd0, yd1, zd0), w(xd0, yd1, zd1),
          w(xd1, yd0, zd0), w(xd1, yd0, zd1), w(xd1, yd1, zd0), w(xd1, yd1, zd1)]
    xi, yi, zi = xl.astype(np.int32), yl.astype(np.int32), zl.astype(np.int32)
    xh = np.where(xd1 > 0, r2, 0).astype(np.int32)   # (x_hi & r2), trilinear_devox.cu:64-75
    yh = np.where(yd1 > 0, r, 0).astype(np.int32)
    zh = np.where(zd1 > 0, 1, 0).astype(np.int32)
    i000 = xi * r2 + yi * r + zi
    i001 = i000 + zh
    i010 = i000 + yh
    i011 = i010 + zh
    i100 = i000 + xh
    i101 = i100 + zh
    i110 = i100 + yh
    i111 = i110 + zh
    idx = [i000, i001, i010, i011, i100, i101, i110, i111]
    outs = np.zeros((B, C, N), f32)
    for b in range(B):
        f = features[b]                                   # [C, r3]
        # SASS order of (:98-102): FMUL t=w001*f001; FFMA t=w000*f000+t; then FFMA for 010..111
        acc = (wg[1][b][None, :] * f[:, idx[1][b]]).astype(f32)
        for k in (0, 2, 3, 4, 5, 6, 7):
            acc = _fma(np.broadcast_to(wg[k][b][None, :], acc.shape), f[:, idx[k][b]], acc)
        outs[b] = acc
    if is_training:
        return outs, np.stack(idx, 1).astype(np.int32), np.stack(wg, 1).astype(f32)
    return outs, np.zeros((1,), np.int32), np.zeros((1,), f32)


# --------------------------------------------------------------------------- sampling
def furthest_point_sampling(coords, m, block=512):
    """SRC/sampling/sampling.cpp:43-58, kernel sampling.cu:86-167.

    coords f32[B,3,N] -> i32[B,M].  Start index 0, distances start at 1e38.  Winner of a
    round = max of min-distance; ties -> smallest (k mod 512), then smallest k (per-thread
    strided scan keeps the first strictly greater, the tree keeps the lower slot on ties).
    """
    coords = np.ascontiguousarray(coords, dtype=f32)
    B, _, N = coords.shape
    out = np.zeros((B, m), np.int32)
    if m <= 0:
        return out
    ks = np.arange(N)
    key = (ks % block) * (N + 1) + ks            # tie-break ordering
    for b in range(B):
        x, y, z = coords[b, 0], coords[b, 1], coords[b, 2]
        dist = np.full(N, 1e38, f32)
        old = 0
        for j in range(1, m):
            d = _sqdist(x - x[old], y - y[old], z - z[old])   # dx = x2 - x1 (candidate - last)
            dist = np.minimum(d, dist)
            if N == 0:
                old = 0
            else:
                best = dist.max()
                if not (best > f32(-1)):
                    old = 0
                else:
                    cand = np.nonzero(dist == best)[0]
                    old = int(cand[np.argmin(key[cand])])
            out[b, j] = old
    return out


def gather_features_forward(features, indices):
    """SRC/sampling/sampling.cpp:6-23, kernel sampling.cu:17-39: out[b,c,j] = feat[b,c,idx[b,j]]."""
    features = np.asarray(features, dtype=f32)
    indices = np.asarray(indices, dtype=np.int32)
    return np.take_along_axis(features, indices[:, None, :].astype(np.int64), axis=2).astype(f32)


# --------------------------------------------------------------------------- ball query / grouping
def ball_query(centers, points, radius, u):
    """SRC/ball_query/ball_query.cpp:6-30, kernel ball_query.cu:19-50.

    centers f32[B,3,M], points f32[B,3,N] -> i32[B,M,U]: the first U points (ascending index)
    with d2 < r2 (strict); the first hit pre-fills all U slots; no hit leaves zeros.
    dx = center - point; r2 = float(radius)*float(radius) (ball_query.cpp:24).
    """
    centers = np.ascontiguousarray(centers, dtype=f32)
    points = np.ascontiguousarray(points, dtype=f32)
    B, _, M = centers.shape
    N = points.shape[2]
    r2 = f32(f32(radius) * f32(radius))
    out = np.zeros((B, M, u), np.int32)
    for b in range(B):
        dx = centers[b, 0][:, None] - points[b, 0][None, :]
        dy = centers[b, 1][:, None] - points[b, 1][None, :]
        dz = centers[b, 2][:, None] - points[b, 2][None, :]
        hit = _sqdist(dx.astype(f32), dy.astype(f32), dz.astype(f32)) < r2      # [M,N]
        for j in range(M):
            ks = np.nonzero(hit[j])[0][:u]
            if ks.size:
                out[b, j, :] = ks[0]
                out[b, j, :ks.size] = ks
    return out


def grouping_forward(features, indices):
    """SRC/grouping/grouping.cpp:6-24, kernel grouping.cu:18-36: out[b,c,m,u] = feat[b,c,idx[b,m,u]]."""
    features = np.asarray(features, dtype=f32)
    indices = np.asarray(indices)
    B, C, N = features.shape
    _, M, U = indices.shape
    flat = indices.reshape(B, 1, M * U).astype(np.int64)
    return np.take_along_axis(features, np.broadcast_to(flat, (B, C, M * U)), axis=2).reshape(B, C, M, U)


# --------------------------------------------------------------------------- 3-NN interpolation
def three_nearest_neighbors_interpolate_forward(points, centers, feats):
    """SRC/interpolate/neighbor_interpolate.cpp:6-40, kernels neighbor_interpolate.cu:20-116.

    points f32[B,3,N], centers f32[B,3,M], feats f32[B,C,M] -> (out f32[B,C,N], idx i32[B,3,N], w f32[B,3,N]).
    Distances are float32 (same FMA chain, dx = point - center); the three bests are kept
    as doubles with strict '<' (earlier centre wins ties); clamp to [1e-10, 1e10]; weights
    are products of the other two distances over their pairwise-product sum (:59-72).
    """
    points = np.ascontiguousarray(points, dtype=f32)
    centers = np.ascontiguousarray(centers, dtype=f32)
    feats = np.ascontiguousarray(feats, dtype=f32)
    B, C, M = feats.shape
    N = points.shape[2]
    idx = np.zeros((B, 3, N), np.int32)
    wts = np.zeros((B, 3, N), f32)
    out = np.zeros((B, C, N), f32)
    for b in range(B):
        dx = (points[b, 0][:, None] - centers[b, 0][None, :]).astype(f32)
        dy = (points[b, 1][:, None] - centers[b, 1][None, :]).astype(f32)
        dz = (points[b, 2][:, None] - centers[b, 2][None, :]).astype(f32)
        d = _sqdist(dx, dy, dz)                                   # [N,M]
        # stable sort == sequential strict-'<' insertion order for ties
        order = np.argsort(d, axis=1, kind="stable")
        bi = np.zeros((N, 3), np.int64)
        bd = np.full((N, 3), 1e40, np.float64)
        take = min(3, M)
        bi[:, :take] = order[:, :take]
        bd[:, :take] = np.take_along_axis(d, order[:, :take], axis=1).astype(np.float64)
        bd = np.maximum(np.minimum(np.float64(f32(1e10)), bd), np.float64(f32(1e-10)))
        d0d1 = (bd[:, 0] * bd[:, 1]).astype(f32)
        d0d2 = (bd[:, 0] * bd[:, 2]).astype(f32)
        d1d2 = (bd[:, 1] * bd[:, 2]).astype(f32)
        inv = (f32(1.0) / ((d0d1 + d0d2).astype(f32) + d1d2).astype(f32)).astype(f32)
        w0, w1, w2 = (d1d2 * inv).astype(f32), (d0d2 * inv).astype(f32), (d0d1 * inv).astype(f32)
        idx[b] = bi.T.astype(np.int32)
        wts[b] = np.stack([w0, w1, w2], 0)
        f = feats[b]
        acc = (f[:, bi[:, 1]] * w1[None, :]).astype(f32)           # FMUL(2nd), FFMA(1st), FFMA(3rd) (:111-113)
        acc = _fma(f[:, bi[:, 0]], np.broadcast_to(w0[None, :], acc.shape), acc)
        acc = _fma(f[:, bi[:, 2]], np.broadcast_to(w2[None, :], acc.shape), acc)
        out[b] = acc
    return out, idx, wts
