"""Drop-in for the reference's pybind11 module `_pvcnn_backend`
(R/grasp_ldm/models/modules/ext/pvcnn/modules/functional/src/bindings.cpp:10-37).

Same function names, argument order, return tuples, output allocation (callee allocates on the
inputs' device) and precondition errors (CUDA device, contiguous, exact dtype -> RuntimeError, as
CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_* do in src/utils.hpp:7-18).  Every call launches on the
CURRENT torch stream (the reference puts voxelize / devoxelize / FPS on the legacy default stream,
vox.cu:114, trilinear_devox.cu:167, sampling.cu:171 - a latent race we do not reproduce).
"""
import torch

from . import _lib


def _chk(x, name, dtype):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if x.dtype != dtype:
        raise RuntimeError(f"{name} must be {'an int' if dtype == torch.int32 else 'a float'} tensor")


def _stream(x):
    return torch.cuda.current_stream(x.device).cuda_stream


def _p(x):
    return x.data_ptr() if x is not None else None


def avg_voxelize_forward(features, coords, resolution):
    """vox.cpp:17-43 -> [out f32[B,C,r^3], ind i32[B,N], cnt i32[B,r^3]]"""
    _chk(features, "features", torch.float32)
    _chk(coords, "coords", torch.int32)
    b, c, n = features.shape
    r = int(resolution)
    with torch.cuda.device(features.device):
        out = torch.empty((b, c, r ** 3), device=features.device, dtype=torch.float32)
        ind = torch.empty((b, n), device=features.device, dtype=torch.int32)
        cnt = torch.empty((b, r ** 3), device=features.device, dtype=torch.int32)
        _lib.call("gldm_avg_voxelize_forward", _p(features), _p(coords), b, c, n, r, _p(out), _p(ind), _p(cnt),
                  _stream(features))
    return [out, ind, cnt]


def avg_voxelize_backward(grad_y, indices, cnt):
    """vox.cpp:54-77"""
    _chk(grad_y, "grad_y", torch.float32)
    _chk(indices, "indices", torch.int32)
    _chk(cnt, "cnt", torch.int32)
    b, c, s = grad_y.shape
    n = indices.shape[1]
    with torch.cuda.device(grad_y.device):
        gx = torch.empty((b, c, n), device=grad_y.device, dtype=torch.float32)
        _lib.call("gldm_avg_voxelize_backward", _p(grad_y), _p(indices), _p(cnt), b, c, n, s, _p(gx),
                  _stream(grad_y))
    return gx


def trilinear_devoxelize_forward(r, is_training, coords, features):
    """trilinear_devox.cpp:18-55 -> [outs f32[B,C,N], inds, wgts] (inds/wgts shape [1] in eval)"""
    _chk(features, "features", torch.float32)
    _chk(coords, "coords", torch.float32)
    b, c = features.shape[:2]
    n = coords.shape[2]
    dev = features.device
    with torch.cuda.device(dev):
        outs = torch.empty((b, c, n), device=dev, dtype=torch.float32)
        if is_training:
            inds = torch.empty((b, 8, n), device=dev, dtype=torch.int32)
            wgts = torch.empty((b, 8, n), device=dev, dtype=torch.float32)
        else:
            inds = torch.zeros((1,), device=dev, dtype=torch.int32)
            wgts = torch.zeros((1,), device=dev, dtype=torch.float32)
        _lib.call("gldm_trilinear_devoxelize_forward", _p(coords), _p(features), b, c, n, int(r),
                  1 if is_training else 0, _p(outs), _p(inds) if is_training else None,
                  _p(wgts) if is_training else None, _stream(features))
    return [outs, inds, wgts]


def trilinear_devoxelize_backward(grad_y, indices, weights, r):
    """trilinear_devox.cpp:67-92"""
    _chk(grad_y, "grad_y", torch.float32)
    _chk(weights, "weights", torch.float32)
    _chk(indices, "indices", torch.int32)
    b, c, n = grad_y.shape
    r3 = int(r) ** 3
    with torch.cuda.device(grad_y.device):
        gx = torch.empty((b, c, r3), device=grad_y.device, dtype=torch.float32)
        _lib.call("gldm_trilinear_devoxelize_backward", _p(grad_y), _p(indices), _p(weights), b, c, n, r3,
                  _p(gx), _stream(grad_y))
    return gx


def furthest_point_sampling(coords, num_samples):
    """sampling.cpp:43-58 -> i32[B,M]"""
    _chk(coords, "coords", torch.float32)
    b, _, n = coords.shape
    m = int(num_samples)
    with torch.cuda.device(coords.device):
        idx = torch.zeros((b, m), device=coords.device, dtype=torch.int32)
        _lib.call("gldm_furthest_point_sampling", _p(coords), b, n, m, _p(idx), _stream(coords))
    return idx


def gather_features_forward(features, indices):
    """sampling.cpp:6-23 -> f32[B,C,M]"""
    _chk(features, "features", torch.float32)
    _chk(indices, "indices", torch.int32)
    b, c, n = features.shape
    m = indices.shape[1]
    with torch.cuda.device(features.device):
        out = torch.empty((b, c, m), device=features.device, dtype=torch.float32)
        _lib.call("gldm_gather_features_forward", _p(features), _p(indices), b, c, n, m, _p(out),
                  _stream(features))
    return out


def gather_features_backward(grad_y, indices, n):
    """sampling.cpp:25-41"""
    _chk(grad_y, "grad_y", torch.float32)
    _chk(indices, "indices", torch.int32)
    b, c, m = grad_y.shape
    with torch.cuda.device(grad_y.device):
        gx = torch.empty((b, c, int(n)), device=grad_y.device, dtype=torch.float32)
        _lib.call("gldm_gather_features_backward", _p(grad_y), _p(indices), b, c, int(n), m, _p(gx),
                  _stream(grad_y))
    return gx


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    """ball_query.cpp:6-30 -> i32[B,M,U]"""
    _chk(centers_coords, "centers_coords", torch.float32)
    _chk(points_coords, "points_coords", torch.float32)
    b, _, m = centers_coords.shape
    n = points_coords.shape[2]
    u = int(num_neighbors)
    with torch.cuda.device(centers_coords.device):
        out = torch.empty((b, m, u), device=centers_coords.device, dtype=torch.int32)
        _lib.call("gldm_ball_query", _p(centers_coords), _p(points_coords), b, n, m, float(radius), u, _p(out),
                  _stream(centers_coords))
    return out


def grouping_forward(features, indices):
    """grouping.cpp:6-24 -> f32[B,C,M,U]"""
    _chk(features, "features", torch.float32)
    _chk(indices, "indices", torch.int32)
    b, c, n = features.shape
    _, m, u = indices.shape
    with torch.cuda.device(features.device):
        out = torch.empty((b, c, m, u), device=features.device, dtype=torch.float32)
        _lib.call("gldm_grouping_forward", _p(features), _p(indices), b, c, n, m, u, _p(out), _stream(features))
    return out


def grouping_backward(grad_y, indices, n):
    """grouping.cpp:26-45"""
    _chk(grad_y, "grad_y", torch.float32)
    _chk(indices, "indices", torch.int32)
    b, c, m, u = grad_y.shape
    with torch.cuda.device(grad_y.device):
        gx = torch.empty((b, c, int(n)), device=grad_y.device, dtype=torch.float32)
        _lib.call("gldm_grouping_backward", _p(grad_y), _p(indices), b, c, int(n), m, u, _p(gx), _stream(grad_y))
    return gx


def three_nearest_neighbors_interpolate_forward(points_coords, centers_coords, centers_features):
    """neighbor_interpolate.cpp:6-40 -> [out f32[B,C,N], idx i32[B,3,N], w f32[B,3,N]]"""
    _chk(points_coords, "points_coords", torch.float32)
    _chk(centers_coords, "centers_coords", torch.float32)
    _chk(centers_features, "centers_features", torch.float32)
    b, c, m = centers_features.shape
    n = points_coords.shape[2]
    dev = points_coords.device
    with torch.cuda.device(dev):
        idx = torch.empty((b, 3, n), device=dev, dtype=torch.int32)
        w = torch.empty((b, 3, n), device=dev, dtype=torch.float32)
        out = torch.empty((b, c, n), device=dev, dtype=torch.float32)
        _lib.call("gldm_three_nn_interpolate_forward", _p(points_coords), _p(centers_coords),
                  _p(centers_features), b, c, m, n, _p(out), _p(idx), _p(w), _stream(points_coords))
    return [out, idx, w]


def three_nearest_neighbors_interpolate_backward(grad_y, indices, weights, m):
    """neighbor_interpolate.cpp:42-64"""
    _chk(grad_y, "grad_y", torch.float32)
    _chk(indices, "indices", torch.int32)
    _chk(weights, "weights", torch.float32)
    b, c, n = grad_y.shape
    with torch.cuda.device(grad_y.device):
        gx = torch.empty((b, c, int(m)), device=grad_y.device, dtype=torch.float32)
        _lib.call("gldm_three_nn_interpolate_backward", _p(grad_y), _p(indices), _p(weights), b, c, n, int(m),
                  _p(gx), _stream(grad_y))
    return gx
