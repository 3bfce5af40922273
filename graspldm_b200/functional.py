"""Operator wrappers with the call signatures of the reference's
R/grasp_ldm/models/modules/ext/pvcnn/modules/functional/{voxelization,devoxelization,sampling,
ball_query,grouping,interpolatation}.py, routed to the sm_100a kernels.

This package implements the generation (inference) path: wrappers do not record autograd graphs;
the backward kernels of the FFI are exposed through `_pvcnn_backend` for parity only.
"""
import torch

from . import _lib, _pvcnn_backend as _backend

__all__ = ["avg_voxelize", "trilinear_devoxelize", "gather", "furthest_point_sample", "ball_query",
           "grouping", "nearest_neighbor_interpolate", "voxelize_fused"]


def avg_voxelize(features, coords, resolution):
    """functional/voxelization.py:11-28 -> f32[B,C,R,R,R]"""
    features = features.contiguous()
    coords = coords.int().contiguous()
    b, c, _ = features.shape
    out, _, _ = _backend.avg_voxelize_forward(features, coords, resolution)
    return out.view(b, c, resolution, resolution, resolution)


def trilinear_devoxelize(features, coords, resolution, is_training=True):
    """functional/devoxelization.py:11-31 -> f32[B,C,N]"""
    B, C = features.shape[:2]
    features = features.contiguous().view(B, C, -1)
    outs, _, _ = _backend.trilinear_devoxelize_forward(resolution, is_training, coords.contiguous(), features)
    return outs


def gather(features, indices):
    """functional/sampling.py:13-27"""
    return _backend.gather_features_forward(features.contiguous(), indices.int().contiguous())


def furthest_point_sample(coords, num_samples):
    """functional/sampling.py:39-50: returns the gathered centre coordinates f32[B,3,M]"""
    coords = coords.contiguous()
    return gather(coords, _backend.furthest_point_sampling(coords, num_samples))


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    """functional/ball_query.py:8-19"""
    return _backend.ball_query(centers_coords.contiguous(), points_coords.contiguous(), radius, num_neighbors)


def grouping(features, indices):
    """functional/grouping.py:10-24"""
    return _backend.grouping_forward(features.contiguous(), indices.contiguous())


def nearest_neighbor_interpolate(points_coords, centers_coords, centers_features):
    """functional/interpolatation.py:10-33"""
    out, _, _ = _backend.three_nearest_neighbors_interpolate_forward(
        points_coords.contiguous(), centers_coords.contiguous(), centers_features.contiguous())
    return out


def voxelize_fused(features, coords, resolution, return_vox=False):
    """Voxelization.forward (R/.../pvcnn/modules/voxelization.py:16-35, normalize=False) and avg_voxelize in
    one launch -> (grid f32[B,C,R,R,R], norm_coords f32[B,3,N][, vox i32[B,3,N]])."""
    _backend._chk(features := features.contiguous(), "features", torch.float32)
    _backend._chk(coords := coords.contiguous(), "coords", torch.float32)
    b, c, n = features.shape
    r = int(resolution)
    dev = features.device
    with torch.cuda.device(dev):
        grid = torch.empty((b, c, r, r, r), device=dev, dtype=torch.float32)
        norm = torch.empty((b, 3, n), device=dev, dtype=torch.float32)
        vox = torch.empty((b, 3, n), device=dev, dtype=torch.int32) if return_vox else None
        _lib.call("gldm_voxelize_fused", features.data_ptr(), coords.data_ptr(), b, c, n, r, grid.data_ptr(),
                  norm.data_ptr(), vox.data_ptr() if return_vox else None,
                  torch.cuda.current_stream(dev).cuda_stream)
    return (grid, norm, vox) if return_vox else (grid, norm)
