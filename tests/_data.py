"""Seeded synthetic inputs shared by the tests, bench.py and the golden generators."""
import numpy as np
import torch


def synthetic_clouds(B, N=1024, seed=1234, dist="S"):
    """(S) points on a sphere surface, per-axis scale U(0.5,1.5), centred; (G) randn*0.7 stress case."""
    g = torch.Generator().manual_seed(seed)
    if dist == "S":
        p = torch.randn(B, N, 3, generator=g)
        p = p / p.norm(dim=-1, keepdim=True)
        p = p * (0.5 + torch.rand(B, 1, 3, generator=g))
        return p - p.mean(1, keepdim=True)
    return torch.randn(B, N, 3, generator=g) * 0.7


def quantised_clouds(B, N, seed, step=0.25):
    """Coordinates snapped to a coarse lattice: many exactly equal distances -> exercises tie-breaks."""
    p = synthetic_clouds(B, N, seed, "G")
    return torch.round(p / step) * step


def op_cases():
    """(name, coords f32[B,3,N]) used for the operator parity tests and the GPU golden file."""
    cases = []
    for name, t in [
        ("sphere1024", synthetic_clouds(3, 1024, 1234, "S")),
        ("gauss1024", synthetic_clouds(2, 1024, 99, "G")),
        ("gauss1000", synthetic_clouds(2, 1000, 5, "G")),
        ("gauss100", synthetic_clouds(2, 100, 6, "G")),
        ("gauss2500", synthetic_clouds(1, 2500, 7, "G")),
        ("ties777", quantised_clouds(2, 777, 8)),
        ("ties1024", quantised_clouds(2, 1024, 9, 0.5)),
    ]:
        cases.append((name, t.transpose(1, 2).contiguous()))
    return cases


FPS_M = {"sphere1024": [1, 16, 256, 1024], "gauss1024": [512], "gauss1000": [128], "gauss100": [100, 130],
         "gauss2500": [300], "ties777": [200], "ties1024": [256]}
BQ = [(0.1, 32), (0.2, 32), (0.4, 64), (0.8, 16)]


def features_for(coords, c, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(coords.shape[0], c, coords.shape[2], generator=g)


def vox_coords(coords, r):
    """Voxelization.forward (normalize=False) with torch ops, as the reference computes it."""
    nc = coords - coords.mean(2, keepdim=True)
    nc = (nc + 1) / 2.0
    nc = torch.clamp(nc * r, 0, r - 1)
    return torch.round(nc).to(torch.int32), nc
