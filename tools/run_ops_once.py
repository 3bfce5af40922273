"""Development helper: every point operator once at B clouds (for ncu captures of the operator kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _data
from graspldm_b200 import _pvcnn_backend as be
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
coords = _data.synthetic_clouds(B, 1024, 1, "S").transpose(1, 2).contiguous().to(dev)
for rep in range(2):
    idx = be.furthest_point_sampling(coords, 256)
    centers = be.gather_features_forward(coords, idx)
    nb = be.ball_query(centers, coords, 0.2, 32)
    feats = torch.randn(B, 32, 1024, device=dev)
    g = be.grouping_forward(feats, nb)
    vc, nc = _data.vox_coords(coords.cpu(), 24)
    vc, nc = vc.to(dev), nc.to(dev).contiguous()
    v1 = be.avg_voxelize_forward(coords, vc, 24)
    grid = torch.randn(B, 48, 24 ** 3, device=dev)
    d = be.trilinear_devoxelize_forward(24, False, nc, grid)
    f48 = torch.randn(B, 48, 1024, device=dev)
    vc12, _ = _data.vox_coords(coords.cpu(), 12)
    v2 = be.avg_voxelize_forward(f48, vc12.to(dev), 12)
    tn = be.three_nearest_neighbors_interpolate_forward(coords, centers, torch.randn(B, 32, 256, device=dev))
torch.cuda.synchronize()
print("ok")
