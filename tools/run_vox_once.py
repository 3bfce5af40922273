"""Development helper: avg_voxelize_forward alone (C = 48 / R = 12 and C = 3 / R = 24) at B clouds, for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _data
from graspldm_b200 import _pvcnn_backend as be
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
coords = _data.synthetic_clouds(B, 1024, 1, "S").transpose(1, 2).contiguous().to(dev)
vc24, _ = _data.vox_coords(coords.cpu(), 24)
vc12, _ = _data.vox_coords(coords.cpu(), 12)
vc24, vc12 = vc24.to(dev), vc12.to(dev)
f48 = torch.randn(B, 48, 1024, device=dev)
for case, (f, vc, r) in {"c3_r24": (coords, vc24, 24), "c48_r12": (f48, vc12, 12)}.items():
    for _ in range(3):
        be.avg_voxelize_forward(f, vc, r)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        be.avg_voxelize_forward(f, vc, r)
    ev[1].record(); torch.cuda.synchronize()
    print(case, "ms", ev[0].elapsed_time(ev[1]) / reps)
