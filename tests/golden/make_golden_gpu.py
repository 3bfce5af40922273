"""GPU golden generator: runs the reference's OWN CUDA kernels (oracle/_ref/_pvcnn_backend.so, built
unmodified from /root/reference by oracle/build_ref.py) on the seeded operator cases of tests/_data.py
and writes gpurun_out/ref_ops_gpu.npz.  The file is then committed as tests/golden/ref_ops_gpu.npz and
pins oracle/ops_np.py (tests/test_oracle_golden.py) and our kernels (tests/test_ops_gpu.py).

Usage (GPU box):  python tests/golden/make_golden_gpu.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data  # noqa: E402
from oracle import build_ref  # noqa: E402


def main():
    ref = build_ref.load()
    assert ref is not None, "oracle/_ref/_pvcnn_backend.so missing"
    dev = torch.device("cuda:0")
    out = {}
    for name, coords in _data.op_cases():
        c = coords.to(dev)
        for m in _data.FPS_M[name]:
            out[f"{name}/fps{m}"] = ref.furthest_point_sampling(c, m).cpu().numpy()
        m = min(_data.FPS_M[name][-1], 128)     # keep the fixture small
        idx = ref.furthest_point_sampling(c, m)
        centers = ref.gather_features_forward(c, idx)
        out[f"{name}/centers"] = centers.cpu().numpy()
        for r, u in _data.BQ:
            nb = ref.ball_query(centers, c, r, u)
            out[f"{name}/bq{r}_{u}"] = nb.cpu().numpy().astype(np.int16 if c.shape[2] < 32768 else np.int32)
        nb = ref.ball_query(centers, c, 0.4, 8)
        f = _data.features_for(coords, 5, 11).to(dev)
        out[f"{name}/group"] = ref.grouping_forward(f, nb).cpu().numpy()
        # 3-NN from the sampled centres back to all points
        cf = _data.features_for(centers.cpu(), 4, 12).to(dev)
        o, i3, w3 = ref.three_nearest_neighbors_interpolate_forward(c, centers, cf)
        out[f"{name}/nn_out"], out[f"{name}/nn_idx"], out[f"{name}/nn_w"] = (o.cpu().numpy(), i3.cpu().numpy().astype(np.int16), w3.cpu().numpy())
        for r, ch in ((24, 3), (12, 6)):
            vc, nc = _data.vox_coords(coords, r)
            feats = coords if ch == 3 else _data.features_for(coords, ch, 13)
            g, ind, cnt = ref.avg_voxelize_forward(feats.to(dev).contiguous(), vc.to(dev), r)
            torch.cuda.synchronize()
            out[f"{name}/vox{r}_ind"] = ind.cpu().numpy().astype(np.int16)
            out[f"{name}/vox{r}_cnt_nz"] = np.stack(np.nonzero(cnt.cpu().numpy()), 0).astype(np.int16)
            out[f"{name}/vox{r}_cnt_v"] = cnt.cpu().numpy()[np.nonzero(cnt.cpu().numpy())].astype(np.int16)
            out[f"{name}/vox{r}_sum"] = g.double().sum((0, 2)).cpu().numpy()   # per-channel checksum
            dv, di, dw = ref.trilinear_devoxelize_forward(r, True, nc.to(dev).contiguous(), g)
            torch.cuda.synchronize()
            out[f"{name}/devox{r}"] = dv.cpu().numpy()
            if name == "gauss100":
                out[f"{name}/devox{r}_inds"], out[f"{name}/devox{r}_wgts"] = di.cpu().numpy(), dw.cpu().numpy()
                out[f"{name}/vox{r}_grid_nz"] = g.cpu().numpy()[:, :, np.unique(ind.cpu().numpy())]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_ops_gpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
