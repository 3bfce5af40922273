"""Operator micro-benchmark on one GPU: our kernels vs the reference's own kernels (oracle/_ref) on the
same inputs, CUDA-event timed.  Writes gpurun_out/ops_bench.json.  Not the headline bench (bench.py)."""
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data  # noqa: E402
from graspldm_b200 import _pvcnn_backend as ours  # noqa: E402
from oracle import build_ref  # noqa: E402


def timeit(fn, iters=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ref = build_ref.load()
    dev = torch.device("cuda:0")
    res = {}
    # (clocks and caches warm before the first timed size: B = 1 is a handful of microseconds per call)
    warm = _data.synthetic_clouds(64, 1024, 1, "S").transpose(1, 2).contiguous().to(dev)
    for _ in range(20):
        ours.furthest_point_sampling(warm, 256)
    torch.cuda.synchronize()
    for B in (1, 64, 1024):
        coords = _data.synthetic_clouds(B, 1024, 1, "S").transpose(1, 2).contiguous().to(dev)
        row = {}
        for name, be in (("ours", ours), ("ref", ref)):
            if be is None:
                continue
            r = {}
            r["fps_1024to256_ms"] = timeit(lambda: be.furthest_point_sampling(coords, 256))
            r["fps_1024to1024_ms"] = timeit(lambda: be.furthest_point_sampling(coords, 1024), iters=5)
            idx = be.furthest_point_sampling(coords, 256)
            centers = be.gather_features_forward(coords, idx)
            r["ball_query_r0.2_u32_ms"] = timeit(lambda: be.ball_query(centers, coords, 0.2, 32))
            nb = be.ball_query(centers, coords, 0.2, 32)
            cfeat = torch.randn(B, 32, 256, device=dev)
            r["three_nn_interpolate_c32_ms"] = timeit(lambda: be.three_nearest_neighbors_interpolate_forward(coords, centers, cfeat))
            feats = torch.randn(B, 32, 1024, device=dev)
            r["grouping_c32_ms"] = timeit(lambda: be.grouping_forward(feats, nb))
            vc, nc = _data.vox_coords(coords.cpu(), 24)
            vc, nc = vc.to(dev), nc.to(dev).contiguous()
            r["avg_voxelize_c3_r24_ms"] = timeit(lambda: be.avg_voxelize_forward(coords, vc, 24))
            grid = torch.randn(B, 48, 24 ** 3, device=dev)
            r["devoxelize_c48_r24_ms"] = timeit(lambda: be.trilinear_devoxelize_forward(24, False, nc, grid))
            f48 = torch.randn(B, 48, 1024, device=dev)
            vc12, nc12 = _data.vox_coords(coords.cpu(), 12)
            vc12 = vc12.to(dev)
            r["avg_voxelize_c48_r12_ms"] = timeit(lambda: be.avg_voxelize_forward(f48, vc12, 12))
            row[name] = r
        # algorithmic bytes per cloud -> achieved GB/s of our kernels
        N, M, U = 1024, 256, 32
        byt = {"avg_voxelize_c3_r24_ms": 4 * (3 * N + 3 * N + 3 * 24 ** 3), "avg_voxelize_c48_r12_ms": 4 * (48 * N + 3 * N + 48 * 12 ** 3),
               "devoxelize_c48_r24_ms": 4 * (48 * 24 ** 3 + 3 * N + 48 * N), "grouping_c32_ms": 4 * (32 * N + M * U + 32 * M * U),
               "fps_1024to256_ms": 12 * N + 4 * M, "ball_query_r0.2_u32_ms": 12 * (N + M) + 4 * M * U}
        row["ours_gbs"] = {k: round(B * v / (row["ours"][k] * 1e-3) / 1e9, 1) for k, v in byt.items()}
        row["ours_rates"] = {"fps_rounds_per_s": B * 256 / (row["ours"]["fps_1024to256_ms"] * 1e-3),
                             "fps_ns_per_round_per_cloud_stream": row["ours"]["fps_1024to256_ms"] * 1e6 / 256,
                             "ball_query_distance_tests_per_s": B * M * N / (row["ours"]["ball_query_r0.2_u32_ms"] * 1e-3)}
        if "ref" in row:
            row["speedup_vs_reference_kernels"] = {k: round(row["ref"][k] / row["ours"][k], 2) for k in row["ours"]}
        res[f"B{B}"] = row
        print(B, json.dumps(row))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ops_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
