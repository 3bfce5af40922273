"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the keys the
driver reads, both arms describe the workload with the same `config` block, and the FLOP bookkeeping behind
`roofline.achieved` matches the figures of SURVEY.md section 8(d)."""
import argparse
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(config=2, ddim_steps=10, objects=256, steps=20, warmup=5, gpus=1)
    d.update(kw)
    return argparse.Namespace(**d)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same configuration block our own arm prints for this --config
    assert d["config"] == bench.workload(_args(config=1), 1)["config"]
    assert "model" not in d["config"] and d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.parametrize("world", [1, 2, 8])
def test_workloads_follow_baseline_json(world):
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) >= 5
    w2 = bench.workload(_args(config=2), world)
    assert (w2["objects"], w2["grasps"], w2["T"], w2["sched"], w2["scaling"]) == (64 * world, 20, 100, "ddpm", "weak")
    assert bench.units_per_step(w2, w2["objects"]) == 1280 * world
    w3 = bench.workload(_args(config=3, ddim_steps=50), world)
    assert (w3["objects"], w3["grasps"], w3["T"], w3["sched"], w3["scaling"]) == (1024, 100, 50, "ddim", "strong")
    w4 = bench.workload(_args(config=4), world)
    assert w4["mode"] == "encoder" and w4["model"] == "ppc" and bench.units_per_step(w4, 4096) == 4096
    w5 = bench.workload(_args(config=5, objects=300), world)
    assert (w5["objects"], w5["grasps"]) == (300 * world, 256)
    for w in (w2, w3, w4, w5):
        assert set(w["config"]) == {"workload", "config_id", "objects", "grasps_per_object", "denoising_steps", "scheduler",
                                    "parallelism"}


def test_flop_bookkeeping_matches_the_survey():
    """8.115 GFLOP per cloud, 7.589 MFLOP per sample and step, 30.70 MFLOP per grasp; a config-2 batch is
    1,530 GFLOP and an object with 256 grasps at T = 100 is 210.3 GFLOP."""
    w2 = bench.workload(_args(config=2), 1)
    assert sum(bench.flops(w2, 64)) == pytest.approx(1530e9, rel=1e-3)
    w5 = bench.workload(_args(config=5), 1)
    assert sum(bench.flops(w5, 1)) == pytest.approx(210.3e9, rel=1e-3)
    w4 = bench.workload(_args(config=4), 1)
    enc, samp, dec = bench.flops(w4, 4096)
    assert enc == pytest.approx(4096 * 8.115e9) and samp == 0 and dec == 0
    assert bench.cpu_sample_objects(w2) == 16          # the bounded CPU sample of VERDICT r1 item 4: >= 16 objects x 20
