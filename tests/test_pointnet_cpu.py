"""CPU: the set-abstraction family (PointNetSAModule / PointNetFPModule / PointNet2SSG / PVCNN2) and the grasp classifier.
tests/golden/pointnet_family.npz holds outputs of the UNMODIFIED reference classes over the CPU operator backend
(tests/golden/make_golden.py::pointnet_golden); here (i) the product's module mirrors must reproduce the reference's seeded,
checkpoint-like weights bit for bit (same modules built in the same order, same state_dict keys) and (ii) the oracle's
restatement of the SA / FP modules must reproduce the fixture."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import _models
from oracle import model_torch as M

G = os.path.join(os.path.dirname(__file__), "golden")


def build_family(which):
    from graspldm_b200.grasp_classifier import PointsBasedGraspClassifier
    from graspldm_b200.pvcnn import PVCNN2, PointNet2SSG, PointNetFPModule, PointNetSAModule
    if which == "sa":
        torch.manual_seed(3)
        return _models.trained_like_(PointNetSAModule(num_centers=128, radius=[0.2, 0.4], num_neighbors=[16, 48], in_channels=5,
                                                      out_channels=[(16, 32), (24, 40)]), 5).eval()
    if which == "fp":
        torch.manual_seed(4)
        return _models.trained_like_(PointNetFPModule(in_channels=72 + 5, out_channels=(32, 16)), 6).eval()
    torch.manual_seed(0)
    if which == "ssg":
        return _models.trained_like_(PointNet2SSG(width_multiplier=0.5), 7).eval()
    if which == "pvcnn2":
        return _models.trained_like_(PVCNN2(extra_feature_channels=0, width_multiplier=0.5, voxel_resolution_multiplier=0.5), 8).eval()
    cls = PointsBasedGraspClassifier(
        num_pc_points=1024 + 64,
        points_backbone_config=dict(type="PVCNN", args=dict(in_channels=3, extra_feature_channels=1, scale_channels=0.25,
                                                            scale_voxel_resolution=0.5, num_blocks=(1, 1, 1, 1))))
    return _models.trained_like_(cls, 9).eval()


@pytest.mark.parametrize("which", ["sa", "fp", "ssg", "pvcnn2", "classifier"])
def test_mirrors_reproduce_the_reference_weights(which):
    man = json.load(open(os.path.join(G, "pointnet_manifest.json")))[which]
    sd = build_family(which).state_dict()
    assert set(sd) == set(man)
    for k, v in sd.items():
        a = v.detach().contiguous().numpy()
        assert list(a.shape) == man[k]["shape"], k
        assert hashlib.sha256(a.tobytes()).hexdigest()[:16] == man[k]["sha"], f"{k}: differs from the reference module tree"


def test_oracle_sa_and_fp_modules_vs_reference():
    g = np.load(os.path.join(G, "pointnet_family.npz"))
    t = lambda k: torch.from_numpy(g[k])
    sa_sd = {k: v.detach() for k, v in build_family("sa").state_dict().items()}
    fp_sd = {k: v.detach() for k, v in build_family("fp").state_dict().items()}
    with torch.no_grad():
        f, c = M.sa_module_forward(sa_sd, "", t("sa_in"), t("coords"), 128, [0.2, 0.4], [16, 48])
        np.testing.assert_array_equal(c.numpy(), g["sa_centers"])
        np.testing.assert_allclose(f.numpy(), g["sa_features"], rtol=1e-5, atol=1e-5)
        o = M.fp_module_forward(fp_sd, "", t("coords"), c, f, t("sa_in"))
        np.testing.assert_allclose(o.numpy(), g["fp_features"], rtol=1e-5, atol=1e-5)
