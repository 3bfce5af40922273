"""GPU: PointNetSAModule (FPS + ball query + the fused grouping / shared-MLP / max kernel), PointNetFPModule, the bare
PointNet2SSG and PVCNN2 networks and the PVCNN-based grasp classifier against fixtures of the unmodified reference classes
(tests/golden/pointnet_family.npz; checkpoint-like weights), plus the fused kernel against the live oracle at other sizes."""
import os

import numpy as np
import pytest
import torch

import _models
from oracle import model_torch as M
from test_pointnet_cpu import build_family

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "pointnet_family.npz"))


def test_sa_and_fp_modules_vs_reference(gold, cuda):
    t = lambda k: torch.from_numpy(gold[k]).to(cuda)
    sa, fp = build_family("sa").to(cuda), build_family("fp").to(cuda)
    f, c = sa((t("sa_in"), t("coords")))
    np.testing.assert_array_equal(c.cpu().numpy(), gold["sa_centers"])                      # FPS: bit-exact indices -> exact gather
    print(f"[SA module] max|err| {np.abs(f.cpu().numpy() - gold['sa_features']).max():.2e}")
    np.testing.assert_allclose(f.cpu().numpy(), gold["sa_features"], rtol=1e-4, atol=2e-5)
    o, pc = fp((t("coords"), c, f, t("sa_in")))
    assert pc.data_ptr() == t("coords").data_ptr() or torch.equal(pc, t("coords"))
    np.testing.assert_allclose(o.cpu().numpy(), gold["fp_features"], rtol=1e-4, atol=2e-5)
    # features=None (coordinates only) and a neighbour count that is not a multiple of the 32-neighbour tile
    from graspldm_b200.pvcnn import PointNetSAModule
    torch.manual_seed(11)
    sa2 = _models.trained_like_(PointNetSAModule(num_centers=37, radius=0.35, num_neighbors=45, in_channels=0,
                                                 out_channels=(20, 300)), 3).eval()
    sd = {k: v.detach().clone() for k, v in sa2.state_dict().items()}
    coords = t("coords")[:, :, :700].contiguous()
    f2, c2 = sa2.to(cuda)((None, coords))
    with torch.no_grad():
        wf, wc = M.sa_module_forward(sd, "", None, coords.cpu(), 37, [0.35], [45])
    np.testing.assert_array_equal(c2.cpu().numpy(), wc.numpy())
    np.testing.assert_allclose(f2.cpu().numpy(), wf.numpy(), rtol=1e-4, atol=2e-5)
    assert f2.shape == (2, 300, 37)


def test_pointnet2ssg_and_pvcnn2_forward_vs_reference(gold, cuda):
    t = lambda k: torch.from_numpy(gold[k]).to(cuda)
    ssg = build_family("ssg").to(cuda)
    out = ssg(t("ssg_in")).cpu().numpy()
    print(f"[PointNet2SSG] max|err| {np.abs(out - gold['ssg_out']).max():.2e} (max|out| {np.abs(gold['ssg_out']).max():.2f})")
    np.testing.assert_allclose(out, gold["ssg_out"], rtol=2e-4, atol=1e-4)
    p2 = build_family("pvcnn2").to(cuda)
    out = p2(t("coords")).cpu().numpy()
    print(f"[PVCNN2] max|err| {np.abs(out - gold['pvcnn2_out']).max():.2e} (max|out| {np.abs(gold['pvcnn2_out']).max():.2f})")
    # voxel ids of points within 1 ulp of a rounding tie may differ from torch's fp32 mean: PVCNN2's
    # normalize=True path computes them with the same torch operations as the reference, so plain fp32 tolerances apply
    np.testing.assert_allclose(out, gold["pvcnn2_out"], rtol=5e-4, atol=2e-4)


def test_grasp_classifier_vs_reference(gold, cuda):
    t = lambda k: torch.from_numpy(gold[k]).to(cuda)
    cls = build_family("classifier").to(cuda)
    pc = t("coords").transpose(1, 2).contiguous()
    preds = cls.classify_grasps(pc, t("cls_grasp_points"))
    print(f"[classifier] preds {preds.cpu().numpy()} vs reference {gold['cls_preds']}")
    assert preds.shape == (2,)
    np.testing.assert_allclose(preds.cpu().numpy(), gold["cls_preds"], rtol=1e-4, atol=1e-5)
    with pytest.raises(NotImplementedError):
        cls(pc, t("cls_grasp_points"))                      # compute_loss=True is training
    with pytest.raises(RuntimeError):
        cls.classify_grasps(pc[:, :100], t("cls_grasp_points"))
