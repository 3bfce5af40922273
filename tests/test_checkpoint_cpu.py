"""Checkpoint ingestion: a Lightning checkpoint of the reference keeps the model under
`ema_model.online_model.` (EMA copy) and `model.` (raw) prefixes (R/tools/inference.py:514-566,
R/grasp_ldm/utils/torch_utils.py:4-37); `inference.load_checkpoint` strips the prefix and loads with strict=True."""
import os

import torch

import _models
from graspldm_b200.inference import fix_state_dict_prefix, load_checkpoint


def _fake_lightning_ckpt(tmp_path, model, ema_scale):
    sd = {}
    for k, v in model.state_dict().items():
        sd["model." + k] = v.clone()
        sd["ema_model.online_model." + k] = v.clone() * ema_scale if v.is_floating_point() else v.clone()
        sd["ema_model.ema_model." + k] = torch.zeros_like(v)            # other prefixes must be ignored
    sd["ema_model.initted"] = torch.tensor(True)
    sd["ema_model.step"] = torch.tensor(123)
    path = os.path.join(tmp_path, "last.ckpt")
    torch.save({"state_dict": sd, "epoch": 7, "optimizer_states": [{}]}, path)
    return path


def test_fix_state_dict_prefix_semantics():
    sd = {"a.b.w": 1, "a.b.c.w": 2, "a.bx.w": 3, "z": 4}
    assert fix_state_dict_prefix(sd, "a.b") == {"w": 1, "c.w": 2}
    assert fix_state_dict_prefix(sd, "a.b", ignore_all_others=False) == {"w": 1, "c.w": 2, "a.bx.w": 3, "z": 4}


def test_load_checkpoint_ema_and_raw(tmp_path):
    src = _models.build("fpc", seed=3)
    path = _fake_lightning_ckpt(str(tmp_path), src, ema_scale=0.5)
    for use_ema, scale in ((True, 0.5), (False, 1.0)):
        dst = _models.build("fpc", seed=11)
        load_checkpoint(dst, path, use_ema_model=use_ema)
        a, b = src.state_dict(), dst.state_dict()
        assert a.keys() == b.keys()
        for k in a:
            want = a[k] * scale if a[k].is_floating_point() else a[k]
            assert torch.equal(b[k], want), k


def test_load_checkpoint_is_strict(tmp_path):
    src = _models.build("fpc", seed=3)
    sd = {"ema_model.online_model." + k: v for k, v in src.state_dict().items()}
    sd.pop(next(iter(sd)))
    path = os.path.join(str(tmp_path), "broken.ckpt")
    torch.save({"state_dict": sd}, path)
    try:
        load_checkpoint(_models.build("fpc", seed=1), path)
    except RuntimeError as e:
        assert "Missing key" in str(e)
    else:
        raise AssertionError("a checkpoint with a missing tensor must not load")


def test_composed_encoder_projection_equals_the_two_layers():
    """conv_downscale (Conv1d 1536 -> 768) followed by out_layer.0 (Conv1d 768 -> 3), pc_encoders.py:60-75: the tensor-core
    path applies them as ONE [3, 1536] projection (engine.compose_affine); same result as the two layers in fp64."""
    import torch
    from graspldm_b200.engine import compose_affine
    g = torch.Generator().manual_seed(5)
    wd, bd = torch.randn(768, 1536, generator=g) * 0.03, torch.randn(768, generator=g) * 0.1
    wo, bo = torch.randn(3, 768, generator=g) * 0.05, torch.randn(3, generator=g) * 0.1
    x = torch.randn(1536, 257, generator=g).double()
    want = wo.double() @ (wd.double() @ x + bd.double()[:, None]) + bo.double()[:, None]
    w, b = compose_affine(wo, bo, wd, bd)
    assert w.shape == (3, 1536) and b.shape == (3,) and w.dtype == torch.float32
    got = w.double() @ x + b.double()[:, None]
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
