"""generate_grasps of InferenceLDM / InferenceVAE (R/tools/inference.py:593-656, :770-815) with the same
arguments and result dictionary.  Experiment / checkpoint / dataset discovery (tools/inference.py:97-328) is
outside the hot path: here a built model is handed in directly; `load_checkpoint`
covers the Lightning .ckpt -> state_dict prefix stripping of tools/inference.py:520-524."""
import torch
from torch import Tensor

from . import engine


def fix_state_dict_prefix(state_dict, prefix, ignore_all_others=True):
    """R/grasp_ldm/utils/torch_utils.py:4-37: keep keys under `prefix.` and strip it."""
    out = {}
    p = prefix + "."
    for k, v in state_dict.items():
        if k.startswith(p):
            out[k[len(p):]] = v
        elif not ignore_all_others:
            out[k] = v
    return out


def load_checkpoint(model, ckpt_path, use_ema_model=True):
    sd = torch.load(ckpt_path, map_location="cpu")["state_dict"]
    sd = fix_state_dict_prefix(sd, "ema_model.online_model" if use_ema_model else "model")
    model.load_state_dict(sd, strict=True)
    return model


def unnormalize_pc(pc: Tensor, metas: dict) -> Tensor:          # tools/inference.py:31-61 (layout/affine only)
    if pc.ndim == 2:
        return pc * metas["pc_std"].to(pc.device) + metas["pc_mean"].to(pc.device)
    return pc * metas["pc_std"].unsqueeze(-2).to(pc.device) + metas["pc_mean"].unsqueeze(-2).to(pc.device)


class _InferenceBase:
    def __init__(self, model, device="cuda:0"):
        self.device = torch.device(device)
        self.model = model.eval().to(self.device)

    def _finish(self, final_grasps, batch_pcs, metas, num_grasps):
        tmrp, cls_logit = final_grasps
        n_pc = batch_pcs.shape[0]
        gm, gs = metas["grasp_mean"].reshape(-1, 6), metas["grasp_std"].reshape(-1, 6)
        if gm.shape[0] != 1 or gs.shape[0] != 1:
            raise NotImplementedError("per-object grasp statistics: the reference datasets use one shared [1,6] row")
        g_un, H, conf = engine.pose_postprocess(tmrp, cls_logit, gm[0], gs[0])
        return dict(grasps=H.view(n_pc, num_grasps, 4, 4), grasp_tmrp=g_un.view(n_pc, num_grasps, 6),
                    confidence=conf.view(n_pc, num_grasps, 1), qualities=None, pc=unnormalize_pc(batch_pcs, metas))


class InferenceLDM(_InferenceBase):
    """tools/inference.py:402-656 (generation part).  use_fast_sampler / num_inference_steps behave as there:
    the step count only takes effect through `model.set_inference_timesteps` (see SURVEY.md finding 2)."""

    def __init__(self, model, device="cuda:0", num_inference_steps=None, fast_sampler=None):
        super().__init__(model, device)
        self.num_inference_steps = num_inference_steps
        self.fast_sampler = fast_sampler

    def generate_grasps(self, pc, metas, num_grasps=10, cls_cond=None, **kwargs):
        batch_pcs = (pc.unsqueeze(0) if pc.ndim == 2 else pc).to(self.device, non_blocking=True)
        metas = {k: v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v for k, v in metas.items()}
        if self.fast_sampler == "DDIM":
            self.model.set_inference_timesteps(self.num_inference_steps)
        final_grasps, _ = self.model.generate_grasps(xyz=batch_pcs, num_grasps=num_grasps, metas=metas, **kwargs)
        out = self._finish(final_grasps, batch_pcs, metas, num_grasps)
        out["all_steps_grasps"] = []
        return out


class InferenceVAE(_InferenceBase):
    """tools/inference.py:770-815."""

    def generate_grasps(self, pc, metas, num_grasps=10, **kwargs):
        batch_pcs = (pc.unsqueeze(0) if pc.ndim == 2 else pc).to(self.device, non_blocking=True)
        metas = {k: v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v for k, v in metas.items()}
        final_grasps = self.model.generate_grasps(batch_pcs, num_grasps, **kwargs)
        return self._finish(final_grasps, batch_pcs, metas, num_grasps)


def default_metas(n_pc=1):
    """Synthetic normalisation statistics of SURVEY.md section 8d (object scale 0.05 m)."""
    return dict(pc_mean=torch.zeros(n_pc, 3), pc_std=torch.full((n_pc, 3), 0.05), grasp_mean=torch.zeros(1, 6),
                grasp_std=torch.tensor([[.05, .05, .05, .5, .5, .5]]), dataset_normalized=True)
