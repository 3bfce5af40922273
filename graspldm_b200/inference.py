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
    def __init__(self, model, device="cuda:0", norm_config=None):
        self.device = torch.device(device)
        self.model = model.eval().to(self.device)
        if norm_config is not None:
            self.set_normalization_params(norm_config)

    def set_normalization_params(self, norm_config):
        """R/grasp_ldm/inference/inference_base.py:103-130; `norm_config` is an object or dict with `pc_shift` [3],
        `grasp_shift` [6], `translation_scale`, `rotation_scale`."""
        get = (lambda k: norm_config[k]) if isinstance(norm_config, dict) else (lambda k: getattr(norm_config, k))
        for k in ("pc_shift", "grasp_shift", "translation_scale", "rotation_scale"):
            try:
                get(k)
            except (KeyError, AttributeError):
                raise AssertionError(f"norm_config should have `{k}`")
        f = dict(dtype=torch.float32, device=self.device)
        self._INPUT_PC_SHIFT = torch.tensor(get("pc_shift"), **f)
        self._INPUT_GRASP_SHIFT = torch.tensor(get("grasp_shift"), **f)
        ones = torch.ones((3,), **f)
        self._INPUT_PC_SCALE = ones * get("translation_scale")
        self._INPUT_GRASP_SCALE = torch.cat((ones * get("translation_scale"), ones * get("rotation_scale")))

    def normalize_input(self, pc):
        """Raw cloud(s) [N,3] / [B,N,3] -> (normalised cloud, metas): centre on the cloud mean, then the dataset shift and
        scale (inference_base.py:182-212; tools/inference.py:570-591 for the batched metas layout).  Two deliberate
        differences from the reference: the caller's tensor is not centred in place, and the stored grasp shift is not
        accumulated across calls (inference_base.py:202-203 adds pc_mean into `_INPUT_GRASP_SHIFT` itself)."""
        assert pc.ndim in (2, 3)
        assert hasattr(self, "_INPUT_PC_SHIFT"), "call set_normalization_params(norm_config) first"
        single = pc.ndim == 2
        src = (pc.unsqueeze(0) if single else pc).to(self.device, non_blocking=True)
        out, pc_mean, grasp_mean = engine.normalize_clouds(src, self._INPUT_PC_SHIFT, self._INPUT_PC_SCALE,
                                                           self._INPUT_GRASP_SHIFT)
        if single:
            out, pc_mean, grasp_mean = out[0], pc_mean[0], grasp_mean[0]
        metas = dict(pc_mean=pc_mean, pc_std=self._INPUT_PC_SCALE if single else self._INPUT_PC_SCALE.unsqueeze(0),
                     grasp_mean=grasp_mean,
                     grasp_std=self._INPUT_GRASP_SCALE if single else self._INPUT_GRASP_SCALE.unsqueeze(0),
                     use_dataset_statistics=False, dataset_normalized=True)
        return out, metas

    def generate_on_pointcloud(self, pc, num_grasps=10, return_intermediate=False, **kwargs):
        """inference_base.py:161-180 (named infer_on_pointcloud in tools/inference.py:658-666)."""
        pc_normalized, metas = self.normalize_input(pc)
        return self.generate_grasps(pc_normalized, metas, num_grasps=num_grasps,
                                    return_intermediate=return_intermediate, **kwargs)

    infer_on_pointcloud = generate_on_pointcloud

    def _finish(self, final_grasps, batch_pcs, metas, num_grasps):
        tmrp, cls_logit = final_grasps
        n_pc = batch_pcs.shape[0]
        gm, gs = metas["grasp_mean"].reshape(-1, 6), metas["grasp_std"].reshape(-1, 6)
        for name, v in (("grasp_mean", gm), ("grasp_std", gs)):
            if v.shape[0] not in (1, n_pc):
                raise ValueError(f"metas['{name}'] has {v.shape[0]} rows for {n_pc} clouds")
        g_un, H, conf = engine.pose_postprocess(tmrp, cls_logit, gm, gs, grasps_per_obj=num_grasps)
        pm, psd = metas["pc_mean"], metas["pc_std"]
        pc_metas = dict(pc_mean=pm.reshape(-1, 3) if pm.ndim == 1 else pm, pc_std=psd.reshape(-1, 3) if psd.ndim == 1 else psd)
        return dict(grasps=H.view(n_pc, num_grasps, 4, 4), grasp_tmrp=g_un.view(n_pc, num_grasps, 6),
                    confidence=conf.view(n_pc, num_grasps, 1), qualities=None, pc=unnormalize_pc(batch_pcs, pc_metas))


class InferenceLDM(_InferenceBase):
    """tools/inference.py:402-656 (generation part).  use_fast_sampler / num_inference_steps behave as there:
    the step count only takes effect through `model.set_inference_timesteps` (see SURVEY.md finding 2)."""

    def __init__(self, model, device="cuda:0", num_inference_steps=None, fast_sampler=None, norm_config=None):
        super().__init__(model, device, norm_config)
        self.num_inference_steps = num_inference_steps
        self.fast_sampler = fast_sampler

    def generate_grasps(self, pc, metas, num_grasps=10, cls_cond=None, return_intermediate=False, **kwargs):
        batch_pcs = (pc.unsqueeze(0) if pc.ndim == 2 else pc).to(self.device, non_blocking=True)
        metas = {k: v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v for k, v in metas.items()}
        if self.fast_sampler == "DDIM":                          # tools/inference.py:605-609
            self.model.set_inference_timesteps(self.num_inference_steps)
        elif self.fast_sampler == "DPMPP":                       # elucidated models: DPM-Solver++ with the configured steps
            kwargs["use_dpmpp"] = True
            kwargs["num_sample_steps"] = self.num_inference_steps
        elif getattr(self.model, "is_elucidated_diffusion", False):
            kwargs.setdefault("use_dpmpp", False)                # the reference's sample() pops the key unconditionally
        final_grasps, step_grasps = self.model.generate_grasps(xyz=batch_pcs, num_grasps=num_grasps, metas=metas,
                                                               return_intermediate=return_intermediate, **kwargs)
        out = self._finish(final_grasps, batch_pcs, metas, num_grasps)
        out["all_steps_grasps"] = []
        if step_grasps:                                          # tools/inference.py:629-641
            if batch_pcs.shape[0] > 1:
                raise NotImplementedError("Batched grasps for all diffusion steps are not implemented")
            gm, gs = metas["grasp_mean"].reshape(-1, 6), metas["grasp_std"].reshape(-1, 6)
            for step_tmrp, step_logit in step_grasps:
                _, H, _ = engine.pose_postprocess(step_tmrp.to(self.device), step_logit.to(self.device), gm, gs,
                                                  grasps_per_obj=num_grasps)
                out["all_steps_grasps"].append(H)
        return out


class InferenceVAE(_InferenceBase):
    """tools/inference.py:770-815."""

    def generate_grasps(self, pc, metas, num_grasps=10, return_intermediate=False, **kwargs):
        batch_pcs = (pc.unsqueeze(0) if pc.ndim == 2 else pc).to(self.device, non_blocking=True)
        metas = {k: v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v for k, v in metas.items()}
        final_grasps = self.model.generate_grasps(batch_pcs, num_grasps, **kwargs)
        return self._finish(final_grasps, batch_pcs, metas, num_grasps)


def default_metas(n_pc=1):
    """Synthetic normalisation statistics of SURVEY.md section 8d (object scale 0.05 m)."""
    return dict(pc_mean=torch.zeros(n_pc, 3), pc_std=torch.full((n_pc, 3), 0.05), grasp_mean=torch.zeros(1, 6),
                grasp_std=torch.tensor([[.05, .05, .05, .5, .5, .5]]), dataset_normalized=True)
