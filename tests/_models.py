"""Seeded model construction shared by tests / bench / smoke (no reference import: the product's own
configs.build_ldm reproduces the reference's random init bit for bit, see test_state_dict_manifest)."""
import torch

from graspldm_b200 import configs


def build(name="fpc", scheduler="ddpm", seed=0, device=None):
    torch.manual_seed(seed)
    m = configs.build_ldm(name, scheduler)
    return m.to(device) if device is not None else m


def split_state_dicts(model):
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    vae = {k[len("vae_model."):]: v for k, v in sd.items() if k.startswith("vae_model.")}
    ddm = {k: v for k, v in sd.items() if k.startswith("diffusion_model.")}
    return vae, ddm
