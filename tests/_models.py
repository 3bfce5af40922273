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


def trained_like_(model, seed=1):
    """Overwrite every normalisation parameter and buffer with checkpoint-like values (a seeded random init leaves
    BatchNorm running statistics at 0 / 1, every norm weight at 1 and bias at 0, LayerNorm.g at 1, so a kernel that
    ignored them would still pass).  Keyed by module NAME, so the reference module tree (tests/golden/make_golden.py)
    and the product's mirror get bit-identical values: BatchNorm running_mean ~ N(0, .3), running_var ~ U(.5, 2);
    every norm weight / LayerNorm.g ~ U(.5, 1.5); norm bias ~ N(0, .2)."""
    import zlib
    from torch import nn
    with torch.no_grad():
        for name, mod in model.named_modules():
            g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
            u = lambda t, lo, hi: t.copy_(lo + (hi - lo) * torch.rand(t.shape, generator=g))
            n = lambda t, s: t.copy_(torch.randn(t.shape, generator=g) * s)
            if isinstance(mod, nn.modules.batchnorm._BatchNorm):
                n(mod.running_mean, 0.3)
                u(mod.running_var, 0.5, 2.0)
                u(mod.weight, 0.5, 1.5)
                n(mod.bias, 0.2)
            elif isinstance(mod, (nn.GroupNorm, nn.LayerNorm)):
                u(mod.weight, 0.5, 1.5)
                n(mod.bias, 0.2)
            elif isinstance(getattr(mod, "g", None), nn.Parameter):      # resnets.py:104-113 LayerNorm (gain only)
                u(mod.g, 0.5, 1.5)
    return model


def build_trained_like(name="fpc", scheduler="ddpm", seed=0, device=None):
    m = trained_like_(build(name, scheduler, seed))
    return m.to(device) if device is not None else m
