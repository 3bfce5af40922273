"""ResNet1D / TimeConditionedResNet1D with the reference's constructor kwargs and state_dict keys
(R/grasp_ldm/models/modules/resnets.py:263-616), executed by the fused CUDA kernels.

The nn.Module tree below exists to OWN parameters under the reference's key names (checkpoints load
with strict=True) and to reproduce the reference's seeded random init (same layer types built in the
same order).  It contains no PyTorch arithmetic: `forward` hands the packed weights to the persistent
kernel through graspldm_b200.engine.  Sub-blocks are parameter holders only.
"""
from typing import Sequence

import torch
from torch import nn

from . import engine


class _Holder(nn.Module):
    """Parameter container; the arithmetic lives in csrc/resnet1d_*.cu."""

    def forward(self, *a, **k):
        raise NotImplementedError(
            f"{type(self).__name__} is a parameter holder; run the enclosing ResNet1D / "
            "TimeConditionedResNet1D, which executes the whole network in one CUDA kernel")


class RandomOrLearnedSinusoidalPosEmb(_Holder):   # resnets.py:44-56
    def __init__(self, dim, is_random=False):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2), requires_grad=not is_random)


class LayerNorm(_Holder):                          # resnets.py:104-113
    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1))


class Block(_Holder):                              # resnets.py:127-177 (proj is weight-standardised in-kernel)
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.proj = nn.Conv1d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)


class ResnetBlock(_Holder):                        # resnets.py:180-208
    def __init__(self, dim, dim_out, *, emb_dim=None, groups=8):
        super().__init__()
        self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(emb_dim, dim_out * 2)) if emb_dim is not None else None
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        if dim != dim_out:
            raise NotImplementedError("res_conv (dim != dim_out) does not occur on the generation path")
        self.res_conv = nn.Identity()


class LinearAttention(_Holder):                    # resnets.py:211-235
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        hidden = heads * dim_head
        self.to_qkv = nn.Conv1d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv1d(hidden, dim, 1), LayerNorm(dim))


class PreNorm(_Holder):                            # resnets.py:116-124
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = LayerNorm(dim)


class Residual(_Holder):                           # resnets.py:59-65
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


def _stage(dim_in, dim_out, emb_dim, groups):
    return nn.ModuleList([
        ResnetBlock(dim_in, dim_in, emb_dim=emb_dim, groups=groups),
        ResnetBlock(dim_in, dim_in, emb_dim=emb_dim, groups=groups),
        Residual(PreNorm(dim_in, LinearAttention(dim_in))),
        nn.Conv1d(dim_in, dim_out, 3, padding=1),
    ])


class _ResNetBase(nn.Module):
    is_time_conditioned = False

    def _build_trunk(self, dim, block_channels, groups, emb_dim):
        dims = (dim,) + tuple(block_channels)
        self.blocks = nn.ModuleList([_stage(a, b, emb_dim, groups) for a, b in zip(dims[:-1], dims[1:])])
        self.final_res_block = ResnetBlock(dims[-1], dims[-1], emb_dim=emb_dim, groups=groups)
        self.final_conv = nn.Conv1d(dims[-1], self.out_channels, 1)
        self._dims = dims
        self._groups = groups

    def _check_supported(self, is_self_conditioned, learned_variance, input_conditioning_dims, out_channels):
        if is_self_conditioned or learned_variance:
            raise NotImplementedError("self conditioning / learned variance are not on the generation path")
        if input_conditioning_dims is None:
            raise NotImplementedError("the generation path is always conditioned on the point-cloud latent")
        if out_channels not in (None, 1):
            raise NotImplementedError("out_channels must be 1")

    # -- engine plumbing ---------------------------------------------------------------------------
    def kernel_cfg(self, seq_len, cond_ch):
        return engine.make_resnet_cfg(L=seq_len, dims=self._dims, emb_dim=self.emb_dim, cond_ch=cond_ch,
                                      cond_dim=self.input_emb_layers[0].in_features, groups=self._groups,
                                      time_cond=self.is_time_conditioned,
                                      fourier_half=(self.time_mlp[0].weights.numel() if self.is_time_conditioned else 0))

    def packed(self, seq_len, cond_ch):
        """Prepared device weights for (seq_len, cond_ch); rebuilt when parameters change."""
        return engine.packed_resnet(self, seq_len, cond_ch)


class ResNet1D(_ResNetBase):
    """resnets.py:263-424.  forward(x [B,1,D], *, z_cond [B,C,Dc]) -> [B,1,D]"""

    def __init__(self, dim: int, init_dim: int = None, out_channels: int = None,
                 block_channels: Sequence = (16, 64, 128, 64, 16), channels: int = 1,
                 input_conditioning_dims: int = None, is_self_conditioned: bool = False,
                 resnet_block_groups: int = 8, learned_variance: bool = False, dropout=None) -> None:
        super().__init__()
        self._check_supported(is_self_conditioned, learned_variance, input_conditioning_dims, out_channels)
        if channels != 1 or init_dim not in (None, dim):
            raise NotImplementedError("channels=1 and init_dim=dim only")
        self.channels, self.is_self_conditioned = channels, False
        self.in_features = self.out_features = dim
        self.init_conv = nn.Conv1d(1, dim, 7, padding=3)
        self.dropout = nn.Dropout(p=dropout, inplace=True) if dropout is not None else None   # identity in eval
        self.emb_dim = dim * 4
        self.is_input_conditioned = True
        self.input_emb_layers = nn.Sequential(nn.Linear(input_conditioning_dims, self.emb_dim), nn.SiLU())
        self.out_channels = 1
        self._build_trunk(dim, block_channels, resnet_block_groups, self.emb_dim)

    @torch.no_grad()
    def forward(self, x, *, z_cond=None, x_self_cond=None):
        return engine.resnet_forward(self, x, None, z_cond)


class TimeConditionedResNet1D(_ResNetBase):
    """resnets.py:427-616.  forward(x [B,1,D], *, time int64[B], z_cond [B,C,Dc]) -> eps [B,1,D]"""
    is_time_conditioned = True

    def __init__(self, dim: int, init_dim: int = None, out_channels: int = None,
                 block_channels: Sequence = (16, 64, 128, 64, 16), channels: int = 1,
                 input_conditioning_dims: int = None, is_self_conditioned: bool = False,
                 resnet_block_groups: int = 8, learned_variance: bool = False, dropout=None,
                 is_time_conditioned: bool = True, learned_sinusoidal_cond: bool = False,
                 random_fourier_features: bool = False, learned_sinusoidal_dim: int = 16) -> None:
        super().__init__()
        self._check_supported(is_self_conditioned, learned_variance, input_conditioning_dims, out_channels)
        if channels != 1 or init_dim not in (None, dim):
            raise NotImplementedError("channels=1 and init_dim=dim only")
        if not is_time_conditioned or not (learned_sinusoidal_cond or random_fourier_features):
            raise NotImplementedError("the denoiser kernel implements the random/learned Fourier time embedding")
        self.channels, self.is_self_conditioned = channels, False
        self.in_features = self.out_features = dim
        self.init_conv = nn.Conv1d(1, dim, 7, padding=3)
        self.dropout = nn.Dropout(p=dropout, inplace=True) if dropout is not None else None
        self.emb_dim = dim * 4
        self.random_or_learned_sinusoidal_cond = True
        self.time_mlp = nn.Sequential(
            RandomOrLearnedSinusoidalPosEmb(learned_sinusoidal_dim, random_fourier_features),
            nn.Linear(learned_sinusoidal_dim + 1, self.emb_dim), nn.GELU(), nn.Linear(self.emb_dim, self.emb_dim))
        self.is_input_conditioned = True
        self.input_emb_layers = nn.Sequential(nn.Linear(input_conditioning_dims, self.emb_dim), nn.SiLU())
        self.out_channels = 1
        self._build_trunk(dim, block_channels, resnet_block_groups, self.emb_dim)

    @torch.no_grad()
    def forward(self, x, *, time=None, z_cond=None, x_self_cond=None, **kwargs):
        assert time is not None
        return engine.resnet_forward(self, x, time, z_cond, precision=kwargs.get("precision", "fp32"))


class ClassTimeConditionedResNet1D(TimeConditionedResNet1D):
    """class_conditioned_resnet.py:9-122: the time-conditioned denoiser plus a class embedding
    cls_embed = SiLU(Linear(1 -> emb)) that is added to the time embedding (:96-98).  forward(x, *, time, z_cond,
    cls_cond [B,1]) - without cls_cond the class is read from kwargs["metas"]["mode_cls"], as the reference does."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.cls_embed = nn.Sequential(nn.Linear(1, self.emb_dim), nn.SiLU())

    @staticmethod
    def class_condition(cls_cond, kwargs, dtype=torch.float32):
        if cls_cond is None:
            assert "metas" in kwargs and "mode_cls" in kwargs["metas"], "Class conditioning tensor is required"
            cls_cond = kwargs["metas"]["mode_cls"].unsqueeze(-1).reshape(-1, 1).to(dtype=dtype)
        return cls_cond

    @torch.no_grad()
    def forward(self, x, *, time=None, z_cond=None, x_self_cond=None, cls_cond=None, **kwargs):
        assert time is not None
        cls_cond = self.class_condition(cls_cond, kwargs, x.dtype)
        return engine.resnet_forward(self, x, time, z_cond, precision=kwargs.get("precision", "fp32"), cls_cond=cls_cond)
