"""GraspCVAE and its sub-networks (R/grasp_ldm/models/grasp_vae.py) - generation side.

Constructor kwargs, attribute names and state_dict keys follow the reference so VAE checkpoints load with
strict=True.  The grasp *encoder*, bottleneck and losses are training-only: their
parameters are kept (same keys, same init order), their forward passes are not implemented here.
"""
from typing import Tuple, Union

import torch
from torch import Tensor, nn

from . import engine
from .pc_encoders import PVCNNEncoder
from .resnets import ResNet1D


def _cfg_get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class ConditionalGraspPoseDecoder(nn.Module):      # grasp_vae.py:353-436
    MODELS = {"ResNet1D": ResNet1D}

    def __init__(self, config, in_features, feature_resolution, num_output_qualities=None) -> None:
        super().__init__()
        ctype = _cfg_get(config, "type")
        if ctype not in self.MODELS:
            raise NotImplementedError(f"Base network arch of type=`{ctype}` is not implemented. Available: {list(self.MODELS)}")
        if num_output_qualities:
            raise NotImplementedError("num_output_qualities > 0 is not used by the generation configs")
        self.in_features, self.feature_resolution = in_features, feature_resolution
        self.in_layer = nn.Linear(in_features, feature_resolution)
        self.net = self.MODELS[ctype](dim=feature_resolution, **dict(_cfg_get(config, "args")))
        self.tmrp = nn.Linear(self.net.out_features, 6)
        self.class_logits = nn.Linear(self.net.out_features, 1)
        self._use_qualities = False
        self.num_qualities = None
        self.out_features = (6, 1)
        self.precision = "fp32"     # "bf16": trunk GEMMs on the tcgen05 tensor cores (bf16 operands, fp32 accumulate)

    @torch.no_grad()
    def forward(self, z_h: Tensor, cond: Tensor = None, *, grasps_per_object: int = 1) -> Tuple[Tensor, Tensor]:
        """z_h [B,D], cond [B,C,Dc] (or one row per object with grasps_per_object=G) -> (tmrp [B,6], logits [B,1]);
        in_layer, the ResNet1D trunk and both heads run in one kernel launch."""
        return engine.decoder_forward(self, z_h, cond, grasps_per_object, precision=self.precision)


class ConditionalGraspPoseEncoder(nn.Module):      # grasp_vae.py:439-536 (training only: parameters kept)
    def __init__(self, config, latent_size: int, feature_resolution: int = 16) -> None:
        super().__init__()
        args = dict(_cfg_get(config, "args"))
        self.in_features = args.pop("in_features")
        self.out_features = latent_size
        self.feature_resolution = feature_resolution
        self.in_layer = nn.Linear(self.in_features, feature_resolution)
        self.net = ResNet1D(dim=feature_resolution, **args)
        self.out_layer = nn.Linear(self.net.out_features, latent_size)

    def forward(self, x, cond):
        raise NotImplementedError("grasp encoding is training-only; outside the generation path")


class VAEBottleneck(nn.Module):                    # grasp_vae.py:539-574 (training only)
    def __init__(self, in_features: int, latent_size: int) -> None:
        super().__init__()
        self.mu = nn.Linear(in_features, latent_size)
        self.logvar = nn.Linear(in_features, latent_size)

    def forward(self, z):
        raise NotImplementedError("the VAE bottleneck is training-only; outside the generation path")


class PcConditionedGraspEncoder(nn.Module):        # grasp_vae.py:258-350
    PC_ENCODERS = {"PVCNNEncoder": PVCNNEncoder}

    def __init__(self, pc_encoder_config, grasp_encoder_config, pc_latent_size: int = 64, grasp_latent_size: int = 4) -> None:
        super().__init__()
        ptype = _cfg_get(pc_encoder_config, "type")
        if ptype not in self.PC_ENCODERS:
            raise NotImplementedError(f"Pointcloud encoder of type=`{ptype}` is not implemented. Available: {list(self.PC_ENCODERS)}")
        self.pc_encoder = self.PC_ENCODERS[ptype](out_features=pc_latent_size, **dict(_cfg_get(pc_encoder_config, "args")))
        self.grasp_encoder = ConditionalGraspPoseEncoder(config=grasp_encoder_config, latent_size=grasp_latent_size)
        self.out_features = grasp_latent_size

    def forward(self, xyz, h, z_pc=None):
        raise NotImplementedError("joint (pc, grasp) encoding is training-only; use encode_pc")

    def encode_pc(self, xyz: Tensor) -> Tensor:
        return self.pc_encoder(xyz)

    def get_conditioning_latent(self, xyz: Tensor) -> Tensor:
        return self.encode_pc(xyz)


class GraspCVAE(nn.Module):                        # grasp_vae.py:17-255
    def __init__(self, grasp_latent_size: int, pc_latent_size: int, grasp_encoder_config: dict, pc_encoder_config: dict,
                 decoder_config: dict, loss_config: dict = None, intermediate_feature_resolution: int = 16,
                 num_output_qualities: Union[int, None] = None) -> None:
        super().__init__()
        self.grasp_latent_size, self.pc_latent_size = grasp_latent_size, pc_latent_size
        self.loss_config = loss_config            # losses carry no parameters; not built (training-only)
        self.encoder = PcConditionedGraspEncoder(pc_encoder_config=pc_encoder_config,
                                                 grasp_encoder_config=grasp_encoder_config,
                                                 pc_latent_size=pc_latent_size, grasp_latent_size=grasp_latent_size)
        self.bottleneck = VAEBottleneck(in_features=self.encoder.out_features, latent_size=grasp_latent_size)
        self.num_output_qualities = num_output_qualities
        self.decoder = ConditionalGraspPoseDecoder(in_features=grasp_latent_size, config=decoder_config,
                                                   num_output_qualities=num_output_qualities,
                                                   feature_resolution=intermediate_feature_resolution)
        self.out_features = self.decoder.out_features

    @property
    def _type(self) -> str:
        return self.__class__.__name__

    @property
    def use_grasp_qualities(self) -> bool:
        return bool(self.decoder._use_qualities)

    def encode(self, xyz, grasp):
        raise NotImplementedError("GraspCVAE.encode is training-only; outside the generation path")

    def forward(self, *a, **k):
        raise NotImplementedError("GraspCVAE.forward (training loss) is outside the generation path")

    def encode_pc(self, xyz: Tensor) -> Tensor:
        return self.encoder.encode_pc(xyz)

    def sample_grasp_latent(self, batch_size: int, device) -> Tensor:
        return torch.randn(batch_size, self.grasp_latent_size, pin_memory=True).to(device, non_blocking=True)

    @torch.no_grad()
    def generate_grasps(self, xyz: Tensor, num_grasps: int = 10, *, z_h: Tensor = None):
        """grasp_vae.py:226-255.  xyz [B,N,3] -> (tmrp [B*G,6], logits [B*G,1]).  z_h may be injected for parity;
        by default it is drawn on the CPU generator and moved, exactly as the reference does (:250)."""
        assert xyz.ndim == 3, f"Input pointcloud should be 3-dim tensor of shape [B, N, 3]. Found a {xyz.ndim} dimensional tensor."
        num_pcs = xyz.shape[0]
        z_pc = self.encode_pc(xyz)                                   # one row per object; never repeated in HBM
        if z_h is None:
            z_h = torch.randn(num_pcs * num_grasps, self.grasp_latent_size, pin_memory=True).to(xyz.device, non_blocking=True)
        return self.decoder(z_h, z_pc, grasps_per_object=num_grasps)
