"""PVCNNEncoder (R/grasp_ldm/models/modules/pc_encoders.py:8-136) on the sm_100a kernels."""
import torch
from torch import nn

from . import engine
from .pvcnn import PVCNN


class PVCNNEncoder(nn.Module):
    def __init__(self, in_features=3, out_features=32, n_points=1024, extra_feature_channels=0, scale_channels=0.25,
                 scale_voxel_resolution=0.75, num_blocks=(1, 1, 1, 1), is_conditioned=False, cond_dims=None,
                 extra_block_channels=None, use_global_attention=False, out_channels=1, load_from_ckpt_path=None) -> None:
        super().__init__()
        if use_global_attention:
            raise NotImplementedError("use_global_attention=False in both generation configs")
        self.pvcnn_modules = PVCNN(extra_feature_channels=extra_feature_channels, scale_channels=scale_channels,
                                   scale_voxel_resolution=scale_voxel_resolution, num_blocks=num_blocks,
                                   is_conditioned=is_conditioned, cond_dims=cond_dims,
                                   extra_block_channels=extra_block_channels)
        self.in_features, self.out_features = in_features, out_features
        mid = int(self.pvcnn_modules.out_channels / 2)
        self.conv_downscale = nn.Conv1d(self.pvcnn_modules.out_channels, mid, kernel_size=1)
        self.global_attention = None
        self.out_layer = nn.Sequential(nn.Conv1d(mid, out_channels, kernel_size=1), nn.Linear(n_points, out_features))
        self.precision = "fp32"     # "bf16": point-wise layers on the tcgen05 tensor cores (bf16 operands, fp32 accumulate)
        if load_from_ckpt_path is not None:
            ckpt = torch.load(load_from_ckpt_path)
            self.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)

    @torch.no_grad()
    def forward(self, out, cond=None):
        """xyz [B,N,3] -> [B,C_out,out_features] ([B,out_features] when C_out == 1)"""
        if self.training:
            raise NotImplementedError("generation path: call .eval() (BatchNorm uses running statistics)")
        return engine.encoder_forward(self, out, precision=self.precision)
