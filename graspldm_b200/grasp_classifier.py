"""PointsBasedGraspClassifier (R/grasp_ldm/models/grasp_classifier.py:13-143) - the grasp-success scorer that follows
generation: object cloud + gripper points -> PVCNN / PVCNN2 point features -> per-point MLP
-> Linear over the points -> sigmoid.  Same constructor arguments, module tree and state_dict keys; `classify_grasps` /
`forward(..., compute_loss=False)` run on the CUDA kernels (point-voxel blocks, shared MLPs, the point reduction), the
training loss is outside the path."""
import torch
from torch import Tensor, nn

from . import _lib, engine
from .pvcnn import PVCNN, PVCNN2, SharedMLP


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class PointsBasedGraspClassifier(nn.Module):
    SUPPORTED_BASE_NETWORKS = {"PVCNN": PVCNN, "PVCNN2": PVCNN2}

    def __init__(self, num_pc_points, points_backbone_config: dict, loss_config: dict = None):
        super().__init__()
        self._loss_config = loss_config            # BCE-with-logits carries no parameters; training-only
        self.num_pc_points = num_pc_points
        self.base_network = self.SUPPORTED_BASE_NETWORKS[_get(points_backbone_config, "type")](
            **dict(_get(points_backbone_config, "args")))
        self._cls_out_dim, self._width_multiplier = 1, 1
        # create_mlp_components(in, [128, 0.5, 1], classifier=True, dim=2)   ext/pvcnn/utils.py:30-62
        self.classifier = nn.Sequential(SharedMLP(self.base_network.out_channels, 128), nn.Dropout(0.5), nn.Conv1d(128, 1, 1),
                                        nn.Linear(self.num_pc_points, 1))
        self.sigmoid = nn.Sigmoid()

    @property
    def _type(self) -> str:
        return self.__class__.__name__

    @torch.no_grad()
    def forward(self, pc: Tensor, grasp_points: Tensor, *, cls_target: Tensor = None, compute_loss: bool = True):
        """pc [B,Np,3], grasp_points [B,Ng,3] (Np + Ng = num_pc_points) -> (None, success probability [B])"""
        if compute_loss:
            raise NotImplementedError("the classification loss is training-only; call with compute_loss=False or classify_grasps")
        if self.training:
            raise NotImplementedError("generation path: call .eval()")
        engine._require_cuda(pc, "pc")
        obj = torch.cat((pc, torch.zeros_like(pc[..., :1])), dim=-1)                 # feature label: 0 = object point
        grp = torch.cat((grasp_points, torch.ones_like(grasp_points[..., :1])), dim=-1)    # 1 = gripper point
        pc_in = torch.cat((obj, grp), dim=-2).transpose(1, 2).contiguous().float()   # [B, 4, N]
        x = self.base_network(pc_in)                                                  # [B, C, N]
        x = self.classifier[0](x)                                                     # SharedMLP -> [B, 128, N]
        conv, lin = self.classifier[2], self.classifier[3]
        dev = x.device
        B, _, N = x.shape
        if N != lin.in_features:
            raise RuntimeError(f"expected {lin.in_features} points (object + gripper), got {N}")
        with torch.cuda.device(dev):
            w = conv.weight.detach().reshape(1, conv.in_channels).float().contiguous()
            h = engine._pw(x, w, None, conv.bias.detach().float().contiguous(), None, 0)      # [B, 1, N]
            logit = torch.empty((B, 1, 1), device=dev, dtype=torch.float32)
            _lib.call("gldm_linear_lastdim_f32", h.data_ptr(), lin.weight.detach().float().contiguous().data_ptr(),
                      lin.bias.detach().float().contiguous().data_ptr(), B, N, 1, logit.data_ptr(), engine._stream(dev))
        return None, torch.sigmoid(logit.squeeze())

    def classify_grasps(self, pc: Tensor, grasp_pose: Tensor) -> Tensor:
        _, preds = self.forward(pc, grasp_pose, compute_loss=False)
        return preds
