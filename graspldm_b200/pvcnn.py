"""PVCNN / PointNet++ building blocks with the reference's names, kwargs and state_dict keys
(R = /root/reference/grasp_ldm/models/modules/ext/pvcnn): modules/{pvconv,shared_mlp,se,voxelization,
ball_query,pointnet}.py, utils.py and pvcnn_base.py.

Dense members (Conv3d, Conv1d, GroupNorm, BatchNorm, Linear) are torch layer objects used as parameter
holders so that seeded init and checkpoints match the reference; the encoder's arithmetic runs in
csrc/encoder_*.cu and csrc/point_ops.cu through graspldm_b200.engine.  The set-abstraction modules
(FPS / ball query / grouping - PointNet++/PVCNN2 side of the extension, not used by the named configs,
SURVEY.md finding 1) run their point operators on our kernels and their shared MLP on the fp32 GEMM.
"""
import torch
from torch import nn

from . import engine, functional as F


class Swish(nn.Module):                   # R/../modules.py:5-7 (parameter-free marker; fused into the GroupNorm kernel)
    def forward(self, x):
        raise NotImplementedError("fused into gldm_groupnorm_swish_f32")


class SE3d(nn.Module):                    # modules/se.py:12-25
    def __init__(self, channel, reduction=8, use_relu=False):
        super().__init__()
        if use_relu:
            raise NotImplementedError("with_se_relu is not used by the generation configs")
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), Swish(),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())


class SharedMLP(nn.Module):               # modules/shared_mlp.py:6-35
    def __init__(self, in_channels, out_channels, dim=1):
        super().__init__()
        conv, bn = (nn.Conv1d, nn.BatchNorm1d) if dim == 1 else (nn.Conv2d, nn.BatchNorm2d)
        if dim not in (1, 2):
            raise ValueError
        if not isinstance(out_channels, (list, tuple)):
            out_channels = [out_channels]
        layers = []
        for oc in out_channels:
            layers.extend([conv(in_channels, oc, 1), bn(oc), nn.ReLU(True)])
            in_channels = oc
        self.layers = nn.Sequential(*layers)

    @torch.no_grad()
    def forward(self, inputs):
        """Conv(k=1) + BatchNorm(eval) + ReLU per layer on [B,C,N] or [B,C,M,U] via the fp32 GEMM kernel."""
        if isinstance(inputs, (list, tuple)):
            return (self.forward(inputs[0]), *inputs[1:])
        if self.training:
            raise NotImplementedError("generation path: call .eval() (BatchNorm uses running statistics)")
        x = inputs
        shp = x.shape
        x = x.reshape(shp[0], shp[1], -1).contiguous().float()
        for i in range(0, len(self.layers), 3):
            conv, bn = self.layers[i], self.layers[i + 1]
            sc, sh = engine._fold_bn(conv, bn)
            w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float().contiguous()
            with torch.cuda.device(x.device):
                x = engine._pw(x, w, sc, sh, None, 1)
        return x.reshape(shp[0], x.shape[1], *shp[2:])


class Voxelization(nn.Module):            # modules/voxelization.py:9-35
    def __init__(self, resolution, normalize=True, eps=0):
        super().__init__()
        self.r = int(resolution)
        self.normalize = normalize
        self.eps = eps

    @torch.no_grad()
    def forward(self, features, coords):
        if self.normalize:
            raise NotImplementedError("normalize=True is not used by PVCNN (pvcnn_base.py:49-56 passes False)")
        grid, norm = F.voxelize_fused(features, coords, self.r)
        return grid, norm

    def extra_repr(self):
        return f"resolution={self.r}"


class PVConv(nn.Module):                  # modules/pvconv.py:13-84
    def __init__(self, in_channels, out_channels, kernel_size, resolution, use_attention=False, dropout=0.1,
                 with_se=False, with_se_relu=False, normalize=True, eps=0):
        super().__init__()
        if use_attention or kernel_size != 3 or not with_se or normalize:
            raise NotImplementedError("PVConv is implemented as configured by PVCNN: k=3, SE, no attention, normalize=False")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.resolution = kernel_size, resolution
        self.voxelization = Voxelization(resolution, normalize=normalize, eps=eps)
        layers = [nn.Conv3d(in_channels, out_channels, kernel_size, stride=1, padding=kernel_size // 2),
                  nn.GroupNorm(num_groups=8, num_channels=out_channels), Swish()]
        layers += [nn.Dropout(dropout)] if dropout is not None else []
        layers += [nn.Conv3d(out_channels, out_channels, kernel_size, stride=1, padding=kernel_size // 2),
                   nn.GroupNorm(num_groups=8, num_channels=out_channels), Swish()]
        layers.append(SE3d(out_channels, use_relu=with_se_relu))
        self.voxel_layers = nn.Sequential(*layers)
        self.point_features = SharedMLP(in_channels, out_channels)


class BallQuery(nn.Module):               # modules/ball_query.py:9-34
    def __init__(self, radius, num_neighbors, include_coordinates=True):
        super().__init__()
        self.radius, self.num_neighbors, self.include_coordinates = radius, num_neighbors, include_coordinates

    @torch.no_grad()
    def forward(self, points_coords, centers_coords, points_features=None):
        points_coords = points_coords.contiguous()
        centers_coords = centers_coords.contiguous()
        idx = F.ball_query(centers_coords, points_coords, self.radius, self.num_neighbors)
        nb_coords = F.grouping(points_coords, idx) - centers_coords.unsqueeze(-1)
        if points_features is None:
            assert self.include_coordinates, "No Features For Grouping"
            return nb_coords
        nb = F.grouping(points_features, idx)
        return torch.cat([nb_coords, nb], dim=1) if self.include_coordinates else nb


class PointNetSAModule(nn.Module):        # modules/pointnet.py:53-114
    def __init__(self, num_centers, radius, num_neighbors, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        if not isinstance(radius, (list, tuple)):
            radius = [radius]
        if not isinstance(num_neighbors, (list, tuple)):
            num_neighbors = [num_neighbors] * len(radius)
        if not isinstance(out_channels, (list, tuple)):
            out_channels = [[out_channels]] * len(radius)
        elif not isinstance(out_channels[0], (list, tuple)):
            out_channels = [out_channels] * len(radius)
        groupers, mlps, total = [], [], 0
        for r, oc, k in zip(radius, out_channels, num_neighbors):
            groupers.append(BallQuery(radius=r, num_neighbors=k, include_coordinates=include_coordinates))
            mlps.append(SharedMLP(in_channels + (3 if include_coordinates else 0), oc, dim=2))
            total += oc[-1]
        self.num_centers, self.out_channels = num_centers, total
        self.groupers, self.mlps = nn.ModuleList(groupers), nn.ModuleList(mlps)

    @torch.no_grad()
    def forward(self, inputs):
        features, coords = inputs
        centers = F.furthest_point_sample(coords, self.num_centers)
        outs = [mlp(g(coords, centers, features)).max(dim=-1).values for g, mlp in zip(self.groupers, self.mlps)]
        return (torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]), centers


class PointNetFPModule(nn.Module):        # modules/pointnet.py:117-135
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.mlp = SharedMLP(in_channels=in_channels, out_channels=out_channels, dim=1)

    @torch.no_grad()
    def forward(self, inputs):
        if len(inputs) == 3:
            points_coords, centers_coords, centers_features = inputs
            points_features = None
        else:
            points_coords, centers_coords, centers_features, points_features = inputs
        interp = F.nearest_neighbor_interpolate(points_coords, centers_coords, centers_features)
        if points_features is not None:
            interp = torch.cat([interp, points_features], dim=1)
        return self.mlp(interp), points_coords


class PVCNN(nn.Module):                   # pvcnn_base.py:15-140 (unconditioned, as PVCNNEncoder builds it)
    def __init__(self, in_channels=3, extra_feature_channels=0, scale_channels=0.25, scale_voxel_resolution=0.75,
                 num_blocks=(1, 2, 1, 1), is_conditioned=False, cond_dims=None, extra_block_channels=None):
        super().__init__()
        if is_conditioned or extra_block_channels is not None:
            raise NotImplementedError("conditioned PVCNN / extra blocks are not used by the generation configs")
        if len(num_blocks) != 4:
            raise ValueError("PVCNN is configured with 4 blocks")
        self.in_channels = in_channels + extra_feature_channels
        c = [int(64 * scale_channels), int(128 * scale_channels), int(1024 * scale_channels), int(2048 * scale_channels)]
        r = [int(32 * scale_voxel_resolution), int(16 * scale_voxel_resolution), None, None]
        assert all(x % 2 == 0 for x in c) and r[0] % 2 == 0 and r[1] % 2 == 0
        self.block_spec = tuple((c[i], num_blocks[i], r[i]) for i in range(4))
        self.out_channels = c[3]
        layers, cin = [], self.in_channels
        for oc, nb, res in self.block_spec:       # utils.py:65-94 create_pointnet_components
            for _ in range(nb):
                if res is None:
                    layers.append(SharedMLP(cin, oc))
                else:
                    layers.append(PVConv(cin, oc, kernel_size=3, resolution=res, with_se=True, normalize=False, eps=0))
                cin = oc
        self.point_features = nn.ModuleList(layers)
        self.is_conditioned = False
