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
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(True) if use_relu else Swish(),
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
        if self.normalize:           # voxelization.py:19-30: scale by twice the largest point norm of the cloud
            nc = coords - coords.mean(2, keepdim=True)
            nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + self.eps) + 0.5
            norm = torch.clamp(nc * self.r, 0, self.r - 1)
            return F.avg_voxelize(features, torch.round(norm).to(torch.int32), self.r), norm
        grid, norm = F.voxelize_fused(features, coords, self.r)
        return grid, norm

    def extra_repr(self):
        return f"resolution={self.r}"


class PVConv(nn.Module):                  # modules/pvconv.py:13-84
    def __init__(self, in_channels, out_channels, kernel_size, resolution, use_attention=False, dropout=0.1,
                 with_se=False, with_se_relu=False, normalize=True, eps=0):
        super().__init__()
        if use_attention or kernel_size != 3 or not with_se:
            raise NotImplementedError("PVConv is implemented as PVCNN / PVCNN2 configure it: k=3, SE, no voxel attention")
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

    @torch.no_grad()
    def forward(self, inputs):
        """(features [B,C,N], coords [B,3,N]) -> (fused [B,C_out,N], coords)   pvconv.py:76-84, strict-fp32 kernels"""
        features, coords = inputs
        if self.training:
            raise NotImplementedError("generation path: call .eval()")
        return engine.pvconv_forward_f32(engine.packed_block(self), features, coords), coords


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


class PointNetAModule(nn.Module):         # modules/pointnet.py:11-50 (global abstraction: MLP over all points, max)
    def __init__(self, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        if not isinstance(out_channels, (list, tuple)):
            out_channels = [[out_channels]]
        elif not isinstance(out_channels[0], (list, tuple)):
            out_channels = [out_channels]
        mlps, total = [], 0
        for oc in out_channels:
            mlps.append(SharedMLP(in_channels + (3 if include_coordinates else 0), oc, dim=1))
            total += oc[-1]
        self.include_coordinates, self.out_channels = include_coordinates, total
        self.mlps = nn.ModuleList(mlps)

    @torch.no_grad()
    def forward(self, inputs):
        features, coords = inputs
        if self.include_coordinates:
            features = torch.cat([features, coords], dim=1)
        zero = torch.zeros((coords.size(0), 3, 1), device=coords.device)
        outs = [mlp(features).max(dim=-1, keepdim=True).values for mlp in self.mlps]
        return (torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]), zero


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
        """(features [B,C,N] | None, coords [B,3,N]) -> ([B, sum C_out, M], centres [B,3,M]).  FPS and ball query are the
        bit-exact operator kernels; grouping, the shared MLP and the max over the neighbours are ONE kernel per radius
        (csrc/set_abstraction.cu): the grouped tensor [B, C+3, M, U] of the reference never reaches HBM."""
        coords = coords.contiguous()
        centers = F.furthest_point_sample(coords, self.num_centers)
        outs = []
        for g, mlp in zip(self.groupers, self.mlps):
            if mlp.training:
                raise NotImplementedError("generation path: call .eval() (BatchNorm uses running statistics)")
            idx = F.ball_query(centers, coords, g.radius, g.num_neighbors)
            outs.append(engine.sa_group_mlp_max(mlp, coords, centers, features, idx, g.include_coordinates))
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

    @torch.no_grad()
    def forward(self, inputs, *, cond=None):
        """[B, 3+C, N] -> [B, C_out, N]   pvcnn_base.py:114-140 (strict-fp32 kernels, block by block)"""
        features = inputs[:, :self.in_channels, :].contiguous()
        coords = features[:, :3, :].contiguous()
        for blk in self.point_features:
            features = blk((features, coords))
            features = features[0] if isinstance(features, tuple) else features
        return features


def _seq_or_single(blocks):
    return blocks[0] if len(blocks) == 1 else nn.Sequential(*blocks)


def create_pointnet2_sa_components(sa_blocks, extra_feature_channels, with_se=False, voxelization_normalize=True, eps=0,
                                   dropout=0.1, width_multiplier=1, voxel_resolution_multiplier=1):
    """utils.py:97-185 (embed_dim = 0, no voxel attention): same modules in the same order -> same seeded init / keys."""
    r, vr = width_multiplier, voxel_resolution_multiplier
    in_channels = extra_feature_channels + 3
    sa_layers, sa_in_channels = [], []
    num_centers = None
    for c, (conv_configs, sa_configs) in enumerate(sa_blocks):
        sa_in_channels.append(in_channels)
        blocks = []
        if conv_configs is not None:
            out_channels, num_blocks, voxel_resolution = conv_configs
            out_channels = int(r * out_channels)
            for k in range(num_blocks):
                # utils.py:139-143: past the first SA stage only the FIRST block of a stage is instantiated (the reference
                # skips the others but still advances the channel count); mirrored for identical keys and arithmetic
                if c == 0 or k == 0:
                    if voxel_resolution is None:
                        blocks.append(SharedMLP(in_channels, out_channels))
                    else:
                        blocks.append(PVConv(in_channels, out_channels, kernel_size=3, resolution=int(vr * voxel_resolution),
                                             dropout=dropout, with_se=with_se, with_se_relu=True, normalize=voxelization_normalize,
                                             eps=eps))
                in_channels = out_channels
            extra_feature_channels = in_channels
        num_centers, radius, num_neighbors, out_channels = sa_configs
        oc = [[int(r * c) for c in o] if isinstance(o, (list, tuple)) else int(r * o) for o in out_channels]
        if num_centers is None:
            blocks.append(PointNetAModule(in_channels=extra_feature_channels, out_channels=oc, include_coordinates=True))
        else:
            blocks.append(PointNetSAModule(num_centers=num_centers, radius=radius, num_neighbors=num_neighbors,
                                           in_channels=extra_feature_channels, out_channels=oc, include_coordinates=True))
        in_channels = extra_feature_channels = blocks[-1].out_channels
        sa_layers.append(_seq_or_single(blocks))
    return sa_layers, sa_in_channels, in_channels, 1 if num_centers is None else num_centers


def create_pointnet2_fp_modules(fp_blocks, in_channels, sa_in_channels, with_se=False, normalize=True, eps=0, dropout=0.1,
                                width_multiplier=1, voxel_resolution_multiplier=1):
    """utils.py:188-247"""
    r, vr = width_multiplier, voxel_resolution_multiplier
    fp_layers = []
    for fp_idx, (fp_configs, conv_configs) in enumerate(fp_blocks):
        blocks = []
        out_channels = tuple(int(r * oc) for oc in fp_configs)
        blocks.append(PointNetFPModule(in_channels=in_channels + sa_in_channels[-1 - fp_idx], out_channels=out_channels))
        in_channels = out_channels[-1]
        if conv_configs is not None:
            out_channels, num_blocks, voxel_resolution = conv_configs
            out_channels = int(r * out_channels)
            for _ in range(num_blocks):
                if voxel_resolution is None:
                    blocks.append(SharedMLP(in_channels, out_channels))
                else:
                    blocks.append(PVConv(in_channels, out_channels, kernel_size=3, resolution=int(vr * voxel_resolution),
                                         dropout=dropout, with_se=with_se, with_se_relu=True, normalize=normalize, eps=eps))
                in_channels = out_channels
        fp_layers.append(_seq_or_single(blocks))
    return fp_layers, in_channels


def _run_blocks(module, inputs):
    """nn.Sequential of tuple-in / tuple-out blocks (the reference chains them the same way)"""
    if isinstance(module, nn.Sequential):
        for m in module:
            inputs = m(inputs)
        return inputs
    return module(inputs)


class PVCNN2(nn.Module):                  # pvcnn_base.py:180-279
    sa_blocks = [((32, 1, 32), (1024, 0.1, 32, (32, 64))), ((64, 2, 16), (256, 0.2, 32, (64, 128))),
                 ((128, 1, 8), (64, 0.4, 32, (128, 256))), (None, (16, 0.8, 32, (256, 256, 512)))]
    fp_blocks = [((256, 256), (256, 1, 8)), ((256, 256), (256, 1, 8)), ((256, 128), (128, 2, 16)), ((128, 128, 64), (64, 1, 32))]

    def __init__(self, in_channels=3, extra_feature_channels=0, width_multiplier=1, voxel_resolution_multiplier=1,
                 use_attention=False, dropout=0.1):
        super().__init__()
        if use_attention:
            raise NotImplementedError("voxel attention is not used")
        self.in_channels = in_channels + extra_feature_channels
        sa_layers, sa_in_channels, channels_sa_features, _ = create_pointnet2_sa_components(
            sa_blocks=self.sa_blocks, extra_feature_channels=extra_feature_channels, with_se=True, voxelization_normalize=True,
            dropout=dropout, width_multiplier=width_multiplier, voxel_resolution_multiplier=voxel_resolution_multiplier)
        self.sa_layers = nn.ModuleList(sa_layers)
        sa_in_channels[0] = extra_feature_channels
        fp_layers, _ = create_pointnet2_fp_modules(
            fp_blocks=self.fp_blocks, in_channels=channels_sa_features, sa_in_channels=sa_in_channels, with_se=True,
            width_multiplier=width_multiplier, voxel_resolution_multiplier=voxel_resolution_multiplier)
        self.fp_layers = nn.ModuleList(fp_layers)
        self.out_channels = self.fp_layers[-1][-1].out_channels

    @torch.no_grad()
    def forward(self, inputs, cond=None):
        if isinstance(inputs, dict):
            inputs = inputs["features"]
        coords, features = inputs[:, :3, :].contiguous(), inputs.contiguous()
        coords_list, in_features_list = [], []
        for sa in self.sa_layers:
            in_features_list.append(features)
            coords_list.append(coords)
            features, coords = _run_blocks(sa, (features, coords))
        in_features_list[0] = inputs[:, 3:, :].contiguous()
        for fp_idx, fp in enumerate(self.fp_layers):
            features, coords = _run_blocks(fp, (coords_list[-1 - fp_idx], coords, features, in_features_list[-1 - fp_idx]))
        return features


class PointNet2SSG(nn.Module):            # pointnet2.py:13-119
    sa_blocks = [(None, (512, 0.2, 64, (64, 64, 128))), (None, (128, 0.4, 64, (128, 128, 256))),
                 (None, (None, None, None, (256, 512, 1024)))]
    fp_blocks = [((256, 256), None), ((256, 128), None), ((128, 128, 128), None)]

    def __init__(self, num_shapes=0, extra_feature_channels=3, width_multiplier=1, voxel_resolution_multiplier=1):
        super().__init__()
        assert extra_feature_channels >= 0
        self.in_channels = extra_feature_channels + 3
        self.num_shapes, self.with_one_hot_shape_id = num_shapes, False
        sa_layers, sa_in_channels, channels_sa_features, _ = create_pointnet2_sa_components(
            sa_blocks=self.sa_blocks, extra_feature_channels=extra_feature_channels, width_multiplier=width_multiplier)
        self.sa_layers = nn.ModuleList(sa_layers)
        fp_layers, _ = create_pointnet2_fp_modules(
            fp_blocks=self.fp_blocks, in_channels=channels_sa_features, sa_in_channels=sa_in_channels,
            width_multiplier=width_multiplier, voxel_resolution_multiplier=voxel_resolution_multiplier)
        self.fp_layers = nn.ModuleList(fp_layers)

    @torch.no_grad()
    def forward(self, inputs):
        features = inputs[:, :self.in_channels, :]
        with_ids = features
        coords, features = features[:, :3, :].contiguous(), features[:, 3:, :].contiguous()
        coords_list, in_features_list = [], []
        for sa in self.sa_layers:
            in_features_list.append(features)
            coords_list.append(coords)
            features, coords = _run_blocks(sa, (features if features.shape[1] > 0 else None, coords))
        in_features_list[0] = with_ids.contiguous()
        for fp_idx, fp in enumerate(self.fp_layers):
            features, coords = _run_blocks(fp, (coords_list[-1 - fp_idx], coords, features, in_features_list[-1 - fp_idx]))
        return features
