"""Seeded synthetic nuScenes-shaped inputs shared by tests/ and bench.py.

points: 'U' = uniform box, 'S' = crude 32-beam LiDAR (ground plane + objects).  calibration: 6 pinhole cameras.
Everything comes from torch.Generator().manual_seed(seed) on the CPU so that the oracle and the CUDA path see
identical bits.
"""
import math

import torch

PC_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
VOXEL_SIZE = [0.075, 0.075, 0.2]


def neck_config(voxel_size=VOXEL_SIZE, pc_range=PC_RANGE, out_size_factor=8):
    """the module kwargs of configs/isfusion/isfusion_0075voxel.py:24-112 (the workload both bench arms and the tests run)."""
    voxel_shape = int((pc_range[3] - pc_range[0]) // voxel_size[0])
    bev = voxel_shape // out_size_factor
    return dict(
        voxel_size=voxel_size, pc_range=pc_range, out_size_factor=out_size_factor, bev_size=bev,
        pts_voxel_encoder=dict(in_channels=5, feat_channels=[64, 64], with_distance=False, voxel_size=voxel_size,
                               with_cluster_center=True, with_voxel_center=True, point_cloud_range=pc_range,
                               norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01)),
        pts_middle_encoder=dict(in_channels=64, sparse_shape=[41, voxel_shape, voxel_shape], base_channels=32,
                                output_channels=256, order=('conv', 'norm', 'act'),
                                encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
                                encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock'),
        fusion_encoder=dict(num_points_in_pillar=12, embed_dims=256, bev_size=bev, num_views=6,
                            region_shape=[(6, 6, 1), (6, 6, 1)], grid_size=[[bev, bev, 1], [bev // 2, bev // 2, 1]],
                            region_drop_info=[{0: {'max_tokens': 36, 'drop_range': (0, 100000)}},
                                              {0: {'max_tokens': 36, 'drop_range': (0, 100000)}}], instance_num=200),
        pts_backbone=dict(in_channels=128, out_channels=[128, 256], layer_nums=[5, 5], layer_strides=[1, 2],
                          norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False)))


def full_model_config(voxel_size=VOXEL_SIZE, pc_range=PC_RANGE, out_size_factor=8):
    """the whole `model = dict(...)` of configs/isfusion/isfusion_0075voxel.py:24-178 (inference keys): image branch
    (SwinTransformer + GeneralizedLSSFPN), the neck (neck_config), SECONDFPN, TransFusionHeadV2 and test_cfg — for
    `isfusion_b200.registry.build_detector`.  tests/test_config_dropin.py checks it key by key against the reference's file."""
    c = neck_config(voxel_size, pc_range, out_size_factor)
    voxel_shape = int((pc_range[3] - pc_range[0]) // voxel_size[0])
    return dict(
        type='ISFusionDetector', detach=True, pc_range=pc_range, voxel_size=voxel_size, out_size_factor=out_size_factor,
        img_backbone=dict(type='SwinTransformer', embed_dims=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7,
                          mlp_ratio=4, qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2,
                          patch_norm=True, out_indices=[1, 2, 3], with_cp=False, convert_weights=False),
        img_neck=dict(type='GeneralizedLSSFPN', in_channels=[192, 384, 768], out_channels=256, start_level=0, num_outs=3),
        pts_voxel_layer=dict(point_cloud_range=pc_range, max_num_points=-1, voxel_size=voxel_size, max_voxels=(-1, -1)),
        pts_voxel_encoder=dict(type='DynamicVFE', **c['pts_voxel_encoder']),
        pts_middle_encoder=dict(type='SparseEncoder', **c['pts_middle_encoder']),
        fusion_encoder=dict(type='ISFusionEncoder', **c['fusion_encoder']),
        pts_backbone=dict(type='SECONDV2', **c['pts_backbone']),
        pts_neck=dict(type='SECONDFPN', in_channels=[128, 256], out_channels=[256, 256], upsample_strides=[1, 2],
                      norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), upsample_cfg=dict(type='deconv', bias=False),
                      use_conv_for_no_stride=True),
        pts_bbox_head=dict(
            type='TransFusionHeadV2', num_proposals=200, auxiliary=True, in_channels=256 * 2, hidden_channel=128, num_classes=10,
            num_decoder_layers=1, num_heads=8, nms_kernel_size=3, ffn_channel=256, dropout=0.1, bn_momentum=0.1, activation='relu',
            common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2)),
            bbox_coder=dict(type='TransFusionBBoxCoder', pc_range=pc_range[:2], voxel_size=voxel_size[:2],
                            out_size_factor=out_size_factor, post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
                            score_threshold=0.0, code_size=10)),
        test_cfg=dict(pts=dict(dataset='nuScenes', grid_size=[voxel_shape, voxel_shape, 40], out_size_factor=out_size_factor,
                               pc_range=pc_range[0:2], voxel_size=voxel_size[:2], nms_type=None, use_rotate_nms=True, nms_thr=0.2,
                               max_num=200)))


def make_points(n, kind='S', seed=0, pc_range=PC_RANGE):
    g = torch.Generator().manual_seed(seed)
    lo = torch.tensor(pc_range[:3])
    hi = torch.tensor(pc_range[3:])
    eps = 1e-3
    if kind == 'U':
        xyz = torch.rand(n, 3, generator=g) * (hi - lo - 2 * eps) + lo + eps
    else:
        beams = torch.linspace(math.radians(-30.0), math.radians(10.0), 32)
        m = int(n * 1.25)
        b = torch.randint(0, 32, (m,), generator=g)
        elev = beams[b] + torch.randn(m, generator=g) * 0.002
        az = torch.rand(m, generator=g) * 2 * math.pi
        # ground return: z = -1.84 => range = -1.84 / sin(elev) for downward beams
        ground_z = -1.84 + torch.randn(m, generator=g) * 0.02
        rng = torch.where(elev < -0.01, ground_z / torch.sin(elev.clamp(max=-0.01)), torch.full((m,), 1e9))
        obj = torch.rand(m, generator=g) < 0.35
        obj_r = torch.rand(m, generator=g) * 57 + 3
        rng = torch.where(obj, torch.minimum(rng, obj_r), rng)
        far = rng > 75
        rng = torch.where(far, torch.rand(m, generator=g) * 65 + 5, rng)
        x = rng * torch.cos(elev) * torch.cos(az)
        y = rng * torch.cos(elev) * torch.sin(az)
        z = rng * torch.sin(elev)
        xyz = torch.stack([x, y, z], 1)
        ok = ((xyz > lo + eps) & (xyz < hi - eps)).all(1)
        xyz = xyz[ok][:n]
    inten = torch.rand(xyz.shape[0], 1, generator=g)
    dt = torch.randint(0, 10, (xyz.shape[0], 1), generator=g).float() * 0.05
    return torch.cat([xyz, inten, dt], 1).contiguous()


def make_calib(batch, img_hw=(256, 704), seed=0):
    """lidar2img [B,6,4,4], img_aug_matrix [B,6,4,4], lidar_aug_matrix [B,4,4] (eval: identity)."""
    g = torch.Generator().manual_seed(seed + 1000)
    yaws = [0.0, 55.0, -55.0, 110.0, -110.0, 180.0]
    f, cx, cy = 1260.0, 800.0, 450.0
    K = torch.tensor([[f, 0, cx], [0, f, cy], [0, 0, 1.0]])
    l2i = torch.zeros(batch, 6, 4, 4)
    for b in range(batch):
        for c, yaw in enumerate(yaws):
            a = math.radians(yaw)
            # camera looks along (cos a, sin a, 0); cam axes: x right, y down, z forward
            fwd = torch.tensor([math.cos(a), math.sin(a), 0.0])
            right = torch.tensor([math.sin(a), -math.cos(a), 0.0])
            down = torch.tensor([0.0, 0.0, -1.0])
            R = torch.stack([right, down, fwd], 0)
            t_cam = torch.tensor([1.0 * math.cos(a), 0.5 * math.sin(a), 1.5]) + (torch.rand(3, generator=g) - 0.5) * 0.2
            T = torch.eye(4)
            T[:3, :3] = K @ R
            T[:3, 3] = -(K @ R) @ t_cam
            l2i[b, c] = T
    s = img_hw[1] / 1600.0 * (0.48 / 0.44) if img_hw == (256, 704) else 0.72
    s = 0.48 if img_hw == (256, 704) else 0.72
    aug = torch.eye(4).repeat(batch, 6, 1, 1)
    new_w, new_h = 1600 * s, 900 * s
    crop_x = (new_w - img_hw[1]) / 2
    crop_y = new_h - img_hw[0]
    aug[:, :, 0, 0] = s
    aug[:, :, 1, 1] = s
    aug[:, :, 0, 3] = -crop_x
    aug[:, :, 1, 3] = -crop_y
    lidar_aug = torch.eye(4).repeat(batch, 1, 1)
    return l2i, aug, lidar_aug


def make_img_feats(batch, img_hw=(256, 704), channels=256, seed=0):
    """the stride-16 FPN level the fusion encoder consumes (img_feats[1]) [B*6, 256, h, w]; the other levels are
    present but unused by the path (fusion_encoder.py:1160)."""
    g = torch.Generator().manual_seed(seed + 2000)
    h, w = img_hw[0] // 16, img_hw[1] // 16
    lvl1 = torch.randn(batch * 6, channels, h, w, generator=g) * 0.5
    lvl0 = torch.zeros(batch * 6, channels, img_hw[0] // 8, img_hw[1] // 8)
    lvl2 = torch.zeros(batch * 6, channels, max(h // 2, 1), max(w // 2, 1))
    return [lvl0, lvl1, lvl2]


def randomize_bn(module, seed=0):
    """BN running stats randomised so that BN is not a no-op (SURVEY §8d 'weights')."""
    g = torch.Generator().manual_seed(seed + 3000)
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return module
