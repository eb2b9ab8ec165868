"""Shared by tests/golden/make_golden_ref.py (runs the REFERENCE's files, build container only) and
tests/test_oracle_vs_reference.py (runs the oracle against the committed fixtures, anywhere).

Weights and the large inputs are NOT stored in the fixtures: both sides regenerate them from numpy's legacy
`RandomState` (bit-stable across numpy versions and machines) keyed by the parameter NAME, so the reference module
and the oracle module get identical values whatever their construction order — and the test also proves that the
two `state_dict`s have the same keys and shapes.
"""
import zlib

import numpy as np
import torch


def _rs(name, seed):
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def fill_state(module, seed=0):
    """Deterministic, name-keyed values for every tensor of module.state_dict() (in place)."""
    sd = module.state_dict()
    for name, t in sd.items():
        if not t.is_floating_point():
            continue                                         # num_batches_tracked
        r = _rs(name, seed)
        shape = tuple(t.shape)
        if name.endswith('running_var'):
            v = r.uniform(0.5, 1.5, shape)
        elif name.endswith('running_mean'):
            v = r.normal(0.0, 0.1, shape)
        elif t.dim() <= 1:
            v = r.uniform(0.5, 1.5, shape) if name.endswith('weight') else r.normal(0.0, 0.1, shape)
        else:
            # fan-in: every axis but the output one (last axis for spconv (kD,kH,kW,Cin,Cout), first otherwise)
            fan_in = t.numel() // (shape[-1] if t.dim() == 5 else shape[0])
            v = r.normal(0.0, 1.0, shape) * (1.5 / np.sqrt(fan_in))
        t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)))
    module.load_state_dict(sd)
    return module


def rand(name, shape, seed=0, scale=1.0):
    return torch.from_numpy((_rs(name, seed).normal(0.0, 1.0, shape) * scale).astype(np.float32))


# ----------------------------------------------------------------------------- small fusion-encoder case
BEV = 24                      # 24x24 BEV cells, 6x6 windows -> 4x4 (+ shifted 5x5) windows; level 1: 12x12
PC_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
IMG_SHAPE = (128, 352)        # network input; stride-16 map 8 x 22
NPTS = 12


def fusion_cfg():
    return dict(num_points_in_pillar=NPTS, embed_dims=256, bev_size=BEV, num_views=6,
                region_shape=[(6, 6, 1), (6, 6, 1)], grid_size=[[BEV, BEV, 1], [BEV // 2, BEV // 2, 1]],
                region_drop_info=[{0: {'max_tokens': 36, 'drop_range': (0, 100000)}},
                                  {0: {'max_tokens': 36, 'drop_range': (0, 100000)}}], instance_num=200)


def backbone_cfg():
    return dict(in_channels=128, out_channels=[128, 256], layer_nums=[5, 5], layer_strides=[1, 2],
                norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False))


def fusion_cfg_full():
    """the config's own sizes (isfusion_0075voxel.py:84-97): 180x180 BEV, 30x30 windows, 90x90 second level."""
    c = fusion_cfg()
    c.update(bev_size=180, grid_size=[[180, 180, 1], [90, 90, 1]])
    return c


def sample_positions(shape, n=4096, seed=0):
    """n flat positions of a tensor of `shape` (fixtures of full-size maps hold samples + moments, not the maps)."""
    total = int(np.prod(shape))
    return _rs('samples%s' % (tuple(shape),), seed).randint(0, total, n)


def fusion_inputs(bs=2, seed=0, bev=None, pillars_per_frame=300):
    """Pillars (zero-padded to 12 points like the hard voxelizer), camera features, LiDAR BEV, calibration."""
    global BEV
    if bev is not None and bev != BEV:          # full-size case: same generator on the bigger grid
        saved, BEV = BEV, bev
        try:
            return fusion_inputs(bs, seed + 100, None, pillars_per_frame)
        finally:
            BEV = saved
    r = _rs('fusion_inputs', seed)
    cell = (PC_RANGE[3] - PC_RANGE[0]) / BEV
    pillars, coors = [], []
    for b in range(bs):
        m = pillars_per_frame + 40 * b
        lin = r.permutation(BEV * BEV)[:m]                    # first-come order is arbitrary: keep it unsorted
        y, x = lin // BEV, lin % BEV
        n = r.randint(1, NPTS + 1, size=m)
        p = np.zeros((m, NPTS, 5), np.float32)
        for i in range(m):
            k = n[i]
            p[i, :k, 0] = PC_RANGE[0] + (x[i] + r.uniform(0, 1, k)) * cell
            p[i, :k, 1] = PC_RANGE[1] + (y[i] + r.uniform(0, 1, k)) * cell
            p[i, :k, 2] = r.uniform(-3.0, 1.0, k)
            p[i, :k, 3] = r.uniform(0, 1, k)
        pillars.append(p)
        coors.append(np.stack([np.full(m, b), np.zeros(m, np.int64), y, x], 1))
    pillars = torch.from_numpy(np.concatenate(pillars))
    coors = torch.from_numpy(np.concatenate(coors)).int()
    h, w = IMG_SHAPE[0] // 16, IMG_SHAPE[1] // 16
    img_feats = [rand('img_feats0', (bs * 6, 256, h * 2, w * 2), seed, 0.5), rand('img_feats1', (bs * 6, 256, h, w), seed, 0.5)]
    lidar = torch.relu(rand('lidar_feats', (bs, 512, BEV, BEV), seed, 0.7))
    # six pinhole cameras looking outwards (yaw 0, +-55, +-110, 180 deg), native image 900x1600, f = 1260 px
    l2i, aug = [], []
    s = IMG_SHAPE[1] / 1600.0
    for b in range(bs):
        mats = []
        for yaw in (0.0, 55.0, -55.0, 110.0, -110.0, 180.0):
            a = np.deg2rad(yaw + 3.0 * b)
            fwd = np.array([np.cos(a), np.sin(a), 0.0])
            left = np.array([-np.sin(a), np.cos(a), 0.0])
            up = np.array([0.0, 0.0, 1.0])
            rot = np.stack([-left, -up, fwd])               # camera axes: x right, y down, z forward
            t = -rot @ np.array([0.5 * np.cos(a), 0.5 * np.sin(a), 1.5])
            k = np.array([[1260.0, 0, 800.0], [0, 1260.0, 450.0], [0, 0, 1.0]])
            m4 = np.eye(4)
            m4[:3, :3] = k @ rot
            m4[:3, 3] = k @ t
            mats.append(m4)
        l2i.append(np.stack(mats))
        a4 = np.eye(4)
        a4[0, 0] = a4[1, 1] = s
        a4[1, 3] = IMG_SHAPE[0] - 900.0 * s                   # resize + bottom crop (eval branch of ImageAug3D)
        aug.append(np.stack([a4] * 6))
    th = np.deg2rad(7.0)
    laug = np.eye(4)
    laug[:2, :2] = 1.05 * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    laug[2, 2] = 1.05
    laug[:3, 3] = [0.3, -0.2, 0.1]
    f32 = lambda a: torch.from_numpy(np.asarray(a, np.float32))
    return dict(pillars=pillars, pillar_coors=coors, img_feats=img_feats, lidar_feats=lidar,
                lidar2img=f32(np.stack(l2i)), img_aug_matrix=f32(np.stack(aug)), lidar_aug_matrix=f32(np.stack([laug] * bs)),
                img_metas=[dict(input_shape=IMG_SHAPE) for _ in range(bs)])


# ----------------------------------------------------------------------------- image backbone
def swin_cfg():
    """configs/isfusion/isfusion_0075voxel.py:33-49 (Swin-T)."""
    return dict(embed_dims=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4, qkv_bias=True, qk_scale=None,
                drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, patch_norm=True, out_indices=[1, 2, 3], with_cp=False,
                convert_weights=False)


def swin_input(seed=0):
    """two 62 x 94 images: not a multiple of the 4-pixel patch nor (after embedding) of the 7-token window."""
    return rand('swin_input', (2, 3, 62, 94), seed).half().float()


# ----------------------------------------------------------------------------- image neck
def lss_cfg():
    return dict(in_channels=[192, 384, 768], out_channels=256, start_level=0, num_outs=3)


def lss_inputs(n_img=3, seed=0):
    """Swin-T stage outputs of a 128x352 image (strides 8 / 16 / 32)."""
    return [rand('lss_in%d' % i, (n_img, c, 128 // s, 352 // s), seed, 0.5).half().float()       # fp16-representable
            for i, (c, s) in enumerate(((192, 8), (384, 16), (768, 32)))]


# ----------------------------------------------------------------------------- head
def head_cfg():
    """configs/isfusion/isfusion_0075voxel.py:115-171 on the 24x24 grid (grid_size 192 / out_size_factor 8)."""
    return dict(num_proposals=200, auxiliary=True, in_channels=512, hidden_channel=128, num_classes=10, num_decoder_layers=1,
                num_heads=8, nms_kernel_size=3, ffn_channel=256, dropout=0.1, bn_momentum=0.1, activation='relu',
                common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2)),
                bbox_coder=dict(type='TransFusionBBoxCoder', pc_range=PC_RANGE[:2], voxel_size=[0.5625, 0.5625], out_size_factor=8,
                                post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.0, code_size=10),
                loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2, alpha=0.25, reduction='mean', loss_weight=1.0),
                loss_bbox=dict(type='L1Loss', reduction='mean', loss_weight=0.25),
                loss_heatmap=dict(type='GaussianFocalLoss', reduction='mean', loss_weight=1.0),
                train_cfg=None,
                test_cfg=dict(dataset='nuScenes', grid_size=[BEV * 8, BEV * 8, 40], out_size_factor=8, pc_range=PC_RANGE[:2],
                              voxel_size=[0.5625, 0.5625], nms_type=None, use_rotate_nms=True, nms_thr=0.2, max_num=200))


def head_input(bs=2, seed=0):
    # fp16-representable values: the B200 head stores its input in fp16, so the same numbers reach every implementation
    return torch.relu(rand('head_input', (bs, 512, BEV, BEV), seed, 0.7)).half().float()


# ----------------------------------------------------------------------------- small VFE / sparse-encoder cases
VFE_VOXEL = [0.5, 0.5, 1.0]
VFE_RANGE = [-8.0, -8.0, -2.0, 8.0, 8.0, 2.0]


def vfe_cfg():
    return dict(in_channels=5, feat_channels=[64, 64], with_distance=False, voxel_size=VFE_VOXEL, with_cluster_center=True,
                with_voxel_center=True, point_cloud_range=VFE_RANGE, norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01))


def vfe_inputs(seed=0, n=1500, bs=2):
    r = _rs('vfe_inputs', seed)
    lo, hi = np.array(VFE_RANGE[:3]), np.array(VFE_RANGE[3:])
    pts = np.concatenate([r.uniform(lo - 0.5, hi + 0.5, (n, 3)), r.uniform(0, 1, (n, 2))], 1).astype(np.float32)
    pts[: n // 3, :3] = (pts[: n // 3, :3] * 0.3).astype(np.float32)      # a dense clump: voxels with many points
    c = np.floor((pts[:, :3].astype(np.float32) - lo.astype(np.float32)) / np.array(VFE_VOXEL, np.float32)).astype(np.int64)
    grid = np.array([32, 32, 4])
    ok = ((c >= 0) & (c < grid)).all(1)
    # out-of-range points must be filtered before DynamicVFE (PointsRangeFilter in the config):
    # map_voxel_center_to_point indexes its canvas with the raw coordinates (voxel_encoder.py:444-449)
    pts, c = pts[ok], c[ok]
    zyx = c[:, ::-1]
    n = len(pts)
    b = np.sort(r.randint(0, bs, n))
    b[-1] = bs - 1
    coors = np.concatenate([b[:, None], zyx], 1)
    return torch.from_numpy(pts), torch.from_numpy(coors).int()


SP_SHAPE = [41, 48, 48]


def sparse_cfg():
    return dict(in_channels=64, sparse_shape=SP_SHAPE, base_channels=32, output_channels=256, order=('conv', 'norm', 'act'),
                encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
                encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')


def sparse_inputs(seed=0, v=1400, bs=2):
    """Clustered active voxels in ascending (b,z,y,x) order (the order DynamicVFE hands them over in)."""
    r = _rs('sparse_inputs', seed)
    centres = r.uniform(0.15, 0.85, (12, 3)) * np.array(SP_SHAPE)
    p = centres[r.randint(0, 12, v * 2)] + r.normal(0, 1.0, (v * 2, 3)) * np.array([1.5, 4.0, 4.0])
    c = np.floor(p).astype(np.int64)
    ok = ((c >= 0) & (c < np.array(SP_SHAPE))).all(1)
    c = c[ok]
    b = r.randint(0, bs, len(c))
    full = np.unique(np.concatenate([b[:, None], c], 1), axis=0)
    full = full[r.permutation(len(full))[:v]]
    full = full[np.lexsort((full[:, 3], full[:, 2], full[:, 1], full[:, 0]))]
    feats = torch.relu(rand('sparse_feats', (len(full), 64), seed))
    return feats, torch.from_numpy(full).int()


# ---------------------------------------------------------------------- input contract: multi-sweep loading
def sweep_inputs(seed=12, n_key=2000, n_sweeps=10, n_sweep=1500):
    """(key_points [N0,5] f32, ts seconds, sweeps): synthetic nuScenes-shaped sweeps — ego motion of a few metres, yaw of a few
    degrees, 50 ms apart, a blob of returns on the ego vehicle itself (|x|,|y| < 1 m: what remove_close drops)."""
    rs = np.random.RandomState(seed)

    def cloud(n):
        p = np.empty((n, 5), np.float32)
        p[:, :2] = rs.uniform(-50, 50, (n, 2))
        p[:, 2] = rs.uniform(-4, 2, n)
        p[: n // 20, :2] = rs.uniform(-1.5, 1.5, (n // 20, 2))        # returns on / next to the ego vehicle
        p[:, 3] = rs.uniform(0, 255, n)
        p[:, 4] = rs.uniform(0, 31, n)                                # ring index in the raw file; overwritten by the time lag
        return p

    ts = 1533151603.547590
    sweeps = []
    for i in range(n_sweeps):
        yaw, pitch = np.deg2rad(rs.uniform(-4, 4)), np.deg2rad(rs.uniform(-0.5, 0.5))
        cz, sz, cy, sy = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
        rot = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        sweeps.append(dict(points=cloud(n_sweep), timestamp=(ts - 0.05 * (i + 1)) * 1e6 + rs.uniform(-2e3, 2e3),
                           sensor2lidar_rotation=rot, sensor2lidar_translation=rs.uniform(-3, 3, 3) * np.array([1, 1, 0.05])))
    return cloud(n_key), ts, sweeps


def camera_images(seed=13, n=2, hw=(225, 400)):
    """synthetic uint8 camera frames (quarter-size nuScenes aspect 900x1600): smooth gradients + edges + noise, so that both
    the clipping of the bicubic overshoot (0 / 255) and the rounding between the two resampling passes are exercised."""
    rs = np.random.RandomState(seed)
    h, w = hw
    yy, xx = np.mgrid[0:h, 0:w]
    out = []
    for i in range(n):
        base = 127 + 120 * np.sin(xx / (7.0 + i) + rs.uniform(0, 3)) * np.cos(yy / (5.0 + 2 * i))
        img = np.stack([base, np.roll(base, 13, 1), 255 - base], -1)
        img[h // 3:h // 3 + 20, w // 4:w // 2] = 255                     # saturated block: overshoot clips at 255 / 0 next to it
        img[h // 2:h // 2 + 9, w // 2:w // 2 + 40] = 0
        img = img + rs.normal(0, 12, img.shape)
        out.append(np.clip(img, 0, 255).astype(np.uint8))
    return out


IMG_NORM = dict(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])                     # config:283-285
IMG_AUG_TEST = dict(final_dim=(96, 264), resize_lim=[0.72, 0.72], bot_pct_lim=[0.0, 0.0], rot_lim=[0.0, 0.0], rand_flip=False,
                    is_train=False)                                                          # config:321-327 at quarter size
