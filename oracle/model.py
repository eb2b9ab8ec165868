"""Oracle: the whole fusion-neck forward (ISFusionDetector.extract_pts_feat minus image branch / pts_neck / head).

TEST INFRASTRUCTURE ONLY.  ref: mmdet3d/models/detectors/isfusion.py:84-176 and
configs/isfusion/isfusion_0075voxel.py:24-112.  This is the CPU baseline `bench.py` times.
"""
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fusion_ops, spconv_ops, voxel_ops

from synth import PC_RANGE, VOXEL_SIZE, neck_config  # noqa: E402,F401  (the workload definition is shared with the product's bench)


class ISFusionNeck(nn.Module):
    """voxelize -> DynamicVFE -> SparseEncoder -> pillarize -> ISFusionEncoder (+SECONDV2 stage1/2)."""

    def __init__(self, cfg=None, shuffle_voxels=True):
        super().__init__()
        cfg = cfg or neck_config()
        self.cfg = cfg
        vs, pr = cfg['voxel_size'], cfg['pc_range']
        self.pts_voxel_layer = voxel_ops.Voxelization(vs, pr, -1, (-1, -1))
        self.pts_voxel_encoder = voxel_ops.DynamicVFE(**cfg['pts_voxel_encoder'])
        self.pts_middle_encoder = spconv_ops.SparseEncoder(**cfg['pts_middle_encoder'])
        self.fusion_encoder = fusion_ops.ISFusionEncoder(shuffle_voxels=shuffle_voxels, **cfg['fusion_encoder'])
        self.pts_backbone = fusion_ops.SECONDV2(**cfg['pts_backbone'])
        osf = cfg['out_size_factor']
        self.pillar_size = [vs[0] * osf, vs[1] * osf, pr[5] - pr[2]]
        self.pts_pillar_layer = voxel_ops.Voxelization(self.pillar_size, pr, cfg['fusion_encoder']['num_points_in_pillar'],
                                                       (30000, 60000))

    def dynamic_voxelize(self, points):
        coors = [F.pad(self.pts_voxel_layer(p), (1, 0), value=i) for i, p in enumerate(points)]
        return torch.cat(points, 0), torch.cat(coors, 0)

    def voxelize_pillars(self, points):
        vs, cs, ns = [], [], []
        for i, p in enumerate(points):
            v, c, n = self.pts_pillar_layer(p)
            vs.append(v)
            cs.append(F.pad(c, (1, 0), value=i))
            ns.append(n)
        return torch.cat(vs, 0), torch.cat(ns, 0), torch.cat(cs, 0)

    @torch.no_grad()
    def forward(self, points, img_feats, img_metas, timings=None, **kwargs):
        t = [time.perf_counter()]

        def tick(name):
            if timings is not None:
                t.append(time.perf_counter())
                timings[name] = timings.get(name, 0.0) + t[-1] - t[-2]
        voxels, coors = self.dynamic_voxelize(points)
        tick('voxelize')
        vf, vc = self.pts_voxel_encoder(voxels, coors)
        tick('vfe')
        bs = int(coors[-1, 0]) + 1
        x, _, _ = self.pts_middle_encoder(vf, vc, bs)
        tick('sparse_encoder')
        pillars, npp, pcoors = self.voxelize_pillars(points)
        tick('pillarize')
        pts_metas = dict(pillars=pillars, pillars_num_points=npp, pillar_coors=pcoors, pts=points, pillar_size=self.pillar_size)
        feats, hm = self.fusion_encoder(img_feats, x, bs, pts_metas=pts_metas, img_metas=img_metas,
                                        pts_backbone=self.pts_backbone, **kwargs)
        tick('fusion_encoder')
        return feats, hm
