"""ISFusionDetector point branch (ref: mmdet3d/models/detectors/isfusion.py:13-176): the fusion-neck hot path
`extract_pts_feat` = dynamic voxelize -> DynamicVFE -> SparseEncoder -> pillarize -> ISFusionEncoder (+SECONDV2
stages) wired from the reference's config keys.  Image backbone / pts_neck / bbox head are outside the path
 and are taken as inputs / left to the caller."""
import torch
import torch.nn.functional as F
from torch import nn

from . import fusion_encoder, sparse_encoder, voxel_encoder  # noqa: F401  (register the modules)
from .registry import DETECTORS, build_backbone, build_fusion_layer, build_middle_encoder, build_voxel_encoder
from .profiling import stage
from .voxel_layer import Voxelization


@DETECTORS.register_module()
class ISFusionDetector(nn.Module):

    def __init__(self, pts_voxel_layer=None, pts_voxel_encoder=None, pts_middle_encoder=None, fusion_encoder=None,
                 pts_backbone=None, pts_neck=None, pts_bbox_head=None, img_backbone=None, img_neck=None,
                 train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None, **kwargs):
        super().__init__()
        self.pts_voxel_layer = Voxelization(**pts_voxel_layer)
        self.pts_voxel_encoder = build_voxel_encoder(dict(pts_voxel_encoder))
        self.pts_middle_encoder = build_middle_encoder(dict(pts_middle_encoder))
        self.fusion_encoder = build_fusion_layer(dict(fusion_encoder))
        self.pts_backbone = build_backbone(dict(pts_backbone))
        self.detach = kwargs.get('detach', False)
        out_size_factor = kwargs.get('out_size_factor', None)
        voxel_size = kwargs.get('voxel_size', None)
        self.pc_range = kwargs.get('pc_range', None)
        self.pillar_size = [voxel_size[0] * out_size_factor, voxel_size[1] * out_size_factor,
                            self.pc_range[5] - self.pc_range[2]]
        self.pts_pillar_layer = Voxelization(max_num_points=self.fusion_encoder.num_points_in_pillar,
                                             voxel_size=self.pillar_size, max_voxels=(30000, 60000),
                                             point_cloud_range=self.pc_range)

    @torch.no_grad()
    def dynamic_voxelize(self, points):
        """ref: isfusion.py:123-146."""
        coors = [F.pad(self.pts_voxel_layer(res), (1, 0), mode='constant', value=i) for i, res in enumerate(points)]
        return torch.cat(points, dim=0), torch.cat(coors, dim=0)

    @torch.no_grad()
    def voxelize(self, points, voxel_type='pillar'):
        """ref: isfusion.py:148-176."""
        voxels, coors, num_points = [], [], []
        for i, res in enumerate(points):
            v, c, n = self.pts_pillar_layer(res)
            voxels.append(v)
            coors.append(F.pad(c, (1, 0), mode='constant', value=i))
            num_points.append(n)
        return torch.cat(voxels, dim=0), torch.cat(num_points, dim=0), torch.cat(coors, dim=0)

    def isfusion(self, pts, pts_feats, img_feats, img_metas, batch_size, **kwargs):
        """ref: isfusion.py:84-101."""
        with stage('a5 pillarize'):
            pillars, pillars_num_points, pillar_coors = self.voxelize(pts, voxel_type='pillar')
        pts_metas = dict(pillars=pillars, pillars_num_points=pillars_num_points, pillar_coors=pillar_coors, pts=pts,
                         pillar_size=self.pillar_size)
        kwargs.update(dict(pts_metas=pts_metas, img_metas=img_metas, pts_backbone=self.pts_backbone))
        return self.fusion_encoder(img_feats, pts_feats, batch_size, **kwargs)

    @torch.no_grad()
    def extract_pts_feat(self, pts, img_feats, img_metas, **kwargs):
        """ref: isfusion.py:103-121 without pts_neck; returns ([feat_0, feat_1], ins_heatmap)."""
        with stage('a1 dynamic_voxelize'):
            voxels, coors = self.dynamic_voxelize(pts)
        batch_size = len(pts)
        with stage('a3 dynamic_vfe'):
            voxel_features, feature_coors = self.pts_voxel_encoder(voxels, coors, pts, img_feats, img_metas,
                                                                   batch_size=batch_size)
        with stage('a4 sparse_encoder'):
            x, _, kwargs = self.pts_middle_encoder(voxel_features, feature_coors, batch_size, dense=False, **kwargs)
        kwargs['lidar_sparse'] = self.pts_middle_encoder.last_out
        return self.isfusion(pts, x, img_feats, img_metas, batch_size, **kwargs)

    forward = extract_pts_feat
