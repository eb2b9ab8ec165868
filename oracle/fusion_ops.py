"""Oracle: ISFusionEncoder (HSF + IGF) and SECONDV2 (rows a6, a7, a10-a14 of SURVEY §8).

TEST INFRASTRUCTURE ONLY.  PINNED by outputs of the reference's own fusion_encoder.py / second.py / second_fpn.py run
in the build container (tests/golden/ref_fusion.npz, tests/test_oracle_vs_reference.py); the deformable-attention
core is the reference's own pure-PyTorch ``ms_deform_attn_core_pytorch`` (ops/functions/ms_deform_attn_func.py:41-61,
pinned by ops/test.py:32-60 and tests/golden/msda_ref.npz).  Literal restatement of
mmdet3d/models/middle_encoders/fusion_encoder.py (line numbers cited per function) and
mmdet3d/models/backbones/second.py:98-239, eval mode, including the quirks of SURVEY Appendix C.
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import sst_ops


class ConvModule(nn.Sequential):
    """mmcv ConvModule: conv(bias=False) -> BN2d (torch defaults eps 1e-5) -> ReLU, names conv/bn/activate
    (fusion_encoder.py:862-958 all use norm_cfg=dict(type='BN2d'))."""

    def __init__(self, cin, cout, k=3, padding=1):
        super().__init__()
        self.add_module('conv', nn.Conv2d(cin, cout, k, padding=padding, bias=False))
        self.add_module('bn', nn.BatchNorm2d(cout))
        self.add_module('activate', nn.ReLU(inplace=True))


def ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    """ref: ops/functions/ms_deform_attn_func.py:41-61 (restated verbatim in behaviour)."""
    N_, S_, M_, D_ = value.shape
    _, Lq_, M_, L_, P_, _ = sampling_locations.shape
    value_list = value.split([int(H_ * W_) for H_, W_ in value_spatial_shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampling_value_list = []
    for lid_, (H_, W_) in enumerate(value_spatial_shapes):
        H_, W_ = int(H_), int(W_)
        value_l_ = value_list[lid_].flatten(2).transpose(1, 2).reshape(N_ * M_, D_, H_, W_)
        sampling_grid_l_ = sampling_grids[:, :, :, lid_].transpose(1, 2).flatten(0, 1)
        sampling_value_list.append(F.grid_sample(value_l_, sampling_grid_l_, mode='bilinear', padding_mode='zeros',
                                                 align_corners=False))
    attention_weights = attention_weights.transpose(1, 2).reshape(N_ * M_, 1, Lq_, L_ * P_)
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) * attention_weights).sum(-1).view(N_, M_ * D_, Lq_)
    return output.transpose(1, 2).contiguous()


class MultiheadAttention(nn.Module):
    """ref: fusion_encoder.py:44-171 + multi_head_attention_forward :221-470 (kv_same branch: q from
    in_proj[:E], k,v from in_proj[E:]; q scaled by head_dim**-0.5; softmax; out_proj).  Parameter names
    in_proj_weight / in_proj_bias / out_proj.* as in the reference."""

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, query, key, value):
        L, B, E = query.shape
        S = key.shape[0]
        H, D = self.num_heads, E // self.num_heads
        q = F.linear(query, self.in_proj_weight[:E], self.in_proj_bias[:E]) * (float(D) ** -0.5)
        k = F.linear(key, self.in_proj_weight[E:2 * E], self.in_proj_bias[E:2 * E])
        v = F.linear(value, self.in_proj_weight[2 * E:], self.in_proj_bias[2 * E:])
        q = q.contiguous().view(L, B * H, D).transpose(0, 1)
        k = k.contiguous().view(S, B * H, D).transpose(0, 1)
        v = v.contiguous().view(S, B * H, D).transpose(0, 1)
        w = F.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
        o = torch.bmm(w, v).transpose(0, 1).contiguous().view(L, B, E)
        return F.linear(o, self.out_proj.weight, self.out_proj.bias), None


class PositionEmbeddingLearned(nn.Module):
    """ref: fusion_encoder.py:173-189."""

    def __init__(self, input_channel, num_pos_feats=288):
        super().__init__()
        self.position_embedding_head = nn.Sequential(nn.Conv1d(input_channel, num_pos_feats, kernel_size=1),
                                                     nn.BatchNorm1d(num_pos_feats), nn.ReLU(inplace=True),
                                                     nn.Conv1d(num_pos_feats, num_pos_feats, kernel_size=1))

    def forward(self, xyz):
        return self.position_embedding_head(xyz.transpose(1, 2).contiguous())


class Instane2SceneAtt(nn.Module):
    """ref: fusion_encoder.py:472-502 (incl. the un-scaled per-channel softmax(S.I^T).I, Appendix C #3)."""

    def __init__(self, d_model, nhead=8, dropout=0.1):
        super().__init__()
        self.multihead_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, query, key, query_scene, bs, bev_size):
        query = query.permute(2, 0, 1)
        key = key.permute(2, 0, 1)
        query2 = self.multihead_attn(query=query, key=key, value=key)[0]
        query = self.norm(query + query2).permute(1, 2, 0)
        query_ins = query.reshape(bs, query.shape[1], bev_size, bev_size)
        attention_weights = F.softmax(torch.matmul(query_scene, query_ins.transpose(2, 3)), dim=-1)
        return query_scene + torch.matmul(attention_weights, query_ins)


class MSDeformAttn(nn.Module):
    """ref: fusion_encoder.py:505-600."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        """ref: :544-558."""
        nn.init.constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(
            1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.)
        nn.init.constant_(self.attention_weights.bias.data, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index):
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        value = self.value_proj(input_flatten).view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        so = self.sampling_offsets(query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        aw = self.attention_weights(query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        aw = F.softmax(aw, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + so / offset_normalizer[None, None, None, :, None, :]
        out = ms_deform_attn_core_pytorch(value, input_spatial_shapes, loc, aw)   # fp32 (:90-128 of the fn wrapper)
        return self.output_proj(out), loc, aw


class DeformableTransformerDecoderLayer(nn.Module):
    """ref: fusion_encoder.py:602-674 (self-attn -> deformable cross-attn -> FFN, post-norm, ReLU)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm3 = nn.LayerNorm(d_model)

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index):
        q = k = tgt + query_pos
        tgt2 = self.self_attn(q.transpose(0, 1), k.transpose(0, 1), tgt.transpose(0, 1))[0].transpose(0, 1)
        tgt = self.norm2(tgt + tgt2)
        tgt2 = self.cross_attn(tgt + query_pos, reference_points, src, src_spatial_shapes, level_start_index)[0]
        tgt = self.norm1(tgt + tgt2)
        tgt2 = self.linear2(F.relu(self.linear1(tgt)))
        return self.norm3(tgt + tgt2)


class InsContextAtt(nn.Module):
    """ref: fusion_encoder.py:768-830."""

    def __init__(self, num_layers=1, embed_dims=128, bev_size=180, n_points=16):
        super().__init__()
        self.bev_size = bev_size
        layer = DeformableTransformerDecoderLayer(d_model=embed_dims, d_ffn=embed_dims, dropout=0.1, n_levels=1,
                                                  n_points=n_points)
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        self.query_pos_embed = PositionEmbeddingLearned(2, embed_dims)
        self.key_pos_embed = PositionEmbeddingLearned(2, embed_dims)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()

    def forward(self, query_feats, query_pos, bev_pos, scene_feats=None):
        scene_feats = scene_feats.permute(0, 1, 3, 2)
        key_pos = self.key_pos_embed(bev_pos / self.bev_size).permute(0, 2, 1)
        output = query_feats.transpose(1, 2)
        reference_points = query_pos / self.bev_size
        query_pos_embed = self.query_pos_embed(reference_points).permute(0, 2, 1)
        bs, c, h, w = scene_feats.shape
        src = scene_feats.flatten(2).transpose(1, 2) + key_pos
        spatial_shapes = torch.as_tensor([(h, w)], dtype=torch.long)
        level_start_index = torch.zeros((1,), dtype=torch.long)
        for layer in self.layers:
            output = layer(output, query_pos_embed, reference_points[:, :, None], src, spatial_shapes, level_start_index)
        return output.transpose(1, 2)


class SECONDV2(nn.Module):
    """ref: mmdet3d/models/backbones/second.py:98-239 (staged execution from inside the fusion encoder)."""

    def __init__(self, in_channels=128, out_channels=[128, 128, 256], layer_nums=[3, 5, 5], layer_strides=[2, 2, 2],
                 norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False), **kw):
        super().__init__()
        eps, mom = norm_cfg.get('eps', 1e-5), norm_cfg.get('momentum', 0.1)
        in_filters = [in_channels, *out_channels[:-1]]
        blocks = []
        for i, layer_num in enumerate(layer_nums):
            if layer_strides[i] == 2:
                self.ds_layer = nn.Sequential(nn.Conv2d(in_filters[i], out_channels[i], 3, stride=2, padding=1, bias=False),
                                              nn.BatchNorm2d(out_channels[i], eps=eps, momentum=mom), nn.ReLU(inplace=True))
                block = []
            else:
                block = [nn.Conv2d(in_filters[i], out_channels[i], 3, stride=layer_strides[i], padding=1, bias=False),
                         nn.BatchNorm2d(out_channels[i], eps=eps, momentum=mom), nn.ReLU(inplace=True)]
            for _ in range(layer_num):
                block += [nn.Conv2d(out_channels[i], out_channels[i], 3, padding=1, bias=False),
                          nn.BatchNorm2d(out_channels[i], eps=eps, momentum=mom), nn.ReLU(inplace=True)]
            blocks.append(nn.Sequential(*block))
        self.blocks = nn.ModuleList(blocks)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    @staticmethod
    def create_dense_coord(x_size, y_size, batch_size):
        bx, by = torch.meshgrid(torch.linspace(0, x_size - 1, x_size), torch.linspace(0, y_size - 1, y_size), indexing='ij')
        base = torch.cat([torch.zeros_like(bx)[None], bx[None], by[None]], dim=0)
        return torch.stack([torch.cat([torch.ones_like(bx)[None] * i, base], dim=0) for i in range(batch_size)])

    @torch.no_grad()
    def forward(self, x, stage=None):
        if stage == 'stage1':
            x = self.blocks[0](x[0])
            rf = self.ds_layer(x)
            N, C, H, W = rf.shape
            rf = rf.reshape(N, C, -1).permute(0, 2, 1).reshape(-1, C)
            coords = self.create_dense_coord(H, W, N).type_as(rf).int()
            rc = torch.cat([coords[k].reshape(4, -1).transpose(1, 0) for k in range(N)], dim=0)
            return rf, rc, x
        elif stage == 'stage2':
            return None, None, self.blocks[1](x[0])
        x1 = self.blocks[0](x)
        return (x1, self.blocks[1](self.ds_layer(x1)))


class SECONDFPN(nn.Module):
    """ref: mmdet3d/models/necks/second_fpn.py:11-98 (config isfusion_0075voxel.py:105-112):
    level 0: 1x1 conv (use_conv_for_no_stride) + BN + ReLU, level 1: 2x2 stride-2 deconv + BN + ReLU, concat,
    then the [Y,X] -> [X,Y] transpose of :96."""

    def __init__(self, in_channels=[128, 128, 256], out_channels=[256, 256, 256], upsample_strides=[1, 2, 4],
                 norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), upsample_cfg=dict(type='deconv', bias=False),
                 conv_cfg=dict(type='Conv2d', bias=False), use_conv_for_no_stride=False, **kw):
        super().__init__()
        eps, mom = norm_cfg.get('eps', 1e-5), norm_cfg.get('momentum', 0.1)
        blocks = []
        for i, oc in enumerate(out_channels):
            st = upsample_strides[i]
            if st > 1 or (st == 1 and not use_conv_for_no_stride):
                up = nn.ConvTranspose2d(in_channels[i], oc, st, stride=st, bias=False)
            else:
                up = nn.Conv2d(in_channels[i], oc, 1, stride=1, bias=False)
            blocks.append(nn.Sequential(up, nn.BatchNorm2d(oc, eps=eps, momentum=mom), nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(blocks)

    @torch.no_grad()
    def forward(self, x):
        ups = [d(x[i]) for i, d in enumerate(self.deblocks)]
        out = torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]
        return [out.permute(0, 1, 3, 2).contiguous()]


class GeneralizedLSSFPN(nn.Module):
    """ref: mmdet3d/models/necks/generalized_lss.py:12-103 (the producer of img_feats; config
    isfusion_0075voxel.py:50-55): top-down, per level: bilinear upsample (align_corners=True) of the level above ->
    concat -> 1x1 ConvModule -> 3x3 ConvModule (conv bias-free + BN2d + ReLU each)."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, **kw):
        super().__init__()
        assert start_level == 0 and end_level == -1
        self.in_channels = in_channels
        last = len(in_channels) - 1
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(last):
            cin = in_channels[i] + (in_channels[i + 1] if i == last - 1 else out_channels)
            self.lateral_convs.append(ConvModule(cin, out_channels, 1, padding=0))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1))

    @torch.no_grad()
    def forward(self, inputs):
        lat = list(inputs)
        for i in range(len(lat) - 2, -1, -1):
            x = F.interpolate(lat[i + 1], size=lat[i].shape[2:], mode='bilinear', align_corners=True)
            lat[i] = self.fpn_convs[i](self.lateral_convs[i](torch.cat([lat[i], x], dim=1)))
        return tuple(lat[:-1])


class ISFusionEncoder(nn.Module):
    """ref: fusion_encoder.py:833-1189."""

    def __init__(self, num_points_in_pillar=10, embed_dims=256, num_classes=10, **kwargs):
        super().__init__()
        self.num_points_in_pillar = num_points_in_pillar
        self.bev_size = kwargs.get('bev_size', 180)
        self.num_views = kwargs.get('num_views', 6)
        region_shape = kwargs.get('region_shape', None)
        grid_size = kwargs.get('grid_size', None)
        region_drop_info = kwargs.get('region_drop_info', None)
        self.embed_dims = embed_dims
        self.conv_fusion = ConvModule(embed_dims * 3, embed_dims // 2)
        self.get_regions = nn.ModuleList()
        self.grid2region_att = nn.ModuleList()
        for l in range(len(region_shape)):
            d = embed_dims // 2 * (l + 1)
            self.get_regions.append(sst_ops.SSTInputLayerV2(window_shape=region_shape[l], sparse_shape=grid_size[l],
                                                            shuffle_voxels=kwargs.get('shuffle_voxels', True),
                                                            drop_info=region_drop_info[l], pos_temperature=1000,
                                                            normalize_pos=False, pos_embed=d))
            self.grid2region_att.append(sst_ops.SSTv2(d_model=[d] * 4, nhead=[8] * 4, num_blocks=1, dim_feedforward=[d] * 4,
                                                      output_shape=grid_size[l][:2],
                                                      in_channel=embed_dims // 2 if l == 0 else None))
        self.instance_num = kwargs.get('instance_num', 200)
        self.nms_kernel_size = 3
        bx, by = torch.meshgrid(torch.linspace(0, self.bev_size - 1, self.bev_size),
                                torch.linspace(0, self.bev_size - 1, self.bev_size), indexing='ij')
        self.bev_pos = torch.cat([(bx + 0.5)[None], (by + 0.5)[None]], dim=0)[None].view(1, 2, -1).permute(0, 2, 1)  # :901-913
        h = embed_dims // 2
        self.conv_ins = ConvModule(h, h)
        self.conv_scene = ConvModule(h, h)
        self.conv_heatmap = ConvModule(h, h)
        self.heatmap_head_1 = ConvModule(h, embed_dims // 4)
        self.heatmap_head_2 = ConvModule(embed_dims // 4, embed_dims // 4)
        self.heatmap_head_3 = nn.Conv2d(embed_dims // 4, num_classes, kernel_size=3, stride=1, padding=1)
        self.instance_att = InsContextAtt(num_layers=2, embed_dims=h, bev_size=self.bev_size)
        self.instance_to_scene_att = Instane2SceneAtt(d_model=h)

    def img_point_sampling(self, reference_voxel, mlvl_feats, num_cam, batch_size, **kwargs):
        """ref: :965-1044 (eval branch: no projection noise)."""
        img_aug_matrix, lidar_aug_matrix, lidar2image = kwargs['img_aug_matrix'], kwargs['lidar_aug_matrix'], kwargs['lidar2img']
        image_size = kwargs['img_metas'][0]['input_shape']
        cams = []
        for b in range(batch_size):
            cur = reference_voxel[b].reshape(-1, 3)[:, :3].clone()
            cur -= lidar_aug_matrix[b][:3, 3]
            cur = torch.inverse(lidar_aug_matrix[b][:3, :3]).matmul(cur.transpose(1, 0))
            cur = lidar2image[b][:, :3, :3].matmul(cur)
            cur += lidar2image[b][:, :3, 3].reshape(-1, 3, 1)
            cur[:, 2, :] = torch.clamp(cur[:, 2, :], 1e-5, 1e5)
            cur[:, :2, :] /= cur[:, 2:3, :]
            cur = img_aug_matrix[b][:, :3, :3].matmul(cur)
            cur += img_aug_matrix[b][:, :3, 3].reshape(-1, 3, 1)
            cur = cur[:, :2, :].transpose(1, 2)
            cur[..., 0] /= image_size[1]
            cur[..., 1] /= image_size[0]
            cams.append((cur - 0.5) * 2)
        feat = mlvl_feats[0]
        _, C, H, W = feat.size()
        feat = feat.view(batch_size, num_cam, C, H, W).transpose(0, 1)
        out = []
        for b in range(batch_size):
            acc = 0
            for k in range(num_cam):
                pts = cams[b][k].view(-1, 1, 2)
                s = F.grid_sample(feat[k][b].unsqueeze(0), pts.unsqueeze(0), align_corners=False)
                acc = acc + s.view(1, C, -1, 1).squeeze(0).squeeze(-1)
            out.append(acc)                                    # mask is computed but never applied (Appendix C #1)
        return out

    def img_fv_to_bev(self, mlvl_feats, bs, **kwargs):
        """ref: :1046-1070."""
        pm = kwargs['pts_metas']
        pillars = pm['pillars'][..., :3]
        pc = pm['pillar_coors']
        npp = pm['pillars'].shape[1]
        ref_3d, pcl = [], []
        for i in range(bs):
            m = pc[:, 0] == i
            pcl.append(pc[m])
            ref_3d.append(pillars[m])
        sampled = self.img_point_sampling(ref_3d, mlvl_feats, self.num_views, bs, **kwargs)
        dec = torch.zeros([bs, self.embed_dims, self.bev_size, self.bev_size]).type_as(mlvl_feats[0])
        for b in range(bs):
            out = sampled[b].reshape(self.embed_dims, -1, npp)
            dec[b, :, pcl[b][:, 2].long(), pcl[b][:, 3].long()] = out.sum(dim=2)
        return dec

    def instance_fusion(self, bev_feats, scene_feats, bs):
        """ref: :1090-1149."""
        bev_pos = self.bev_pos.repeat(bs, 1, 1)
        out = bev_feats.permute(0, 1, 3, 2).contiguous()
        hm = self.heatmap_head_3(self.heatmap_head_2(self.heatmap_head_1(self.conv_heatmap(out.clone()))))
        heatmap = hm.sigmoid()
        pad = self.nms_kernel_size // 2
        local_max = torch.zeros_like(heatmap)
        local_max[:, :, pad:(-pad), pad:(-pad)] = F.max_pool2d(heatmap, kernel_size=self.nms_kernel_size, stride=1, padding=0)
        if self.num_views == 6:
            local_max[:, 8] = heatmap[:, 8]
            local_max[:, 9] = heatmap[:, 9]
        heatmap = heatmap * (heatmap == local_max)
        heatmap = heatmap.view(bs, heatmap.shape[1], -1)
        top = heatmap.view(bs, -1).argsort(dim=-1, descending=True)[..., :self.instance_num]
        top_idx = natural_idx = top % heatmap.shape[-1]
        # test hook: downstream arithmetic on a GIVEN selection (the discrete top-K flips on near-ties between fp16 and fp32
        # heat-maps; a parity test that wants to isolate the arithmetic passes the selection of the path under test)
        if getattr(self, 'force_top_idx', None) is not None:
            top_idx = self.force_top_idx.to(top_idx.device).long()
        query_pos = bev_pos.gather(index=top_idx[:, None, :].permute(0, 2, 1).expand(-1, -1, 2), dim=1)
        qpn = torch.zeros_like(query_pos)
        qpn[..., 0] = query_pos[..., 1]
        qpn[..., 1] = query_pos[..., 0]
        x_scene = self.conv_scene(bev_feats.permute(0, 1, 3, 2))
        x_ins = x_scene.reshape(bs, x_scene.shape[1], -1).gather(index=top_idx[:, None, :].expand(-1, x_scene.shape[1], -1), dim=-1)
        x_ins = self.instance_att(x_ins, qpn, bev_pos, scene_feats=x_scene)
        bf = self.conv_ins(bev_feats).flatten(2, 3)
        ret = self.instance_to_scene_att(bf, x_ins, scene_feats, bs, self.bev_size)
        self._dbg = dict(heatmap=hm, top_idx=natural_idx, top_idx_used=top_idx, x_ins=x_ins, x_scene=x_scene)
        return ret, hm

    @torch.no_grad()
    def forward(self, img_mlvl_feats, lidar_feats, bs, **kwargs):
        """ref: :1152-1189."""
        img_bev = self.img_fv_to_bev([img_mlvl_feats[1]], bs, **kwargs)
        bev = self.conv_fusion(torch.cat([img_bev, lidar_feats], dim=1))
        grid_features = bev.flatten(2, 3).permute(0, 2, 1).reshape(-1, bev.shape[1])
        coords = SECONDV2.create_dense_coord(self.bev_size, self.bev_size, bs).type_as(grid_features).int()
        grid_coords = torch.cat([coords[k].reshape(4, -1).transpose(1, 0) for k in range(bs)], dim=0)
        pts_backbone = kwargs['pts_backbone']
        ins_hm = None
        feats = []
        self._dbg_fwd = dict(img_bev=img_bev, bev=bev)
        for i in range(len(self.get_regions)):
            x = self.get_regions[i](grid_features, grid_coords, bs)
            x = self.grid2region_att[i](x)
            if i == 0:
                self._dbg_fwd['sst0'] = x[0]
                x[0], ins_hm = self.instance_fusion(bev, x[0], bs)
                self._dbg_fwd['igf'] = x[0]
            grid_features, grid_coords, this_feat = pts_backbone(x, 'stage{}'.format(i + 1))
            feats.append(this_feat)
        return feats, ins_hm
