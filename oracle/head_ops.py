"""Oracle: TransFusionHeadV2 inference forward + box decoding (the consumer of the neck).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED by outputs of the reference's own
mmdet3d/models/dense_heads/transfusion_head_v2.py and mmdet3d/core/bbox/coders/transfusion_bbox_coder.py run in the
build container (tests/golden/ref_head.npz, tests/test_oracle_vs_reference.py).  Restates, eval mode:
  forward_single (transfusion_head_v2.py:771-892): shared 3x3 conv -> heat-map head -> sigmoid + 3x3 local-max NMS
  (classes 8, 9 exempt for nuScenes) -> top-`num_proposals` over class x position -> query = feature + class
  encoding -> decoder layer(s) (self-attention over the queries, cross-attention to the 180x180 BEV tokens, FFN;
  learned positional embeddings of the (x+.5, y+.5) grid) -> per-attribute Conv1d heads; centre += query position.
  get_bboxes / TransFusionBBoxCoder.decode (transfusion_head_v2.py:1278-1310, transfusion_bbox_coder.py:41-125):
  score = sigmoid(heatmap) * query_heatmap_score * one_hot(label), boxes in metres, post_center_range filter.
Parameter names follow the reference so its checkpoints load (state_dict keys checked against the reference module).
"""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from .fusion_ops import ConvModule, MultiheadAttention, PositionEmbeddingLearned


class TransformerDecoderLayer(nn.Module):
    """ref: transfusion_head_v2.py:42-120 (post-norm; positional embeddings added to q, k AND v of the cross attention)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation='relu', self_posembed=None,
                 cross_posembed=None):
        super().__init__()
        self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.activation = {'relu': F.relu, 'gelu': F.gelu}[activation]
        self.self_posembed = self_posembed
        self.cross_posembed = cross_posembed

    def forward(self, query, key, query_pos, key_pos):
        qpe = self.self_posembed(query_pos).permute(2, 0, 1)
        kpe = self.cross_posembed(key_pos).permute(2, 0, 1)
        query = query.permute(2, 0, 1)
        key = key.permute(2, 0, 1)
        q = query + qpe
        query = self.norm1(query + self.self_attn(q, q, q)[0])
        query = self.norm2(query + self.multihead_attn(query + qpe, key + kpe, key + kpe)[0])
        query = self.norm3(query + self.linear2(self.activation(self.linear1(query))))
        return query.permute(1, 2, 0)


class FFN(nn.Module):
    """ref: transfusion_head_v2.py:505-590: per attribute (num_conv-1) x [Conv1d(bias=False) + BN1d + ReLU] + Conv1d."""

    def __init__(self, in_channels, heads, head_conv=64, final_kernel=1):
        super().__init__()
        self.heads = heads
        for head, (classes, num_conv) in heads.items():
            layers, c_in = [], in_channels
            for _ in range(num_conv - 1):
                m = nn.Sequential()
                m.add_module('conv', nn.Conv1d(c_in, head_conv, final_kernel, padding=final_kernel // 2, bias=False))
                m.add_module('bn', nn.BatchNorm1d(head_conv))
                m.add_module('activate', nn.ReLU(inplace=True))
                layers.append(m)
                c_in = head_conv
            layers.append(nn.Conv1d(head_conv, classes, final_kernel, padding=final_kernel // 2, bias=True))
            setattr(self, head, nn.Sequential(*layers))

    def forward(self, x):
        return {head: getattr(self, head)(x) for head in self.heads}


class TransFusionHeadV2(nn.Module):
    """ref: transfusion_head_v2.py:593-892 (inference side)."""

    def __init__(self, num_proposals=128, auxiliary=True, in_channels=128 * 3, hidden_channel=128, num_classes=4,
                 num_decoder_layers=3, num_heads=8, nms_kernel_size=1, ffn_channel=256, dropout=0.1, bn_momentum=0.1,
                 activation='relu', common_heads=dict(), num_heatmap_convs=2, test_cfg=None, bbox_coder=None, **kwargs):
        super().__init__()
        self.num_classes, self.num_proposals, self.auxiliary = num_classes, num_proposals, auxiliary
        self.num_decoder_layers, self.nms_kernel_size, self.test_cfg = num_decoder_layers, nms_kernel_size, test_cfg
        self.shared_conv = nn.Conv2d(in_channels, hidden_channel, 3, padding=1, bias=True)      # bias='auto' is truthy (:655)
        self.heatmap_head = nn.Sequential(ConvModule(hidden_channel, hidden_channel),
                                          nn.Conv2d(hidden_channel, num_classes, 3, padding=1, bias=True))
        self.class_encoding = nn.Conv1d(num_classes, hidden_channel, 1)
        self.decoder = nn.ModuleList([
            TransformerDecoderLayer(hidden_channel, num_heads, ffn_channel, dropout, activation,
                                    self_posembed=PositionEmbeddingLearned(2, hidden_channel),
                                    cross_posembed=PositionEmbeddingLearned(2, hidden_channel))
            for _ in range(num_decoder_layers)])
        self.prediction_heads = nn.ModuleList()
        for _ in range(num_decoder_layers):
            heads = copy.deepcopy(common_heads)
            heads.update(dict(heatmap=(num_classes, num_heatmap_convs)))
            self.prediction_heads.append(FFN(hidden_channel, heads))
        x_size = test_cfg['grid_size'][0] // test_cfg['out_size_factor']
        y_size = test_cfg['grid_size'][1] // test_cfg['out_size_factor']
        bx, by = torch.meshgrid(torch.linspace(0, x_size - 1, x_size), torch.linspace(0, y_size - 1, y_size), indexing='ij')
        self.bev_pos = torch.cat([(bx + 0.5)[None], (by + 0.5)[None]], dim=0)[None].view(1, 2, -1).permute(0, 2, 1)
        self.coder = bbox_coder

    @torch.no_grad()
    def forward_single(self, inputs):
        bs = inputs.shape[0]
        lidar_feat = self.shared_conv(inputs)
        flat = lidar_feat.view(bs, lidar_feat.shape[1], -1)
        bev_pos = self.bev_pos.repeat(bs, 1, 1)
        dense_heatmap = self.heatmap_head(lidar_feat)
        heatmap = dense_heatmap.sigmoid()
        pad = self.nms_kernel_size // 2
        local_max = torch.zeros_like(heatmap)
        inner = F.max_pool2d(heatmap, kernel_size=self.nms_kernel_size, stride=1, padding=0)
        local_max[:, :, pad:(-pad), pad:(-pad)] = inner
        if self.test_cfg['dataset'] == 'nuScenes':
            local_max[:, 8] = heatmap[:, 8]
            local_max[:, 9] = heatmap[:, 9]
        heatmap = heatmap * (heatmap == local_max)
        heatmap = heatmap.view(bs, heatmap.shape[1], -1)
        top = heatmap.view(bs, -1).argsort(dim=-1, descending=True)[..., :self.num_proposals]
        top_class = top // heatmap.shape[-1]
        top_index = top % heatmap.shape[-1]
        query_pos = bev_pos.gather(index=top_index[:, None, :].permute(0, 2, 1).expand(-1, -1, 2), dim=1)
        query_feat = flat.gather(index=top_index[:, None, :].expand(-1, flat.shape[1], -1), dim=-1)
        self.query_labels = top_class
        one_hot = F.one_hot(top_class, num_classes=self.num_classes).permute(0, 2, 1)
        query_feat = query_feat + self.class_encoding(one_hot.float())
        rets = []
        for i in range(self.num_decoder_layers):
            query_feat = self.decoder[i](query_feat, flat, query_pos, bev_pos)
            res = self.prediction_heads[i](query_feat)
            res['center'] = res['center'] + query_pos.permute(0, 2, 1)
            rets.append(res)
            query_pos = res['center'].clone().permute(0, 2, 1)
        rets[0]['query_heatmap_score'] = heatmap.gather(index=top_index[:, None, :].expand(-1, self.num_classes, -1), dim=-1)
        rets[0]['dense_heatmap'] = dense_heatmap
        if not self.auxiliary:
            return [rets[-1]]
        out = {}
        for k in rets[0]:
            out[k] = rets[0][k] if k in ('dense_heatmap', 'query_heatmap_score') else torch.cat([r[k] for r in rets], dim=-1)
        return [out]

    @torch.no_grad()
    def forward(self, feats, img_feats=None, metas=None):
        if isinstance(feats, torch.Tensor):
            feats = [feats]
        assert len(feats) == 1, 'only support one level features.'
        return ([self.forward_single(feats[0])[0]],)

    @torch.no_grad()
    def get_bboxes(self, preds_dicts):
        """-> per sample dict(bboxes [K,9 or 7] (x,y,z_bottom,w,l,h,yaw[,vx,vy]), scores [K], labels [K]); test_cfg
        nms_type=None (config:170), so no NMS."""
        p = preds_dicts[0][0]
        n = self.num_proposals
        score = p['heatmap'][..., -n:].sigmoid()
        one_hot = F.one_hot(self.query_labels, num_classes=self.num_classes).permute(0, 2, 1)
        score = score * p['query_heatmap_score'] * one_hot
        return decode_boxes(self.coder, score, p['rot'][..., -n:].clone(), p['dim'][..., -n:].clone(), p['center'][..., -n:].clone(),
                            p['height'][..., -n:].clone(), p['vel'][..., -n:].clone() if 'vel' in p else None)


def decode_boxes(coder, heatmap, rot, dim, center, height, vel):
    """ref: transfusion_bbox_coder.py:41-125 with filter=True."""
    labels = heatmap.max(1).indices
    scores = heatmap.max(1).values
    osf, vs, pr = coder['out_size_factor'], coder['voxel_size'], coder['pc_range']
    center = torch.stack([center[:, 0] * osf * vs[0] + pr[0], center[:, 1] * osf * vs[1] + pr[1]], 1)
    dim = dim.exp()
    height = height - dim[:, 2:3] * 0.5
    rot = torch.atan2(rot[:, 0:1], rot[:, 1:2])
    parts = [center, height, dim, rot] + ([vel] if vel is not None else [])
    boxes = torch.cat(parts, dim=1).permute(0, 2, 1)
    rng = torch.tensor(coder['post_center_range'], dtype=boxes.dtype)
    mask = (boxes[..., :3] >= rng[:3]).all(2) & (boxes[..., :3] <= rng[3:]).all(2)
    if coder.get('score_threshold'):
        mask &= scores > coder['score_threshold']
    return [dict(bboxes=boxes[i, mask[i]], scores=scores[i, mask[i]], labels=labels[i, mask[i]]) for i in range(boxes.shape[0])]
