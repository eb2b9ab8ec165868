"""Oracle: SST region partition + window attention (rows a8, a9 of SURVEY §8).

TEST INFRASTRUCTURE ONLY.  PINNED by outputs of the reference's own files run in the build container
(tests/golden/ref_sst.npz: window / slot indices, masks, positional table exact; SSTv2 level output to 2e-4;
tests/test_oracle_vs_reference.py).  Literal restatement
of mmdet3d/ops/sst/sst_ops.py:22-287, mmdet3d/models/sst/sst_input_layer_v2.py, sst_basic_block_v2.py and
mmdet3d/models/backbones/sst_v2.py, with the reference's nondeterministic pieces made canonical:
inner-window index = stable rank (the CUDA op hands ranks out in atomic order, ingroup_inds_kernel.cu:17-30) and
the token shuffle (sst_input_layer_v2.py:77-81) optional (outputs are shuffle-invariant up to fp reassociation).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def get_inner_win_inds(win_inds):
    """ref: IngroupIndicesFunction (sst_ops.py:197-217): rank of each voxel inside its window."""
    order = torch.argsort(win_inds, stable=True)
    sorted_w = win_inds[order]
    head = torch.ones_like(sorted_w, dtype=torch.bool)
    head[1:] = sorted_w[1:] != sorted_w[:-1]
    starts = torch.nonzero(head).squeeze(1)
    seg = torch.cumsum(head.long(), 0) - 1
    rank = torch.arange(len(sorted_w)) - starts[seg]
    out = torch.empty_like(win_inds)
    out[order] = rank
    return out


def make_continuous_inds(inds):
    """ref: sst_ops.py:270-287."""
    unique_inds, _ = torch.sort(torch.unique(inds))
    canvas = -torch.ones((int(unique_inds.max()) + 1,), dtype=inds.dtype)
    canvas[unique_inds] = torch.arange(len(unique_inds), dtype=inds.dtype)
    conti = canvas[inds]
    assert conti.max() == len(torch.unique(conti)) - 1 and conti.min() == 0
    return conti


def get_flat2win_inds(batch_win_inds, voxel_drop_lvl, drop_info):
    """ref: sst_ops.py:22-60."""
    out = {}
    for dl in drop_info:
        dl_mask = voxel_drop_lvl == dl
        if not dl_mask.any():
            continue
        conti = make_continuous_inds(batch_win_inds[dl_mask])
        max_tokens = drop_info[dl]['max_tokens']
        inner = get_inner_win_inds(conti)
        f2w = conti * max_tokens + inner
        out[dl] = (f2w, torch.where(dl_mask))
        assert inner.max() < max_tokens
    return out


def get_flat2win_inds_v2(batch_win_inds, voxel_drop_lvl, drop_info):
    d = get_flat2win_inds(batch_win_inds, voxel_drop_lvl, drop_info)
    d['voxel_drop_level'] = voxel_drop_lvl
    d['batching_info'] = drop_info
    return d


def flat2window_v2(feat, inds_dict):
    """ref: sst_ops.py:63-97,133-137."""
    drop_info = inds_dict['batching_info']
    lvl = inds_dict['voxel_drop_level']
    out = {}
    for dl in drop_info:
        dl_mask = lvl == dl
        if not dl_mask.any():
            continue
        this_inds = inds_dict[dl][0]
        max_tokens = drop_info[dl]['max_tokens']
        num_windows = int((this_inds // max_tokens).max()) + 1
        f3 = torch.zeros((num_windows * max_tokens, feat.shape[-1]), dtype=feat.dtype)
        f3[this_inds] = feat[dl_mask]
        out[dl] = f3.reshape(num_windows, max_tokens, feat.shape[-1])
    return out


def window2flat_v2(feat_3d_dict, inds_dict):
    """ref: sst_ops.py:99-127,129-131."""
    inds_v1 = {k: inds_dict[k] for k in inds_dict if not isinstance(k, str)}
    n = sum(v[0].shape[0] for v in inds_v1.values())
    first = feat_3d_dict[list(feat_3d_dict.keys())[0]]
    flat = torch.zeros((n, first.shape[-1]), dtype=first.dtype)
    for dl in feat_3d_dict:
        feat = feat_3d_dict[dl]
        inds, flat_pos = inds_v1[dl]
        flat[flat_pos] = feat.reshape(-1, feat.shape[-1])[inds]
    return flat


def get_window_coors(coors, sparse_shape, window_shape, do_shift):
    """ref: sst_ops.py:219-268."""
    if len(window_shape) == 2:
        win_shape_x, win_shape_y = window_shape
        win_shape_z = sparse_shape[-1]
    else:
        win_shape_x, win_shape_y, win_shape_z = window_shape
    sparse_shape_x, sparse_shape_y, sparse_shape_z = sparse_shape
    max_num_win_x = int(np.ceil((sparse_shape_x / win_shape_x)) + 1)
    max_num_win_y = int(np.ceil((sparse_shape_y / win_shape_y)) + 1)
    max_num_win_z = int(np.ceil((sparse_shape_z / win_shape_z)) + 1)
    max_num_win_per_sample = max_num_win_x * max_num_win_y * max_num_win_z
    if do_shift:
        shift_x, shift_y, shift_z = win_shape_x // 2, win_shape_y // 2, win_shape_z // 2
    else:
        shift_x, shift_y, shift_z = win_shape_x, win_shape_y, win_shape_z
    if sparse_shape_z == win_shape_z:
        shift_z = 0
    sx = coors[:, 3] + shift_x
    sy = coors[:, 2] + shift_y
    sz = coors[:, 1] + shift_z
    wx, wy, wz = sx // win_shape_x, sy // win_shape_y, sz // win_shape_z
    batch_win_inds = coors[:, 0] * max_num_win_per_sample + wx * max_num_win_y * max_num_win_z + wy * max_num_win_z + wz
    coors_in_win = torch.stack([sz % win_shape_z, sy % win_shape_y, sx % win_shape_x], dim=-1)
    return batch_win_inds, coors_in_win


class SSTInputLayerV2(nn.Module):
    """ref: mmdet3d/models/sst/sst_input_layer_v2.py:18-312."""

    def __init__(self, drop_info, window_shape, sparse_shape, shuffle_voxels=True, debug=False, normalize_pos=False,
                 pos_temperature=10000, mute=False, pos_embed=256, **kwargs):
        super().__init__()
        self.drop_info = drop_info
        self.sparse_shape = sparse_shape
        self.shuffle_voxels = shuffle_voxels
        self.window_shape = window_shape
        self.normalize_pos = normalize_pos
        self.pos_temperature = pos_temperature
        self.pos_embed_channels = pos_embed

    @torch.no_grad()
    def forward(self, voxel_feats, voxel_coors, batch_size=None):
        voxel_coors = voxel_coors.long()
        if self.shuffle_voxels:
            shuffle_inds = torch.randperm(len(voxel_feats))
            voxel_feats = voxel_feats[shuffle_inds]
            voxel_coors = voxel_coors[shuffle_inds]
        info = {}
        for i in range(2):
            bwi, ciw = get_window_coors(voxel_coors, self.sparse_shape, self.window_shape, i == 1)
            info[f'batch_win_inds_shift{i}'] = bwi
            info[f'coors_in_win_shift{i}'] = ciw
        info['voxel_feats'] = voxel_feats
        info['voxel_coors'] = voxel_coors
        info = self.drop_voxel(info)
        voxel_feats = info['voxel_feats']
        for i in range(2):
            info[f'flat2win_inds_shift{i}'] = get_flat2win_inds_v2(info[f'batch_win_inds_shift{i}'],
                                                                   info[f'voxel_drop_level_shift{i}'], self.drop_info)
            info[f'pos_dict_shift{i}'] = self.get_pos_embed(info[f'flat2win_inds_shift{i}'],
                                                            info[f'coors_in_win_shift{i}'], self.pos_embed_channels,
                                                            voxel_feats.dtype)
            info[f'key_mask_shift{i}'] = self.get_key_padding_mask(info[f'flat2win_inds_shift{i}'])
        if self.shuffle_voxels:
            info['shuffle_inds'] = shuffle_inds
        return info

    def drop_single_shift(self, batch_win_inds):
        """ref: sst_input_layer_v2.py:112-133."""
        drop_lvl = -torch.ones_like(batch_win_inds)
        inner = get_inner_win_inds(batch_win_inds)
        bincount = torch.bincount(batch_win_inds)
        num_before = bincount[batch_win_inds]
        target = torch.zeros_like(batch_win_inds)
        for dl in self.drop_info:
            max_tokens = self.drop_info[dl]['max_tokens']
            lower, upper = self.drop_info[dl]['drop_range']
            m = (num_before >= lower) & (num_before < upper)
            target[m] = max_tokens
            drop_lvl[m] = dl
        return inner < target, drop_lvl

    def drop_voxel(self, info):
        """ref: sst_input_layer_v2.py:135-211 (two shifts)."""
        s0 = info['batch_win_inds_shift0']
        n_all = s0.shape[0]
        keep_inds = torch.arange(n_all, dtype=torch.long)
        k0, l0 = self.drop_single_shift(s0)
        l0, keep_inds, s0 = l0[k0], keep_inds[k0], s0[k0]
        s1 = info['batch_win_inds_shift1'][k0]
        k1, l1 = self.drop_single_shift(s1)
        l0, keep_inds, s0 = l0[k1], keep_inds[k1], s0[k1]
        l1, s1 = l1[k1], s1[k1]
        info['voxel_keep_inds'] = keep_inds
        info['voxel_drop_level_shift0'], info['batch_win_inds_shift0'] = l0, s0
        info['voxel_drop_level_shift1'], info['batch_win_inds_shift1'] = l1, s1
        n_before = len(info['voxel_coors'])
        info['voxel_feats'] = info['voxel_feats'][keep_inds]
        info['voxel_coors'] = info['voxel_coors'][keep_inds]
        for k, v in info.items():
            if isinstance(v, torch.Tensor) and len(v) == n_before:
                info[k] = v[keep_inds]
        return info

    def get_pos_embed(self, inds_dict, coors_in_win, feat_dim, dtype):
        """ref: sst_input_layer_v2.py:223-290."""
        win_x, win_y = self.window_shape[:2]
        win_z = 0
        ndim = 2
        assert self.window_shape[-1] == 1 or len(self.window_shape) == 2
        y, x = coors_in_win[:, 1] - win_y / 2, coors_in_win[:, 2] - win_x / 2
        pos_length = feat_dim // ndim
        inv_freq = torch.arange(pos_length, dtype=torch.float32)
        inv_freq = self.pos_temperature ** (2 * (inv_freq // 2) / pos_length)
        embed_x = x[:, None] / inv_freq[None, :]
        embed_y = y[:, None] / inv_freq[None, :]
        embed_x = torch.stack([embed_x[:, ::2].sin(), embed_x[:, 1::2].cos()], dim=-1).flatten(1)
        embed_y = torch.stack([embed_y[:, ::2].sin(), embed_y[:, 1::2].cos()], dim=-1).flatten(1)
        pos = torch.cat([embed_x, embed_y], dim=-1).to(dtype)
        assert feat_dim - pos.size(1) == 0
        return flat2window_v2(pos, inds_dict)

    def get_key_padding_mask(self, ind_dict):
        """ref: sst_input_layer_v2.py:292-303."""
        n = len(ind_dict['voxel_drop_level'])
        kp = torch.ones((n, 1)).bool()
        d = flat2window_v2(kp, ind_dict)
        return {k: v.logical_not().squeeze(2) for k, v in d.items()}


class WindowAttention(nn.Module):
    """ref: sst_basic_block_v2.py:14-76."""

    def __init__(self, d_model, nhead, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)

    def forward(self, feat_2d, pos_dict, ind_dict, key_padding_dict):
        out = {}
        f3d = flat2window_v2(feat_2d, ind_dict)
        for name in f3d:
            pos = pos_dict[name].permute(1, 0, 2)
            f = f3d[name].permute(1, 0, 2)
            q = k = f + pos
            o, _ = self.self_attn(q, k, value=f, key_padding_mask=key_padding_dict[name])
            out[name] = o.permute(1, 0, 2)
        return window2flat_v2(out, ind_dict)


class EncoderLayer(nn.Module):
    """ref: sst_basic_block_v2.py:78-126 (post-norm)."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout=0.0, activation='gelu'):
        super().__init__()
        self.win_attn = WindowAttention(d_model, nhead, dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.activation = F.gelu if activation == 'gelu' else F.relu

    def forward(self, src, pos_dict, ind_dict, key_mask_dict):
        src2 = self.win_attn(src, pos_dict, ind_dict, key_mask_dict)
        src = self.norm1(src + src2)
        src2 = self.linear2(self.activation(self.linear1(src)))
        return self.norm2(src + src2)


class BasicShiftBlockV2(nn.Module):
    """ref: sst_basic_block_v2.py:129-169."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout=0.0, activation='gelu'):
        super().__init__()
        self.encoder_list = nn.ModuleList([EncoderLayer(d_model, nhead, dim_feedforward, dropout, activation)
                                           for _ in range(2)])

    def forward(self, src, pos_dict_list, ind_dict_list, key_mask_dict_list):
        out = src
        for i in range(2):
            out = self.encoder_list[i](out, pos_dict_list[i], ind_dict_list[i], key_mask_dict_list[i])
        return out


class SSTv2(nn.Module):
    """ref: mmdet3d/models/backbones/sst_v2.py:11-133."""

    def __init__(self, d_model=[], nhead=[], num_blocks=6, dim_feedforward=[], dropout=0.0, activation='gelu',
                 output_shape=None, debug=True, in_channel=None, checkpoint_blocks=[], layer_cfg=dict()):
        super().__init__()
        if in_channel is not None:
            self.linear0 = nn.Linear(in_channel, d_model[0])
        self.block_list = nn.ModuleList([BasicShiftBlockV2(d_model[i], nhead[i], dim_feedforward[i], dropout, activation)
                                         for i in range(num_blocks)])
        for name, p in self.named_parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.output_shape = output_shape

    @torch.no_grad()
    def forward(self, voxel_info, **kwargs):
        batch_size = int(voxel_info['voxel_coors'][:, 0].max()) + 1
        out = voxel_info['voxel_feats']
        inds = [voxel_info[f'flat2win_inds_shift{i}'] for i in range(2)]
        masks = [voxel_info[f'key_mask_shift{i}'] for i in range(2)]
        poss = [voxel_info[f'pos_dict_shift{i}'] for i in range(2)]
        if hasattr(self, 'linear0'):
            out = self.linear0(out)
        for block in self.block_list:
            out = block(out, poss, inds, masks)
        return [self.recover_bev(out, voxel_info['voxel_coors'], batch_size)]

    def recover_bev(self, voxel_feat, coors, batch_size):
        """ref: sst_v2.py:97-133."""
        ny, nx = self.output_shape
        c = voxel_feat.shape[-1]
        canv = []
        for b in range(batch_size):
            canvas = torch.zeros(c, nx * ny, dtype=voxel_feat.dtype)
            m = coors[:, 0] == b
            tc = coors[m, :]
            idx = (tc[:, 2] * nx + tc[:, 3]).long()
            canvas[:, idx] = voxel_feat[m, :].t()
            canv.append(canvas)
        return torch.stack(canv, 0).view(batch_size, c, ny, nx)
