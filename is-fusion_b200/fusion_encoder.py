"""ISFusionEncoder (HSF + IGF), SSTInputLayerV2 / SSTv2 and SECONDV2 on the B200 kernels.

ref: mmdet3d/models/middle_encoders/fusion_encoder.py:833-1189, mmdet3d/models/sst/sst_input_layer_v2.py,
mmdet3d/models/sst/sst_basic_block_v2.py, mmdet3d/models/backbones/sst_v2.py, backbones/second.py:98-239.
Same registry names, constructor kwargs, forward signature and state_dict keys.

Data layout: every BEV map is channels-last fp16 [B, Y, X, C]; a BEV map IS the SST token matrix
[B*Y*X, C] (token = b*G*G + y*G + x, the order of fusion_encoder.py:1167), so region partition needs no sort /
unique / scatter: window id and in-window slot are closed forms of (y, x).  The [X,Y]
"transposed frame" the reference uses for the heat-map / conv_scene branch (Appendix C #2) is realised by
transposing the 3x3 kernels (conv(x^T, W)^T == conv(x, W^T)) instead of the activations.
"""
import ctypes
import math
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from . import conv2d
from .gemm import gemm, pack_weight
from .profiling import stage
from .registry import (BACKBONES, FUSION_LAYERS, MIDDLE_ENCODERS, ConvModule, build_backbone, build_conv_layer,
                       build_middle_encoder, build_norm_layer)


def _fold_conv_bn(conv, bn, transpose_kernel=False):
    """eval-mode BN folded into an fp16 channels-last conv weight + fp16 bias."""
    w = conv.weight.detach().float()
    if bn is not None:
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale
        w = w * scale.view(-1, 1, 1, 1)
        b = shift if conv.bias is None else shift + conv.bias * scale
    else:
        b = conv.bias if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if transpose_kernel:
        w = w.transpose(2, 3)
    b = b.detach()
    return (w.to(L.DTYPE).contiguous(memory_format=torch.channels_last), b.to(L.DTYPE).contiguous(), b.float().contiguous())   # + fp32 shift


def _conv(x_nhwc, packed, stride=1, relu=True):
    """3x3 conv + folded BN (+ReLU) on a channels-last fp16 map [B,Y,X,C]: the library's own tcgen05 kernels (conv2d.py).
    cuDNN is reachable only as the A/B baseline (ISF_CONV=cudnn)."""
    if conv2d.BACKEND != 'cudnn':
        return conv2d.conv3x3(x_nhwc, packed, stride, relu)
    w, b = packed[0], packed[1]
    x = x_nhwc.permute(0, 3, 1, 2)
    if relu:     # cuDNN runtime-fused conv + bias + ReLU (one kernel)
        y = torch.cudnn_convolution_relu(x, w, b, [stride, stride], [1, 1], [1, 1], 1)
    else:
        y = F.conv2d(x, w, b, stride=stride, padding=1)
    return y.permute(0, 2, 3, 1)        # stays channels-last in memory: a free view


class _Packable(nn.Module):
    """modules that cache folded / packed fp16 weights for the eval-mode CUDA path."""

    def _invalidate(self):
        self.__dict__['_pk'] = None

    def train(self, mode=True):
        self._invalidate()
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        self._invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)


# =====================================================================================================
# SST
# =====================================================================================================
@MIDDLE_ENCODERS.register_module()
class SSTInputLayerV2(nn.Module):
    """Region partition.  On the dense BEV grid of the fusion encoder every quantity the reference derives with
    sorts / uniques / atomics (sst_input_layer_v2.py:62-110, sst_ops.py:22-60,219-287) is analytic, so this
    module only validates its input and hands the token matrix on; SSTv2 applies the closed forms in-kernel.
    `drop_info` with max_tokens >= window area never drops on a dense grid (sst_input_layer_v2.py:121-132)."""

    def __init__(self, drop_info, window_shape, sparse_shape, shuffle_voxels=True, debug=False, normalize_pos=False,
                 pos_temperature=10000, mute=False, pos_embed=256, **kwargs):
        super().__init__()
        self.meta_drop_info = drop_info
        self.sparse_shape = sparse_shape
        self.shuffle_voxels = shuffle_voxels     # the shuffle is output-invariant; not performed
        self.window_shape = window_shape
        self.normalize_pos = normalize_pos
        self.pos_temperature = pos_temperature
        self.pos_embed_channels = pos_embed
        wx, wy = window_shape[0], window_shape[1]
        if (wx, wy) != (6, 6) or normalize_pos or any(d['max_tokens'] < wx * wy for d in drop_info.values()):
            raise L.IsfError('SSTInputLayerV2: built for 6x6 windows, max_tokens >= 36 (isfusion_0075voxel.py:17-21)')

    def pos_table(self, device):
        """[36, C] in-window sinusoid, slot = iy*6 + ix (sst_input_layer_v2.py:245-276); input independent -> cached."""
        cached = self.__dict__.get('_pos_cache')
        if cached is not None and cached.device == device:
            return cached
        t = self._pos_table(device)
        self.__dict__['_pos_cache'] = t
        return t

    def _pos_table(self, device):
        c = self.pos_embed_channels
        pos_length = c // 2
        inv = torch.arange(pos_length, dtype=torch.float32, device=device)
        inv = self.pos_temperature ** (2 * (inv // 2) / pos_length)
        iy, ix = torch.meshgrid(torch.arange(6, device=device), torch.arange(6, device=device), indexing='ij')
        x = (ix.flatten().float() - 3.0)[:, None] / inv[None]
        y = (iy.flatten().float() - 3.0)[:, None] / inv[None]
        ex = torch.stack([x[:, ::2].sin(), x[:, 1::2].cos()], dim=-1).flatten(1)
        ey = torch.stack([y[:, ::2].sin(), y[:, 1::2].cos()], dim=-1).flatten(1)
        return torch.cat([ex, ey], dim=-1)

    def forward(self, voxel_feats, voxel_coors=None, batch_size=None):
        g = self.sparse_shape[0]
        if voxel_feats.shape[0] != batch_size * g * self.sparse_shape[1]:
            raise L.IsfError('SSTInputLayerV2: the B200 path takes the dense BEV token matrix [B*Y*X, C]')
        return dict(voxel_feats=voxel_feats, batch_size=batch_size, grid=g, pos=self.pos_table(voxel_feats.device))


class _WinAttn(nn.Module):
    def __init__(self, d_model, nhead, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)


class EncoderLayer(nn.Module):
    """parameter container with the reference's names (sst_basic_block_v2.py:78-104)."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout=0.0):
        super().__init__()
        self.win_attn = _WinAttn(d_model, nhead, dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class BasicShiftBlockV2(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, dropout=0.0):
        super().__init__()
        self.encoder_list = nn.ModuleList([EncoderLayer(d_model, nhead, dim_feedforward, dropout) for _ in range(2)])


@BACKBONES.register_module()
class SSTv2(_Packable):

    def __init__(self, d_model=[], nhead=[], num_blocks=6, dim_feedforward=[], dropout=0.0, activation='gelu',
                 output_shape=None, debug=True, in_channel=None, checkpoint_blocks=[], layer_cfg=dict()):
        super().__init__()
        if activation != 'gelu' or any(h != 8 for h in nhead[:num_blocks]):
            raise L.IsfError('SSTv2: built for GELU, 8 heads')
        self.d_model, self.nhead = d_model, nhead
        if in_channel is not None:
            self.linear0 = nn.Linear(in_channel, d_model[0])
        self.block_list = nn.ModuleList([BasicShiftBlockV2(d_model[i], nhead[i], dim_feedforward[i], dropout)
                                         for i in range(num_blocks)])
        for name, p in self.named_parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.output_shape = output_shape
        self._pk = None

    def _pack(self, pos):
        if self.__dict__.get('_pk') is None:
            pk = {}
            if hasattr(self, 'linear0'):
                pk['l0'] = (pack_weight(self.linear0.weight.detach()), self.linear0.bias.detach().float().contiguous())
            layers = []
            for blk in self.block_list:
                for enc in blk.encoder_list:
                    mha = enc.win_attn.self_attn
                    e = mha.embed_dim
                    wi = mha.in_proj_weight.detach().float()
                    layers.append(dict(
                        win=pack_weight(wi), bin=mha.in_proj_bias.detach().float().contiguous(),
                        posq=(pos @ wi[:e].t()).contiguous(), posk=(pos @ wi[e:2 * e].t()).contiguous(),
                        wo=pack_weight(mha.out_proj.weight.detach()), bo=mha.out_proj.bias.detach().float().contiguous(),
                        ln1=(enc.norm1.weight.detach().float().contiguous(), enc.norm1.bias.detach().float().contiguous(), enc.norm1.eps),
                        w1=pack_weight(enc.linear1.weight.detach()), b1=enc.linear1.bias.detach().float().contiguous(),
                        w2=pack_weight(enc.linear2.weight.detach()), b2=enc.linear2.bias.detach().float().contiguous(),
                        ln2=(enc.norm2.weight.detach().float().contiguous(), enc.norm2.bias.detach().float().contiguous(), enc.norm2.eps)))
            pk['layers'] = layers
            self.__dict__['_pk'] = pk
        return self.__dict__['_pk']

    @torch.no_grad()
    def forward(self, voxel_info, **kwargs):
        if self.training:
            raise L.IsfError('SSTv2: eval-mode CUDA path only')
        x = voxel_info['voxel_feats']
        bs, g = voxel_info['batch_size'], voxel_info['grid']
        pk = self._pack(voxel_info['pos'])
        dt = pk['layers'][0]['win'].dtype          # fp16, or bf16 when packed under _lib.set_dtype(torch.bfloat16)
        if x.dtype != dt:
            x = x.to(dt)
        x = x.contiguous()
        if 'l0' in pk:
            x = gemm(x, pk['l0'][0], shift=pk['l0'][1])
        c = x.shape[1]
        lib = L.lib()
        for li, ly in enumerate(pk['layers']):
            shifted = li % 2                              # BasicShiftBlockV2: layer 0 aligned, layer 1 shifted
            qkv = gemm(x, ly['win'], shift=ly['bin'])
            att = torch.empty_like(x)
            L.check(L.fn16('isf_window_attention', dt)(L.ptr(qkv), L.ptr(ly['posq']), L.ptr(ly['posk']), bs, g, c, 8, shifted,
                                                       L.ptr(att), L.stream()))
            x = gemm(att, ly['wo'], shift=ly['bo'], residual=x, ln=ly['ln1'])
            h = gemm(x, ly['w1'], shift=ly['b1'], act='gelu')
            x = gemm(h, ly['w2'], shift=ly['b2'], residual=x, ln=ly['ln2'])
        # recover_bev (sst_v2.py:97-133) is the identity on a dense grid: [B*Y*X, C] -> [B, C, Y, X] view
        return [x.view(bs, g, g, c).permute(0, 3, 1, 2)]


# =====================================================================================================
# SECONDV2
# =====================================================================================================
@BACKBONES.register_module()
class SECONDV2(_Packable):
    """ref: backbones/second.py:98-239 (staged execution; BN folded, cuDNN channels-last fp16 convs)."""

    def __init__(self, in_channels=128, out_channels=[128, 128, 256], layer_nums=[3, 5, 5], layer_strides=[2, 2, 2],
                 norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False),
                 init_cfg=None, pretrained=None):
        super().__init__()
        assert len(layer_strides) == len(layer_nums) == len(out_channels)
        in_filters = [in_channels, *out_channels[:-1]]
        blocks = []
        for i, layer_num in enumerate(layer_nums):
            if layer_strides[i] == 2:
                self.ds_layer = nn.Sequential(
                    build_conv_layer(conv_cfg, in_filters[i], out_channels[i], 3, stride=layer_strides[i], padding=1),
                    build_norm_layer(norm_cfg, out_channels[i])[1], nn.ReLU(inplace=True))
                block = []
            else:
                block = [build_conv_layer(conv_cfg, in_filters[i], out_channels[i], 3, stride=layer_strides[i], padding=1),
                         build_norm_layer(norm_cfg, out_channels[i])[1], nn.ReLU(inplace=True)]
            for _ in range(layer_num):
                block.append(build_conv_layer(conv_cfg, out_channels[i], out_channels[i], 3, padding=1))
                block.append(build_norm_layer(norm_cfg, out_channels[i])[1])
                block.append(nn.ReLU(inplace=True))
            blocks.append(nn.Sequential(*block))
        self.blocks = nn.ModuleList(blocks)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        self._pk = None

    def _pack(self):
        if self.__dict__.get('_pk') is None:
            def seq(s):
                mods = list(s)
                return [(_fold_conv_bn(mods[i], mods[i + 1]), mods[i].stride[0]) for i in range(0, len(mods), 3)]
            self.__dict__['_pk'] = dict(b0=seq(self.blocks[0]), ds=seq(self.ds_layer), b1=seq(self.blocks[1]))
        return self.__dict__['_pk']

    @staticmethod
    def _run(x, convs):
        for packed, stride in convs:
            x = _conv(x, packed, stride=stride)
        return x

    @torch.no_grad()
    def forward(self, x, stage=None):
        if self.training:
            raise L.IsfError('SECONDV2: eval-mode CUDA path only')
        pk = self._pack()
        nhwc = lambda t: t.permute(0, 2, 3, 1)       # NCHW-logical (channels-last memory) -> [B,Y,X,C] view
        if stage == 'stage1':
            f = self._run(nhwc(x[0]), pk['b0'])
            ds = self._run(f, pk['ds'])
            n, h, w, c = ds.shape
            return ds.reshape(n * h * w, c), None, f.permute(0, 3, 1, 2)
        if stage == 'stage2':
            return None, None, self._run(nhwc(x[0]), pk['b1']).permute(0, 3, 1, 2)
        f = self._run(nhwc(x), pk['b0'])
        f2 = self._run(self._run(f, pk['ds']), pk['b1'])
        return (f.permute(0, 3, 1, 2), f2.permute(0, 3, 1, 2))


# =====================================================================================================
# SECONDFPN (pts_neck)
# =====================================================================================================
from .registry import NECKS  # noqa: E402


@NECKS.register_module()
class SECONDFPN(_Packable):
    """ref: mmdet3d/models/necks/second_fpn.py:11-98 for the isfusion_0075voxel.py:105-112 configuration: a 1x1 conv
    (level 0) and a 2x2 stride-2 deconv (level 1), each + BN + ReLU, concatenated and transposed to [X,Y].
    Both are GEMMs over the BEV token matrix, so the whole neck is five launches of the tcgen05 GEMM with the BN
    fold + ReLU in the epilogue and the concat / pixel-shuffle / transpose done by the epilogue's row scatter
    straight into the channels-last [B, X, Y, 512] output."""

    def __init__(self, in_channels=[128, 128, 256], out_channels=[256, 256, 256], upsample_strides=[1, 2, 4],
                 norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), upsample_cfg=dict(type='deconv', bias=False),
                 conv_cfg=dict(type='Conv2d', bias=False), use_conv_for_no_stride=False, init_cfg=None):
        super().__init__()
        if list(upsample_strides) != [1, 2] or not use_conv_for_no_stride:
            raise L.IsfError('SECONDFPN: built for upsample_strides=[1,2], use_conv_for_no_stride=True')
        self.in_channels, self.out_channels = in_channels, out_channels
        blocks = []
        for i, oc in enumerate(out_channels):
            st = upsample_strides[i]
            up = (nn.ConvTranspose2d(in_channels[i], oc, st, stride=st, bias=False) if st > 1 else
                  nn.Conv2d(in_channels[i], oc, 1, stride=1, bias=False))
            blocks.append(nn.Sequential(up, build_norm_layer(norm_cfg, oc)[1], nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(blocks)
        self._pk = None

    def _pack(self):
        if self.__dict__.get('_pk') is None:
            from .spconv import fold_bn
            c0, b0 = self.deblocks[0][0], self.deblocks[0][1]
            c1, b1 = self.deblocks[1][0], self.deblocks[1][1]
            pk = dict(w0=pack_weight(c0.weight.detach()[:, :, 0, 0]), sb0=fold_bn(b0), sb1=fold_bn(b1),
                      w1=[pack_weight(c1.weight.detach()[:, :, dy, dx].t().contiguous()) for dy in range(2) for dx in range(2)])
            self.__dict__['_pk'] = pk
            self.__dict__['_rows'] = {}
        return self.__dict__['_pk']

    def _row_maps(self, bs, g, device):
        """output row (in the transposed [B, X, Y] frame) of every input token."""
        key = (bs, g, device)
        rows = self.__dict__['_rows'].get(key)
        if rows is None:
            b = torch.arange(bs, device=device).view(bs, 1, 1)
            y = torch.arange(g, device=device).view(1, g, 1)
            x = torch.arange(g, device=device).view(1, 1, g)
            r0 = ((b * g + x) * g + y).reshape(-1).int().contiguous()                 # level 0: (b,y,x) -> (b,x,y)
            h = g // 2
            yh = torch.arange(h, device=device).view(1, h, 1)
            xh = torch.arange(h, device=device).view(1, 1, h)
            r1 = [((b * g + 2 * xh + dx) * g + 2 * yh + dy).reshape(-1).int().contiguous()
                  for dy in range(2) for dx in range(2)]
            rows = (r0, r1)
            self.__dict__['_rows'][key] = rows
        return rows

    @torch.no_grad()
    def forward(self, x, swin_format=False, **kwargs):
        """x = [feat_0 [B,128,G,G], feat_1 [B,256,G/2,G/2]] (channels-last memory) -> [out [B,512,G,G]] in the [X,Y] frame."""
        if self.training:
            raise L.IsfError('SECONDFPN: eval-mode CUDA path only')
        pk = self._pack()
        f0, f1 = x[0].permute(0, 2, 3, 1), x[1].permute(0, 2, 3, 1)
        bs, g = f0.shape[0], f0.shape[1]
        t0 = f0.reshape(bs * g * g, f0.shape[-1]).half().contiguous()
        t1 = f1.reshape(bs * (g // 2) ** 2, f1.shape[-1]).half().contiguous()
        oc0, oc1 = self.out_channels[0], self.out_channels[1]
        out = torch.empty((bs * g * g, oc0 + oc1), dtype=torch.float16, device=t0.device)
        r0, r1 = self._row_maps(bs, g, t0.device)
        gemm(t0, pk['w0'], scale=pk['sb0'][0], shift=pk['sb0'][1], act='relu', out=out, out_rows=r0, ldo=oc0 + oc1)
        up = out[:, oc0:]                                   # column offset: same rows, pitch oc0+oc1
        for w, r in zip(pk['w1'], r1):
            gemm(t1, w, scale=pk['sb1'][0], shift=pk['sb1'][1], act='relu', out=up, out_rows=r, ldo=oc0 + oc1)
        return [out.view(bs, g, g, oc0 + oc1).permute(0, 3, 1, 2)]


# =====================================================================================================
# IGF parameter containers (names of fusion_encoder.py:44-171, 173-189, 472-502, 505-674, 768-830)
# =====================================================================================================
class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class PositionEmbeddingLearned(nn.Module):
    def __init__(self, input_channel, num_pos_feats=288):
        super().__init__()
        self.position_embedding_head = nn.Sequential(nn.Conv1d(input_channel, num_pos_feats, kernel_size=1),
                                                     nn.BatchNorm1d(num_pos_feats), nn.ReLU(inplace=True),
                                                     nn.Conv1d(num_pos_feats, num_pos_feats, kernel_size=1))

    def forward(self, xyz):      # [B, N, 2] -> [B, C, N]   (tiny; eval BN through torch)
        return self.position_embedding_head(xyz.transpose(1, 2).contiguous())


class Instane2SceneAtt(nn.Module):
    def __init__(self, d_model, nhead=8, dropout=0.1):
        super().__init__()
        self.multihead_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
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


def _msda_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    b, s, h, d = value.shape
    _, q, _, lv, p, _ = sampling_locations.shape
    out = torch.empty((b, q, h * d), dtype=torch.float32, device=value.device)
    L.check(L.lib().isf_ms_deform_attn_forward(
        L.ptr(value), 1 if value.dtype == torch.float16 else 0, L.ptr(spatial_shapes), L.ptr(level_start_index),
        L.ptr(sampling_locations), L.ptr(attention_weights), b, s, h, d, lv, q, p, L.ptr(out), L.stream()))
    return out


class MultiScaleDeformableAttnFunction(torch.autograd.Function):
    """ref: MultiScaleDeformableAttnFunction_fp32 (multi_scale_deformable_attn_function.py:90-171): forward = K9,
    backward = K9b (isf_ms_deform_attn_backward) — gradients of value, sampling locations and attention weights."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
        ctx.save_for_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
        return _msda_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lvl_start, loc, aw = ctx.saved_tensors
        b, s, h, d = value.shape
        _, q, _, lv, p, _ = loc.shape
        grad_value = torch.zeros(value.shape, dtype=torch.float32, device=value.device)     # accumulated into
        grad_loc = torch.empty_like(loc)
        grad_aw = torch.empty_like(aw)
        L.check(L.lib().isf_ms_deform_attn_backward(
            L.ptr(value), 1 if value.dtype == torch.float16 else 0, L.ptr(shapes), L.ptr(lvl_start), L.ptr(loc), L.ptr(aw),
            L.ptr(grad_output.float().contiguous()), b, s, h, d, lv, q, p, L.ptr(grad_value), L.ptr(grad_loc), L.ptr(grad_aw),
            L.stream()))
        return grad_value.to(value.dtype), None, None, grad_loc, grad_aw


def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step=64):
    """ref: MultiScaleDeformableAttnFunction_fp32 (multi_scale_deformable_attn_function.py:90-171) ->
    mmcv._ext.ms_deform_attn_forward / _backward.  value [B,S,H,D] fp32|fp16 -> [B,Q,H*D] fp32; differentiable in value,
    sampling_locations and attention_weights (the autograd function is only entered when a gradient is asked for)."""
    L.require_cuda(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    args = (value.contiguous(), spatial_shapes.long().contiguous(), level_start_index.long().contiguous(),
            sampling_locations.float().contiguous(), attention_weights.float().contiguous())
    if torch.is_grad_enabled() and any(t.requires_grad for t in (value, sampling_locations, attention_weights)):
        return MultiScaleDeformableAttnFunction.apply(*args)
    return _msda_forward(*args)


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation='relu', n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    def forward(self, tgt, query_pos, reference_points, value, spatial_shapes, level_start_index):
        """200-query decoder layer; `value` = value_proj(src) [B,S,H,D] fp16 precomputed on the tcgen05 GEMM."""
        q = k = tgt + query_pos
        tgt2 = self.self_attn(q.transpose(0, 1), k.transpose(0, 1), tgt.transpose(0, 1))[0].transpose(0, 1)
        tgt = self.norm2(tgt + tgt2)
        ca = self.cross_attn
        qq = tgt + query_pos
        n, lq, _ = qq.shape
        so = ca.sampling_offsets(qq).view(n, lq, ca.n_heads, ca.n_levels, ca.n_points, 2)
        aw = F.softmax(ca.attention_weights(qq).view(n, lq, ca.n_heads, ca.n_levels * ca.n_points), -1)
        aw = aw.view(n, lq, ca.n_heads, ca.n_levels, ca.n_points)
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + so / norm[None, None, None, :, None, :]
        tgt2 = ca.output_proj(ms_deform_attn(value, spatial_shapes, level_start_index, loc, aw))
        tgt = self.norm1(tgt + tgt2)
        tgt2 = self.linear2(F.relu(self.linear1(tgt)))
        return self.norm3(tgt + tgt2)


# a11: one fused launch (csrc/decoder.cu) instead of ~60 small torch launches; ISF_DECODER=torch keeps the module-by-module
# path (the second statement of the same arithmetic, used by the parity test)
FUSED_DECODER = os.environ.get('ISF_DECODER', 'fused') == 'fused'


def _frag_pack(w):
    """[N, 128] weight (nn.Linear layout) -> fp16 in the B-fragment order of mma.sync.m16n8k16 per 8 output rows: lane
    (n % 8) * 4 + l4 holds, contiguously, the halves k = ks*16 + j*8 + l4*2 + {0,1} for ks = 0..7, j = 0..1 (csrc/decoder.cu gemm32)."""
    n, k = w.shape
    assert k == 128 and n % 8 == 0
    v = w.detach().half().view(n // 8, 8, 8, 2, 4, 2)          # (ntile, n_in, ks, j, l4, pair)
    return v.permute(0, 1, 4, 2, 3, 5).reshape(n, 128).contiguous()


def pack_decoder(ia, mha):
    """packed fp16 weights + fp32 parameters of isf_ins_context_decoder for InsContextAtt `ia` (+ the k / v rows of the
    instance->scene MultiheadAttention `mha`); row / offset plan = the constants of csrc/decoder.cu (namespace dec)."""
    c = ia.layers[0].norm1.normalized_shape[0]
    pe = ia.query_pos_embed.position_embedding_head
    conv1, bn, conv2 = pe[0], pe[1], pe[3]
    s1 = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    t1 = bn.bias - bn.running_mean * s1 + conv1.bias * s1
    rows = [_frag_pack(conv2.weight.detach()[:, :, 0])]
    fpar = [conv1.weight.detach()[:, :, 0].reshape(-1), s1.detach(), t1.detach(), conv2.bias.detach()]
    for l in ia.layers:
        sa, ca = l.self_attn, l.cross_attn
        rows += [_frag_pack(sa.in_proj_weight.detach()[:2 * c]), _frag_pack(sa.in_proj_weight.detach()[2 * c:]),
                 _frag_pack(sa.out_proj.weight.detach()),
                 _frag_pack(torch.cat([ca.sampling_offsets.weight.detach(), ca.attention_weights.weight.detach()], 0)),
                 _frag_pack(ca.output_proj.weight.detach()), _frag_pack(l.linear1.weight.detach()), _frag_pack(l.linear2.weight.detach())]
        fpar += [sa.in_proj_bias.detach()[:2 * c], sa.in_proj_bias.detach()[2 * c:], sa.out_proj.bias.detach(),
                 l.norm2.weight.detach(), l.norm2.bias.detach(),
                 torch.cat([ca.sampling_offsets.bias.detach(), ca.attention_weights.bias.detach()], 0), ca.output_proj.bias.detach(),
                 l.norm1.weight.detach(), l.norm1.bias.detach(), l.linear1.bias.detach(), l.linear2.bias.detach(),
                 l.norm3.weight.detach(), l.norm3.bias.detach()]
    rows.append(_frag_pack(mha.in_proj_weight.detach()[c:]))
    fpar.append(mha.in_proj_bias.detach()[c:])
    wp = torch.cat(rows, 0).contiguous()
    fp = torch.cat([t.float().reshape(-1) for t in fpar]).contiguous()
    nb, nf = ctypes.c_size_t(0), ctypes.c_size_t(0)
    L.check(L.lib().isf_ins_context_decoder_sizes(len(ia.layers), ctypes.byref(nb), ctypes.byref(nf)))
    if wp.numel() * 2 != nb.value or fp.numel() != nf.value:
        raise L.IsfError('pack_decoder: layout mismatch with the library (%d / %d bytes, %d / %d floats)'
                         % (wp.numel() * 2, nb.value, fp.numel(), nf.value))
    return wp, fp


class InsContextAtt(nn.Module):
    def __init__(self, num_layers=1, embed_dims=128, bev_size=180, n_points=16):
        super().__init__()
        import copy
        self.bev_size = bev_size
        self.num_feature_levels = 1
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


# =====================================================================================================
# ISFusionEncoder
# =====================================================================================================
@FUSION_LAYERS.register_module()
class ISFusionEncoder(_Packable):

    def __init__(self, num_points_in_pillar=10, embed_dims=256, num_classes=10, **kwargs):
        super().__init__()
        self.num_points_in_pillar = num_points_in_pillar
        self.bev_size = kwargs.get('bev_size', 180)
        self.num_views = kwargs.get('num_views', 6)
        region_shape = kwargs.get('region_shape', None)
        grid_size = kwargs.get('grid_size', None)
        region_drop_info = kwargs.get('region_drop_info', None)
        self.random_noise = 1.0
        self.embed_dims = embed_dims
        self.num_classes = num_classes
        bn2d = dict(type='BN2d')
        self.conv_fusion = ConvModule(embed_dims * 3, embed_dims // 2, kernel_size=3, padding=1, norm_cfg=bn2d)
        self.get_regions = nn.ModuleList()
        self.grid2region_att = nn.ModuleList()
        for l in range(len(region_shape)):
            d = embed_dims // 2 * (l + 1)
            self.get_regions.append(build_middle_encoder(dict(
                type='SSTInputLayerV2', window_shape=region_shape[l], sparse_shape=grid_size[l], shuffle_voxels=True,
                drop_info=region_drop_info[l], pos_temperature=1000, normalize_pos=False, pos_embed=d)))
            self.grid2region_att.append(build_backbone(dict(
                type='SSTv2', d_model=[d] * 4, nhead=[8] * 4, num_blocks=1, dim_feedforward=[d] * 4,
                output_shape=grid_size[l][:2], in_channel=embed_dims // 2 if l == 0 else None)))
        self.instance_num = kwargs.get('instance_num', 200)
        self.nms_kernel_size = 3
        bx, by = torch.meshgrid(torch.linspace(0, self.bev_size - 1, self.bev_size),
                                torch.linspace(0, self.bev_size - 1, self.bev_size), indexing='ij')
        self.bev_pos = torch.cat([(bx + 0.5)[None], (by + 0.5)[None]], dim=0)[None].view(1, 2, -1).permute(0, 2, 1)
        h = embed_dims // 2
        self.conv_ins = ConvModule(h, h, kernel_size=3, padding=1, norm_cfg=bn2d)
        self.conv_scene = ConvModule(h, h, kernel_size=3, padding=1, norm_cfg=bn2d)
        self.conv_heatmap = ConvModule(h, h, kernel_size=3, padding=1, norm_cfg=bn2d)
        self.heatmap_head_1 = ConvModule(h, embed_dims // 4, kernel_size=3, padding=1, norm_cfg=bn2d)
        self.heatmap_head_2 = ConvModule(embed_dims // 4, embed_dims // 4, kernel_size=3, padding=1, norm_cfg=bn2d)
        self.heatmap_head_3 = nn.Conv2d(embed_dims // 4, num_classes, kernel_size=3, stride=1, padding=1)
        self.instance_att = InsContextAtt(num_layers=2, embed_dims=h, bev_size=self.bev_size)
        self.instance_to_scene_att = Instane2SceneAtt(d_model=h)
        self._pk = None

    # ------------------------------------------------------------------ packing
    def _pack(self, device):
        if self.__dict__.get('_pk') is None:
            pk = {}
            pk['fusion'] = _fold_conv_bn(self.conv_fusion.conv, self.conv_fusion.bn)
            pk['ins'] = _fold_conv_bn(self.conv_ins.conv, self.conv_ins.bn)
            # [X,Y]-frame branches run on the [Y,X] map with transposed kernels (Appendix C #2)
            pk['scene'] = _fold_conv_bn(self.conv_scene.conv, self.conv_scene.bn, True)
            pk['hm0'] = _fold_conv_bn(self.conv_heatmap.conv, self.conv_heatmap.bn, True)
            pk['hm1'] = _fold_conv_bn(self.heatmap_head_1.conv, self.heatmap_head_1.bn, True)
            pk['hm2'] = _fold_conv_bn(self.heatmap_head_2.conv, self.heatmap_head_2.bn, True)
            pk['hm3'] = _fold_conv_bn(self.heatmap_head_3, None, True)
            g = self.bev_size
            bev_pos = self.bev_pos.to(device)                                      # [1, G*G, 2]
            ia = self.instance_att
            key_pos = ia.key_pos_embed(bev_pos / ia.bev_size).permute(0, 2, 1)    # [1, N, C] input independent
            pk['key_pos'] = key_pos[0].half().contiguous()
            pk['bev_pos'] = bev_pos
            wv = torch.cat([l.cross_attn.value_proj.weight.detach() for l in ia.layers], 0)
            bv = torch.cat([l.cross_attn.value_proj.bias.detach() for l in ia.layers], 0)
            pk['value_w'] = pack_weight(wv)
            pk['value_b'] = bv.float().contiguous()
            mha = self.instance_to_scene_att.multihead_attn
            e = mha.embed_dim
            pk['i2s_wq'] = pack_weight(mha.in_proj_weight.detach()[:e])
            pk['i2s_bq'] = mha.in_proj_bias.detach()[:e].float().contiguous()
            pk['i2s_wkv'] = mha.in_proj_weight.detach()[e:].float()
            pk['i2s_bkv'] = mha.in_proj_bias.detach()[e:].float()
            pk['i2s_wo'] = pack_weight(mha.out_proj.weight.detach())
            pk['i2s_bo'] = mha.out_proj.bias.detach().float().contiguous()
            n = self.instance_to_scene_att.norm
            pk['i2s_ln'] = (n.weight.detach().float().contiguous(), n.bias.detach().float().contiguous(), n.eps)
            pk['dec_w'], pk['dec_f'] = pack_decoder(ia, mha)
            pk['shapes'] = torch.tensor([[g, g]], dtype=torch.long, device=device)
            pk['lvl0'] = torch.zeros(1, dtype=torch.long, device=device)
            self.__dict__['_pk'] = pk
        return self.__dict__['_pk']

    def new_fusion_buffer(self, bs, device):
        """zeroed channels-last [B,G,G,768] buffer: image BEV in channels 0..255, LiDAR BEV in 256..767."""
        g = self.bev_size
        return torch.zeros((bs, g, g, self.embed_dims * 3), dtype=L.DTYPE, device=device)

    # ------------------------------------------------------------------ a6: camera -> BEV
    def camera_projection(self, lidar2img, img_aug_matrix, lidar_aug_matrix, bs, device):
        """host-side 3x3 algebra of img_point_sampling (:976-1006) folded into 20 floats per (frame, camera):
        q = M p + t with M = lidar2img[:3,:3].inv(lidar_aug[:3,:3]), then the image-aug affine rows."""
        l2i = torch.as_tensor(lidar2img).float().cpu().reshape(bs, self.num_views, 4, 4)
        aug = torch.as_tensor(img_aug_matrix).float().cpu().reshape(bs, self.num_views, 4, 4)
        laug = torch.as_tensor(lidar_aug_matrix).float().cpu().reshape(bs, 4, 4)
        rinv = torch.inverse(laug[:, :3, :3])
        m = l2i[:, :, :3, :3] @ rinv[:, None]
        t = l2i[:, :, :3, 3] - (m @ laug[:, None, :3, 3, None]).squeeze(-1)
        proj = torch.cat([m.reshape(bs, self.num_views, 9), t, aug[:, :, :2, :3].reshape(bs, self.num_views, 6),
                          aug[:, :, :2, 3]], -1).contiguous()
        return proj.to(device)

    def img_fv_to_bev(self, mlvl_feats, bs, out=None, **kwargs):
        """ref: img_point_sampling + img_fv_to_bev (fusion_encoder.py:965-1070) as one fused kernel.
        Returns the channels-last fp16 buffer [B, G, G, ldo] whose first 256 channels hold the image BEV.
        pts_metas carries either the reference's concatenated pillars/pillar_coors or, on the sync-free path,
        'pillars_list' = [(pillars [cap,P,nf], coors [cap,3], device count)] per sample."""
        pm = kwargs['pts_metas']
        feat = mlvl_feats[0]
        dev = feat.device
        c, fh, fw = feat.shape[1], feat.shape[2], feat.shape[3]
        dt = out.dtype if out is not None else L.DTYPE
        feat_nhwc = feat.view(bs, self.num_views, c, fh, fw).permute(0, 1, 3, 4, 2).contiguous().to(dt)
        image_size = kwargs['img_metas'][0]['input_shape']
        proj = kwargs.get('cam_proj')
        if proj is None:
            proj = self.camera_projection(kwargs['lidar2img'], kwargs['img_aug_matrix'], kwargs['lidar_aug_matrix'], bs, dev)
        g = self.bev_size
        if out is None:
            out = torch.zeros((bs, g, g, self.embed_dims), dtype=dt, device=dev)
        if 'pillars_list' in pm:
            jobs = [(p, co, cnt, 3, b) for b, (p, co, cnt) in enumerate(pm['pillars_list'])]
        else:
            jobs = [(pm['pillars'].contiguous(), pm['pillar_coors'].contiguous(), None, 4, 0)]
        for pillars, pcoors, cnt, cdim, b in jobs:
            L.require_cuda(pillars, pcoors)
            L.check(L.fn16('isf_cam2bev', dt)(L.ptr(pillars), pillars.shape[0], L.ptr(cnt), pillars.shape[1], pillars.shape[2],
                                              L.ptr(pcoors), cdim, b, L.ptr(feat_nhwc), L.ptr(proj), bs, self.num_views, fh, fw,
                                              c, ctypes.c_float(image_size[1]), ctypes.c_float(image_size[0]), g, g, L.ptr(out),
                                              out.shape[-1], 0, L.stream()))
        return out

    # ------------------------------------------------------------------ a10-a12: instance-guided fusion
    def instance_fusion(self, bev, scene, bs, **kwargs):
        """bev, scene: channels-last fp16 [B, G, G, 128] ([Y,X] frame) -> (fused scene [B,G,G,128], ins_heatmap)."""
        pk = self._pack(bev.device)
        g, c, k = self.bev_size, bev.shape[-1], self.instance_num
        lib = L.lib()
        dev = bev.device
        # heat-map branch (a10)
        hm = _conv(_conv(_conv(bev, pk['hm0']), pk['hm1']), pk['hm2'])
        logits = _conv(hm, pk['hm3'], relu=False).permute(0, 3, 1, 2).float().contiguous()      # [B,10,Y,X]
        ws_bytes = ctypes.c_size_t(0)
        L.check(lib.isf_heatmap_topk_workspace_size(bs, self.num_classes, g, g, ctypes.byref(ws_bytes)))
        ws = L.workspace(ws_bytes.value, dev)
        top_yx = torch.empty((bs, k), dtype=torch.int32, device=dev)
        plain_from = 8 if self.num_views == 6 else self.num_classes      # classes 8,9: 1x1 NMS (:1107-1115)
        L.check(lib.isf_heatmap_topk(L.ptr(logits), bs, self.num_classes, g, g, plain_from, k, L.ptr(top_yx),
                                     L.ptr(None), L.ptr(ws), ctypes.c_size_t(ws.numel()), L.stream()))
        ty, tx = top_yx // g, top_yx % g                        # n = y*G + x  (reference index t = x*G + y)
        # conv_scene in the [Y,X] frame == the reference's x_scene transposed back (InsContextAtt :797)
        x_scene = _conv(bev, pk['scene'])                                                     # [B,Y,X,C]
        toks = x_scene.reshape(bs, g * g, c)
        ia = self.instance_att
        src = (toks + pk['key_pos'][None]).reshape(bs * g * g, c)
        value = gemm(src, pk['value_w'], shift=pk['value_b'])                                 # both layers: [T, 2C]
        if FUSED_DECODER:
            # a11 in one launch: query position embedding + both decoder layers + the k | v projection of the i2s attention
            # (reference: query_pos = bev_pos[t] = (x+.5, y+.5); query_pos_new swaps -> (y+.5, x+.5); / bev_size — in-kernel)
            out = torch.empty((bs, k, c), dtype=torch.float32, device=dev)
            kv = torch.empty((bs, k, 2 * c), dtype=torch.float16, device=dev)
            toks_c = toks if toks.is_contiguous() else toks.contiguous()
            L.check(lib.isf_ins_context_decoder(L.ptr(toks_c), L.ptr(top_yx), L.ptr(value), value.stride(0), L.ptr(pk['dec_w']),
                                                L.ptr(pk['dec_f']), bs, g, k, c, 8, ia.layers[0].cross_attn.n_points, len(ia.layers),
                                                L.ptr(out), L.ptr(kv), L.stream()))
        else:
            tl = top_yx.long()
            ref_pts = torch.stack([ty.float() + 0.5, tx.float() + 0.5], -1) / g
            x_ins = torch.gather(toks, 1, tl[..., None].expand(-1, -1, c)).float()            # [B,K,C]
            qpe = ia.query_pos_embed(ref_pts).permute(0, 2, 1)                                # [B,K,C]
            out = x_ins
            for li, layer in enumerate(ia.layers):
                v = value[:, li * c:(li + 1) * c].reshape(bs, g * g, 8, c // 8)
                out = layer(out, qpe, ref_pts[:, :, None], v.contiguous(), pk['shapes'], pk['lvl0'])
            kv = F.linear(out, pk['i2s_wkv'], pk['i2s_bkv']).half().contiguous()              # [B,K,2C]
        # instance -> scene attention (a12)
        ins_tok = _conv(bev, pk['ins']).reshape(bs * g * g, c)                                # query tokens
        q = gemm(ins_tok, pk['i2s_wq'], shift=pk['i2s_bq'])
        att = torch.empty_like(q)
        L.check(lib.isf_i2s_attention(L.ptr(q), L.ptr(kv), bs, g * g, k, c, 8, L.ptr(att), L.stream()))
        query_ins = gemm(att, pk['i2s_wo'], shift=pk['i2s_bo'], residual=ins_tok, ln=pk['i2s_ln'])
        fused = torch.empty((bs, g, g, c), dtype=torch.float16, device=dev)
        scene = scene.contiguous()
        cab = ctypes.c_size_t(0)
        L.check(lib.isf_channel_attention_workspace_size(bs, g, c, ctypes.byref(cab)))
        caws = torch.empty(cab.value, dtype=torch.uint8, device=dev)
        L.check(lib.isf_channel_attention(L.ptr(scene), c, L.ptr(query_ins), c, bs, g, c, L.ptr(fused), c, L.ptr(caws),
                                          ctypes.c_size_t(caws.numel()), L.stream()))
        ins_heatmap = logits.transpose(2, 3)                  # back to the reference's [X,Y] frame
        self._dbg = dict(top_idx=(tx * g + ty).long(), x_ins=out)
        return fused, ins_heatmap

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, img_mlvl_feats, lidar_feats, bs, **kwargs):
        """ref: fusion_encoder.py:1152-1189.  lidar_feats [B,512,G,G] (any layout) or, on the fast path,
        kwargs['lidar_sparse'] = SparseConvTensor of conv_out written straight into the fused buffer."""
        if self.training:
            raise L.IsfError('ISFusionEncoder: eval-mode CUDA path only')
        dev = img_mlvl_feats[1].device
        pk = self._pack(dev)
        g = self.bev_size
        cat = kwargs.get('img_bev_buffer')      # camera->BEV already done on a side stream (detector fast path)
        if cat is None:
            with stage('a6 cam2bev'):
                cat = self.new_fusion_buffer(bs, dev)
                self.img_fv_to_bev([img_mlvl_feats[1]], bs, out=cat, **kwargs)
        sp = kwargs.get('lidar_sparse')
        if sp is not None:
            d = sp.spatial_shape[0]
            L.check(L.lib().isf_sparse_to_bev_nhwc(L.ptr(sp.features), L.ptr(sp.indices), sp.features.shape[0],
                                                   L.ptr(sp.num_dev), sp.features.shape[1], d, g, g, L.ptr(cat), cat.shape[-1],
                                                   self.embed_dims, L.stream()))
        else:
            cat[..., self.embed_dims:] = lidar_feats.permute(0, 2, 3, 1)
        with stage('a7 conv_fusion'):
            bev = _conv(cat, pk['fusion'])                                 # [B,G,G,128]
        pts_backbone = kwargs.get('pts_backbone', None)
        grid_features = bev.reshape(bs * g * g, -1)
        ins_hm = None
        return_feats = []
        for i in range(len(self.get_regions)):
            with stage('a8+a9 sst level %d' % i):
                x = self.get_regions[i](grid_features, None, bs)
                x = self.grid2region_att[i](x)
            if i == 0:
                with stage('a10-a13 instance fusion'):
                    fused, ins_hm = self.instance_fusion(bev, x[0].permute(0, 2, 3, 1), bs, **kwargs)
                x[0] = fused.permute(0, 3, 1, 2)
            with stage('a14 secondv2 stage%d' % (i + 1)):
                grid_features, _, this_feat = pts_backbone(x, 'stage{}'.format(i + 1))
            return_feats.append(this_feat)
        return return_feats, ins_hm
