"""SwinTransformer (the image backbone) with its linear layers on the tcgen05 GEMM.

ref: mmdet3d/models/backbones/swin.py:20-763, mmdet3d/models/utils/transformer.py:62-385 (AdaptivePadding, PatchEmbed,
PatchMerging); configs/isfusion/isfusion_0075voxel.py:33-49 (Swin-T: 96 channels, depths 2/2/6/2, heads 3/6/12/24,
7x7 windows).  Same registry name, constructor kwargs, forward signature and state_dict keys as the reference.

Where the ~218 GFLOP of a 6-camera frame go: qkv / proj / FFN / patch-merging linears are ~90 % of them and run on
`isf_gemm_f16` (bias, GELU and the FFN's residual fused in the epilogue; the 4x4 patch embedding is a GEMM over the
unfolded 48-value patches).  LayerNorm, the window partition / cyclic shift glue and the 49-token attention core
(`F.scaled_dot_product_attention` with relative-position bias + shift mask as an additive mask, cached per resolution)
are torch ops by default; `ISF_SWIN_ATTN=own` replaces the whole pad / roll / partition / attention / reverse sequence by
`isf_swin_window_attention` (csrc/fusion.cu: the mma.sync tile code of the validated SST `window_attention_kernel`, with the
reference's padding, cyclic shift, region mask and crop as closed forms of (window, slot) — checked on the host against the
reference's index sequence), which also lets the `proj` GEMM add the residual in its epilogue.
Eval mode only (DropPath / dropout are the identity); activations are fp16 token matrices [B, L, C]."""
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from .fusion_encoder import _Packable
from .gemm import gemm, pack_weight
from .registry import BACKBONES


# opt-in: the library's shifted-window attention kernel (csrc/fusion.cu swin_window_attention_kernel) instead of the torch
# pad / roll / partition / SDPA / reverse sequence; with it a block is LN -> GEMM -> kernel -> GEMM(+residual) -> LN -> 2 GEMMs
OWN_WINDOW_ATTENTION = os.environ.get('ISF_SWIN_ATTN', 'own') == 'own'     # default since round 2 (GPU-verified, image branch 12.1 -> 5.7 ms); 'torch' = SDPA baseline


class _Lin:
    """packed fp16 weight + fp32 bias of one nn.Linear for the GEMM."""

    def __init__(self, lin):
        self.w = pack_weight(lin.weight.detach())
        self.b = lin.bias.detach().float().contiguous() if lin.bias is not None else None
        self.k = lin.in_features

    def __call__(self, x2d, act=None, residual=None):
        return gemm(x2d, self.w, k=self.k, shift=self.b, act=act, residual=residual)


class PatchEmbed(nn.Module):
    def __init__(self, in_channels, embed_dims, patch, norm=True):
        super().__init__()
        self.patch = patch
        self.projection = nn.Conv2d(in_channels, embed_dims, patch, stride=patch)
        self.norm = nn.LayerNorm(embed_dims) if norm else None


class PatchMerging(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.norm = nn.LayerNorm(4 * in_channels)
        self.reduction = nn.Linear(4 * in_channels, out_channels, bias=False)


class WindowMSA(nn.Module):
    def __init__(self, embed_dims, num_heads, window):
        super().__init__()
        self.embed_dims, self.num_heads, self.window = embed_dims, num_heads, window
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window - 1) ** 2, num_heads))
        ys, xs = torch.meshgrid(torch.arange(window), torch.arange(window), indexing='ij')
        p = torch.stack([ys.reshape(-1), xs.reshape(-1)], 1)
        rel = p[:, None, :] - p[None, :, :] + (window - 1)
        self.register_buffer('relative_position_index', rel[..., 0] * (2 * window - 1) + rel[..., 1])
        self.qkv = nn.Linear(embed_dims, 3 * embed_dims)
        self.proj = nn.Linear(embed_dims, embed_dims)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class ShiftWindowMSA(nn.Module):
    def __init__(self, embed_dims, num_heads, window, shift):
        super().__init__()
        self.window, self.shift = window, shift
        self.w_msa = WindowMSA(embed_dims, num_heads, window)


class _FFN(nn.Module):
    def __init__(self, embed_dims, hidden):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, hidden), nn.GELU(), nn.Dropout(0.0)),
                                    nn.Linear(hidden, embed_dims), nn.Dropout(0.0))


class SwinBlock(nn.Module):
    def __init__(self, embed_dims, num_heads, hidden, window, shift):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dims)
        self.attn = ShiftWindowMSA(embed_dims, num_heads, window, window // 2 if shift else 0)
        self.norm2 = nn.LayerNorm(embed_dims)
        self.ffn = _FFN(embed_dims, hidden)


class _Stage(nn.Module):
    def __init__(self, embed_dims, num_heads, hidden, depth, window, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([SwinBlock(embed_dims, num_heads, hidden, window, shift=(i % 2 == 1)) for i in range(depth)])
        self.downsample = downsample


def _shift_mask(hp, wp, ws, shift, device):
    """additive -100 mask between the regions a cyclic shift brings into one window (swin.py:228-247): [nW, N, N]."""
    region = torch.zeros((hp, wp), device=device)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            region[hs, wsl] = cnt
            cnt += 1
    mw = region.view(hp // ws, ws, wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


@BACKBONES.register_module()
class SwinTransformer(_Packable):
    act_dtype = torch.float16        # activation storage; the CPU logic test of this module runs it in fp32

    def __init__(self, pretrain_img_size=224, in_channels=3, embed_dims=96, patch_size=4, window_size=7, mlp_ratio=4,
                 depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), strides=(4, 2, 2, 2), out_indices=(0, 1, 2, 3), qkv_bias=True,
                 qk_scale=None, patch_norm=True, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, use_abs_pos_embed=False,
                 act_cfg=dict(type='GELU'), norm_cfg=dict(type='LN'), with_cp=False, pretrained=None, convert_weights=False,
                 frozen_stages=-1, init_cfg=None):
        super().__init__()
        if (use_abs_pos_embed or not qkv_bias or qk_scale is not None or act_cfg != dict(type='GELU') or norm_cfg != dict(type='LN')
                or tuple(strides) != (patch_size, 2, 2, 2)):
            raise L.IsfError('SwinTransformer: built for the isfusion_0075voxel.py configuration (config:33-49)')
        if embed_dims % 32 or any((embed_dims << i) // h % 8 for i, h in enumerate(num_heads)):
            raise L.IsfError('SwinTransformer: embed_dims must be a multiple of 32 and every head dimension a multiple of 8')
        self.out_indices = list(out_indices)
        self.window_size = window_size
        self.patch_embed = PatchEmbed(in_channels, embed_dims, patch_size, patch_norm)
        self.stages = nn.ModuleList()
        self.num_features = []
        c = embed_dims
        for i, depth in enumerate(depths):
            down = PatchMerging(c, 2 * c) if i < len(depths) - 1 else None
            self.stages.append(_Stage(c, num_heads[i], mlp_ratio * c, depth, window_size, down))
            self.num_features.append(c)
            if down is not None:
                c *= 2
        for i in self.out_indices:
            self.add_module('norm%d' % i, nn.LayerNorm(self.num_features[i]))
        self._pk = None
        self._masks = {}

    # ------------------------------------------------------------------ packing
    def _pack(self):
        if self.__dict__.get('_pk') is None:
            pe = self.patch_embed
            pk = dict(embed_w=pack_weight(pe.projection.weight.detach().reshape(pe.projection.out_channels, -1)),
                      embed_b=pe.projection.bias.detach().float().contiguous(), stages=[])
            for st in self.stages:
                blocks = []
                for blk in st.blocks:
                    m = blk.attn.w_msa
                    n = m.window * m.window
                    bias = m.relative_position_bias_table.detach()[m.relative_position_index.view(-1)].view(n, n, -1)
                    blocks.append(dict(qkv=_Lin(m.qkv), proj=_Lin(m.proj), fc1=_Lin(blk.ffn.layers[0][0]), fc2=_Lin(blk.ffn.layers[1]),
                                       bias=bias.permute(2, 0, 1).contiguous().float()))           # [heads, N, N]
                pk['stages'].append(dict(blocks=blocks, red=_Lin(st.downsample.reduction) if st.downsample is not None else None))
            self.__dict__['_pk'] = pk
            self.__dict__['_masks'] = {}
        return self.__dict__['_pk']

    def _attn_mask(self, key, bias, hp, wp, shift, dtype):
        """relative-position bias (+ shift mask) as ONE additive mask [1, nW, heads, N, N], cached per (block, resolution)."""
        cache = self.__dict__['_masks']
        ent = cache.get((key, hp, wp))
        if ent is None:
            ws = self.window_size
            m = bias[None, None]                                                     # [1, 1, heads, N, N]
            if shift > 0:
                m = m + _shift_mask(hp, wp, ws, shift, bias.device)[None, :, None]   # [1, nW, heads, N, N]
            else:
                m = m.expand(1, (hp // ws) * (wp // ws), -1, -1, -1)
            ent = m.to(dtype).contiguous()
            cache[(key, hp, wp)] = ent
        return ent

    # ------------------------------------------------------------------ forward
    def _block(self, x, hw, blk, pk, key):
        """x [B, L, C] fp16 -> [B, L, C] fp16 (swin.py:355-376)."""
        b, l, c = x.shape
        h, w = hw
        ws, shift, heads = self.window_size, blk.attn.shift, blk.attn.w_msa.num_heads
        if OWN_WINDOW_ATTENTION and ws == 7 and c == heads * 32 and heads % 3 == 0 and x.dtype == torch.float16:
            y = F.layer_norm(x, (c,), blk.norm1.weight.to(x.dtype), blk.norm1.bias.to(x.dtype), blk.norm1.eps).view(b * l, c)
            qkv = pk['qkv'](y.contiguous())                                            # token-major [B*L, 3C]
            att = torch.empty((b * l, c), dtype=torch.float16, device=x.device)
            L.check(L.lib().isf_swin_window_attention(L.ptr(qkv), L.ptr(pk['qkv'].b), L.ptr(pk['bias']), b, h, w, c, heads, shift,
                                                      L.ptr(att), L.stream()))
            x = pk['proj'](att, residual=x.reshape(b * l, c).contiguous()).view(b, l, c)   # + identity in the epilogue
            y = F.layer_norm(x, (c,), blk.norm2.weight.to(x.dtype), blk.norm2.bias.to(x.dtype), blk.norm2.eps).view(b * l, c)
            hid = pk['fc1'](y, act='gelu')
            return pk['fc2'](hid, residual=x.view(b * l, c).contiguous()).view(b, l, c)
        y = F.layer_norm(x, (c,), blk.norm1.weight.to(x.dtype), blk.norm1.bias.to(x.dtype), blk.norm1.eps).view(b, h, w, c)
        y = F.pad(y, (0, 0, 0, (-w) % ws, 0, (-h) % ws))
        hp, wp = y.shape[1], y.shape[2]
        if shift > 0:
            y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
        nwh, nww = hp // ws, wp // ws
        win = y.view(b, nwh, ws, nww, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, c)            # [B*nW*N, C] window-major rows
        qkv = pk['qkv'](win.contiguous()).view(b, nwh * nww, ws * ws, 3, heads, c // heads)
        q, k, v = (qkv[:, :, :, i].permute(0, 1, 3, 2, 4) for i in range(3))                      # [B, nW, heads, N, d]
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=self._attn_mask(key, pk['bias'], hp, wp, shift, q.dtype))
        o = o.permute(0, 1, 3, 2, 4).reshape(-1, c)                                               # [B*nW*N, C]
        o = pk['proj'](o.contiguous()).view(b, nwh, nww, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, hp, wp, c)
        if shift > 0:
            o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
        x = x + o[:, :h, :w, :].reshape(b, l, c)
        y = F.layer_norm(x, (c,), blk.norm2.weight.to(x.dtype), blk.norm2.bias.to(x.dtype), blk.norm2.eps).view(b * l, c)
        hid = pk['fc1'](y, act='gelu')
        return pk['fc2'](hid, residual=x.view(b * l, c).contiguous()).view(b, l, c)              # + identity in the epilogue

    @torch.no_grad()
    def forward(self, x):
        """x [B*views, 3, H, W] -> list of [B*views, C_i, H_i, W_i] fp16 maps (channels-last memory) for `out_indices`."""
        if self.training:
            raise L.IsfError('SwinTransformer: eval-mode CUDA path only')
        if not L.is_device(x):
            raise L.IsfError('expected a CUDA tensor (the product path has no CPU fallback)')
        pk = self._pack()
        pe = self.patch_embed
        p = pe.patch
        hh, ww = x.shape[-2:]
        x = F.pad(x.to(self.act_dtype), [0, (-ww) % p, 0, (-hh) % p])                                         # AdaptivePadding 'corner'
        b = x.shape[0]
        hw = (x.shape[2] // p, x.shape[3] // p)
        rows = F.unfold(x, kernel_size=p, stride=p).transpose(1, 2).reshape(b * hw[0] * hw[1], -1)   # (c, kh, kw)-major = conv weight order
        if rows.shape[1] % 8:
            raise L.IsfError('SwinTransformer: in_channels * patch_size^2 must be a multiple of 8')
        t = gemm(rows.contiguous(), pk['embed_w'], k=rows.shape[1], shift=pk['embed_b'])
        c = t.shape[1]
        if pe.norm is not None:
            t = F.layer_norm(t, (c,), pe.norm.weight.to(t.dtype), pe.norm.bias.to(t.dtype), pe.norm.eps)
        t = t.view(b, hw[0] * hw[1], c)
        outs = []
        for i, (st, spk) in enumerate(zip(self.stages, pk['stages'])):
            for j, blk in enumerate(st.blocks):
                t = self._block(t, hw, blk, spk['blocks'][j], (i, j))
            if i in self.out_indices:
                nrm = getattr(self, 'norm%d' % i)
                o = F.layer_norm(t, (t.shape[-1],), nrm.weight.to(t.dtype), nrm.bias.to(t.dtype), nrm.eps)
                outs.append(o.view(b, hw[0], hw[1], -1).permute(0, 3, 1, 2))                      # NCHW view of channels-last memory
            if st.downsample is not None:
                c = t.shape[-1]
                m = t.view(b, hw[0], hw[1], c).permute(0, 3, 1, 2)
                m = F.pad(m, [0, hw[1] % 2, 0, hw[0] % 2])
                hw = (m.shape[2] // 2, m.shape[3] // 2)
                m = F.unfold(m, kernel_size=2, stride=2).transpose(1, 2)                          # [B, L/4, 4C] channel-major
                ds = st.downsample
                m = F.layer_norm(m, (4 * c,), ds.norm.weight.to(m.dtype), ds.norm.bias.to(m.dtype), ds.norm.eps)
                t = spk['red'](m.reshape(-1, 4 * c).contiguous()).view(b, hw[0] * hw[1], 2 * c)
        return outs
