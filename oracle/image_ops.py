"""Oracle: the image backbone SwinTransformer (config isfusion_0075voxel.py:33-49).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED by outputs of the reference's own
mmdet3d/models/backbones/swin.py + models/utils/transformer.py (PatchEmbed, PatchMerging) run in the build container
(tests/golden/ref_swin.npz, tests/test_oracle_vs_reference.py).  Eval-mode restatement with the reference's parameter
names (patch_embed.projection / norm, stages.i.blocks.j.{norm1, attn.w_msa.{relative_position_bias_table,
relative_position_index, qkv, proj}, norm2, ffn.layers.0.0, ffn.layers.1}, stages.i.downsample.{norm, reduction},
norm{i}).  DropPath / dropout are the identity in eval mode.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _pad_to_multiple(x, k_h, k_w):
    """AdaptivePadding 'corner' for kernel == stride (utils/transformer.py:62-131): zero-pad bottom / right."""
    h, w = x.shape[-2:]
    ph, pw = (-h) % k_h, (-w) % k_w
    return F.pad(x, [0, pw, 0, ph]) if ph or pw else x


class PatchEmbed(nn.Module):
    """ref: utils/transformer.py:134-257 (conv k = s = patch, corner padding, LayerNorm)."""

    def __init__(self, in_channels, embed_dims, patch, norm=True):
        super().__init__()
        self.patch = patch
        self.projection = nn.Conv2d(in_channels, embed_dims, patch, stride=patch)
        self.norm = nn.LayerNorm(embed_dims) if norm else None

    def forward(self, x):
        x = self.projection(_pad_to_multiple(x, self.patch, self.patch))
        hw = (x.shape[2], x.shape[3])
        x = x.flatten(2).transpose(1, 2)
        return (self.norm(x) if self.norm is not None else x), hw


class PatchMerging(nn.Module):
    """ref: utils/transformer.py:260-385: 2x2 unfold (channel-major: index c*4 + kh*2 + kw) -> LN(4C) -> Linear(4C, 2C)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.norm = nn.LayerNorm(4 * in_channels)
        self.reduction = nn.Linear(4 * in_channels, out_channels, bias=False)

    def forward(self, x, hw):
        b, _, c = x.shape
        h, w = hw
        x = _pad_to_multiple(x.view(b, h, w, c).permute(0, 3, 1, 2), 2, 2)
        h, w = x.shape[-2:]
        x = F.unfold(x, kernel_size=2, stride=2).transpose(1, 2)
        return self.reduction(self.norm(x)), (h // 2, w // 2)


class WindowMSA(nn.Module):
    """ref: swin.py:20-123."""

    def __init__(self, embed_dims, num_heads, window):
        super().__init__()
        self.num_heads, self.window = num_heads, window
        self.scale = (embed_dims // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window - 1) ** 2, num_heads))
        ys, xs = torch.meshgrid(torch.arange(window), torch.arange(window), indexing='ij')
        p = torch.stack([ys.reshape(-1), xs.reshape(-1)], 1)                 # [N, 2]
        rel = p[:, None, :] - p[None, :, :] + (window - 1)                   # [N, N, 2] in 0..2w-2
        self.register_buffer('relative_position_index', rel[..., 0] * (2 * window - 1) + rel[..., 1])
        self.qkv = nn.Linear(embed_dims, 3 * embed_dims)
        self.proj = nn.Linear(embed_dims, embed_dims)

    def forward(self, x, mask=None):
        bw, n, c = x.shape
        qkv = self.qkv(x).reshape(bw, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        attn = q @ k.transpose(-2, -1)
        bias = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1)
        attn = attn + bias.unsqueeze(0)
        if mask is not None:
            nw = mask.shape[0]
            attn = (attn.view(bw // nw, nw, self.num_heads, n, n) + mask[None, :, None]).view(-1, self.num_heads, n, n)
        return self.proj((attn.softmax(-1) @ v).transpose(1, 2).reshape(bw, n, c))


class ShiftWindowMSA(nn.Module):
    """ref: swin.py:126-283 (pad to a multiple of the window, cyclic shift by window // 2, -100 mask between regions)."""

    def __init__(self, embed_dims, num_heads, window, shift):
        super().__init__()
        self.window, self.shift = window, shift
        self.w_msa = WindowMSA(embed_dims, num_heads, window)

    def _windows(self, x):
        b, h, w, c = x.shape
        ws = self.window
        return x.view(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, c)

    def forward(self, x, hw):
        b, _, c = x.shape
        h, w = hw
        ws = self.window
        x = x.view(b, h, w, c)
        x = F.pad(x, (0, 0, 0, (-w) % ws, 0, (-h) % ws))
        hp, wp = x.shape[1], x.shape[2]
        mask = None
        if self.shift > 0:
            x = torch.roll(x, shifts=(-self.shift, -self.shift), dims=(1, 2))
            region = torch.zeros((1, hp, wp, 1))
            cnt = 0
            for hs in (slice(0, -ws), slice(-ws, -self.shift), slice(-self.shift, None)):
                for wsl in (slice(0, -ws), slice(-ws, -self.shift), slice(-self.shift, None)):
                    region[:, hs, wsl, :] = cnt
                    cnt += 1
            mw = self._windows(region).view(-1, ws * ws)
            diff = mw.unsqueeze(1) - mw.unsqueeze(2)
            mask = torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))
        y = self.w_msa(self._windows(x), mask)
        y = y.view(b, hp // ws, wp // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, hp, wp, c)
        if self.shift > 0:
            y = torch.roll(y, shifts=(self.shift, self.shift), dims=(1, 2))
        return y[:, :h, :w, :].reshape(b, h * w, c)


class _FFN(nn.Module):
    """mmcv FFN with num_fcs = 2: Linear -> GELU -> Linear (+ identity); names layers.0.0 / layers.1."""

    def __init__(self, embed_dims, hidden):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, hidden), nn.GELU(), nn.Dropout(0.0)),
                                    nn.Linear(hidden, embed_dims), nn.Dropout(0.0))

    def forward(self, x, identity):
        return identity + self.layers(x)


class SwinBlock(nn.Module):
    """ref: swin.py:286-376 (pre-norm)."""

    def __init__(self, embed_dims, num_heads, hidden, window, shift):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dims)
        self.attn = ShiftWindowMSA(embed_dims, num_heads, window, window // 2 if shift else 0)
        self.norm2 = nn.LayerNorm(embed_dims)
        self.ffn = _FFN(embed_dims, hidden)

    def forward(self, x, hw):
        x = x + self.attn(self.norm1(x), hw)
        return self.ffn(self.norm2(x), x)


class _Stage(nn.Module):
    def __init__(self, embed_dims, num_heads, hidden, depth, window, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([SwinBlock(embed_dims, num_heads, hidden, window, shift=(i % 2 == 1)) for i in range(depth)])
        self.downsample = downsample

    def forward(self, x, hw):
        for blk in self.blocks:
            x = blk(x, hw)
        if self.downsample is not None:
            xd, hwd = self.downsample(x, hw)
            return xd, hwd, x, hw
        return x, hw, x, hw


class SwinTransformer(nn.Module):
    """ref: swin.py:465-763 (forward :744-763): patch embed -> 4 stages -> LayerNorm'ed NCHW maps of `out_indices`."""

    def __init__(self, in_channels=3, embed_dims=96, patch_size=4, window_size=7, mlp_ratio=4, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), out_indices=(0, 1, 2, 3), patch_norm=True, **kwargs):
        super().__init__()
        self.out_indices = list(out_indices)
        self.patch_embed = PatchEmbed(in_channels, embed_dims, patch_size, patch_norm)
        self.stages = nn.ModuleList()
        c = embed_dims
        self.num_features = []
        for i, depth in enumerate(depths):
            down = PatchMerging(c, 2 * c) if i < len(depths) - 1 else None
            self.stages.append(_Stage(c, num_heads[i], mlp_ratio * c, depth, window_size, down))
            self.num_features.append(c)
            if down is not None:
                c *= 2
        for i in self.out_indices:
            self.add_module('norm%d' % i, nn.LayerNorm(self.num_features[i]))

    @torch.no_grad()
    def forward(self, x):
        x, hw = self.patch_embed(x)
        outs = []
        for i, stage in enumerate(self.stages):
            x, hw, out, ohw = stage(x, hw)
            if i in self.out_indices:
                out = getattr(self, 'norm%d' % i)(out)
                outs.append(out.view(-1, ohw[0], ohw[1], self.num_features[i]).permute(0, 3, 1, 2).contiguous())
        return outs
