"""Dense 3x3 BEV convolutions on the library's own tcgen05 kernels (no cuDNN on the product path).

The neck's dense convolutions (conv_fusion 768->128, the heat-map / scene / instance convs, SECONDV2: 203 GFLOP per
frame, reference fusion_encoder.py:862-958, backbones/second.py:98-239) are nn.Conv2d modules in the reference.  Here:

* `BACKEND = 'tma'` (default): `isf_conv3x3_nhwc_f16` (csrc/conv2d.cu) — implicit GEMM with TMA-tiled halos: the input
  tile (+ halo) of a CTA comes in as 4-D / 5-D tensor boxes straight into the swizzled UMMA layout, zero padding is the
  TMA's out-of-bounds fill, the three dy taps re-use one box, folded-BN shift + ReLU in the epilogue, TMA store.
* `BACKEND = 'gather'` (ISF_CONV=gather; also the fallback for input widths that are not a multiple of 64): the sparse
  convolution kernel's special case "every site active, 9 kernel offsets" — the token matrix [B*H*W, C] is the feature
  matrix, the neighbour table is a closed form of (b, y, x) cached per (B, H, W, stride) (`isf_spconv_implicit_gemm`).
  Widths that are not a power of two (768 = 3 x 256) are split into channel groups = virtual offsets (9 x 3 = 27).
* `BACKEND = 'cudnn'` (ISF_CONV=cudnn): the library baseline kept for A/B measurements only (scripts/bench_conv.py).

Same folded fp16 weights and fp32 accumulation everywhere; the host-side index / packing logic is pure torch and is
checked on the CPU against F.conv2d (tests/test_conv2d_own.py)."""
import ctypes
import os

import torch

from . import _lib as L

BACKEND = os.environ.get('ISF_CONV', 'tma')          # 'tma' | 'gather' | 'cudnn' (A/B baseline only)
_plans = {}
_weights = {}


def channel_groups(cin):
    """smallest split of the input width into power-of-two groups the kernel accepts (>= 8 channels each)."""
    g = 1
    while cin % g == 0 and g <= 3:
        c = cin // g
        if c >= 8 and c & (c - 1) == 0:
            return g
        g += 1
    return None


def dense_rulebook(bs, h, w, stride, groups, device):
    """neighbour table of a 3x3, padding-1 convolution over a dense [bs, h, w] grid whose feature rows are
    (b*h + y)*w + x (times `groups`, plus the group): nbr [9*groups, rows_pad] int32 (-1 = outside), oh, ow."""
    oh, ow = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
    o = torch.arange(bs * oh * ow, device=device)
    b, oy, ox = o // (oh * ow), (o // ow) % oh, o % ow
    rows = []
    for ky in range(3):
        for kx in range(3):
            iy, ix = oy * stride - 1 + ky, ox * stride - 1 + kx
            ok = (iy >= 0) & (iy < h) & (ix >= 0) & (ix < w)
            r = (b * h + iy) * w + ix
            for g in range(groups):
                rows.append(torch.where(ok, r * groups + g, torch.full_like(r, -1)))
    n = o.numel()
    pad = max((n + 127) // 128 * 128, 128)
    nbr = torch.full((9 * groups, pad), -1, dtype=torch.int32, device=device)
    nbr[:, :n] = torch.stack(rows).int()
    return nbr.contiguous(), oh, ow


def pack_dense_weight(w, bias):
    """[Cout, Cin, 3, 3] (+ bias [Cout]) -> K-major fp16 [Cout_pad, Kpad] with column (ky*3 + kx)*Cin + c, fp32 shift
    [Cout_pad]; Cout_pad = next of {32, 64, 128, 256}."""
    cout, cin = w.shape[0], w.shape[1]
    cpad = next((c for c in (32, 64, 128, 256) if c >= cout), None)
    if cpad is None:
        return None
    k = 9 * cin
    kpad = (k + 63) // 64 * 64
    wp = torch.zeros((cpad, kpad), dtype=w.dtype if w.dtype == torch.bfloat16 else torch.float16, device=w.device)
    wp[:cout, :k] = w.permute(0, 2, 3, 1).reshape(cout, k).to(wp.dtype)
    shift = torch.zeros(cpad, dtype=torch.float32, device=w.device)
    if bias is not None:
        shift[:cout] = bias.float()
    return wp.contiguous(), shift.contiguous(), cpad


def _packed_weight(w, bias, shift_src):
    """packed weight + fp32 shift of a folded conv, cached per weight tensor.  The entry keeps the source tensors alive: a
    key built from data_ptr() alone would alias a NEW tensor the allocator places at a freed weight's address."""
    key = (w.data_ptr(), w._version, tuple(w.shape), bias.data_ptr() if bias is not None else 0, bias._version if bias is not None else 0)
    hit = _weights.get(key)
    if hit is None:
        ent = pack_dense_weight(w, shift_src)
        if ent is None:
            return None
        if len(_weights) > 256:
            _weights.clear()
        hit = (ent, w, bias, shift_src)
        _weights[key] = hit
    return hit[0]


def emulate(x_nhwc, nbr, wp, shift, groups, cout, relu):
    """pure-torch statement of what the kernel computes from (nbr, wp, shift) — used by the CPU test of this module."""
    bs, h, w, cin = x_nhwc.shape
    cg = cin // groups
    a = x_nhwc.reshape(bs * h * w * groups, cg).float()
    taps = nbr.shape[0]
    n = nbr.shape[1]
    acc = torch.zeros((n, wp.shape[0]), dtype=torch.float32)
    for t in range(taps):
        r = nbr[t].long()
        rows = torch.where((r >= 0)[:, None], a[r.clamp(min=0)], torch.zeros(1, cg))
        acc += rows @ wp[:, t * cg:(t + 1) * cg].float().t()
    acc = acc + shift
    if relu:
        acc = torch.relu(acc)
    return acc[:, :cout]


def conv3x3_gather(x_nhwc, packed, stride=1, relu=True):
    """3x3 / padding 1 conv (+ folded-BN bias) (+ReLU) of a channels-last fp16 map [B,H,W,Cin] -> [B,OH,OW,Cout] on the
    sparse-conv implicit-GEMM kernel, or None when the shape is outside what the kernel takes."""
    w, bias = packed[0], packed[1]
    cout, cin = w.shape[0], w.shape[1]
    if tuple(w.shape[2:]) != (3, 3) or stride not in (1, 2):
        return None
    groups = channel_groups(cin)
    if groups is None or 9 * groups > 27:
        return None
    ent = _packed_weight(w, bias, bias)
    if ent is None:
        return None
    wp, shift, cpad = ent
    bs, h, wd, _ = x_nhwc.shape
    pkey = (x_nhwc.device, bs, h, wd, stride, groups)
    plan = _plans.get(pkey)
    if plan is None:
        plan = dense_rulebook(bs, h, wd, stride, groups, x_nhwc.device)
        _plans[pkey] = plan
    nbr, oh, ow = plan
    x = x_nhwc if (x_nhwc.is_contiguous() and x_nhwc.dtype == torch.float16) else x_nhwc.half().contiguous()
    L.require_cuda(x, wp, shift, nbr)
    n_out = bs * oh * ow
    out = torch.empty((n_out, cpad), dtype=torch.float16, device=x.device)
    L.check(L.lib().isf_spconv_implicit_gemm(
        L.ptr(x), cin // groups, L.ptr(wp), wp.shape[1], cpad, L.ptr(nbr), ctypes.c_longlong(nbr.shape[1]), 9 * groups, n_out,
        L.ptr(None), L.ptr(None), L.ptr(shift), L.ptr(None), 0, 1 if relu else 0, L.ptr(out), cpad, L.stream()))
    return out.view(bs, oh, ow, cpad)[..., :cout]


def conv3x3_tma(x_nhwc, packed, stride=1, relu=True):
    """the same conv on `isf_conv3x3_nhwc_f16` (TMA-tiled halos), or None when the shape is outside what that kernel takes
    (input width not a multiple of 64, more than 256 output channels, odd map with stride 2)."""
    w, bias = packed[0], packed[1]
    cout, cin = w.shape[0], w.shape[1]
    bs, h, wd, cx = x_nhwc.shape
    if tuple(w.shape[2:]) != (3, 3) or stride not in (1, 2) or cin % 64 or cx != cin or cout > 256:
        return None
    if stride == 2 and (h % 2 or wd % 2):
        return None
    ent = _packed_weight(w, bias, packed[2] if len(packed) > 2 else bias)      # fp32 folded shift when the caller kept it
    if ent is None:
        return None
    wp, shift, cpad = ent
    x = x_nhwc if (x_nhwc.is_contiguous() and x_nhwc.dtype == wp.dtype) else x_nhwc.to(wp.dtype).contiguous()
    L.require_cuda(x, wp, shift)
    oh, ow = (h - 1) // stride + 1, (wd - 1) // stride + 1
    ldo = (cout + 7) // 8 * 8
    out = torch.empty((bs, oh, ow, ldo), dtype=wp.dtype, device=x.device)
    L.check(L.fn16('isf_conv3x3_nhwc_f16', wp.dtype)(L.ptr(x), bs, h, wd, cin, cin, L.ptr(wp), cpad, L.ptr(shift), stride, 1 if relu else 0,
                                         L.ptr(out), ldo, ldo, L.stream()))
    return out if ldo == cout else out[..., :cout]


def conv3x3(x_nhwc, packed, stride=1, relu=True):
    """3x3 / padding 1 conv (+ folded-BN shift) (+ReLU) of a channels-last fp16 map on the library's own kernels."""
    y = None
    if BACKEND != 'gather':
        y = conv3x3_tma(x_nhwc, packed, stride, relu)
    if y is None:
        y = conv3x3_gather(x_nhwc, packed, stride, relu)
    if y is None:
        raise L.IsfError('conv3x3: no kernel for input %s / weight %s / stride %d' % (tuple(x_nhwc.shape), tuple(packed[0].shape), stride))
    return y
