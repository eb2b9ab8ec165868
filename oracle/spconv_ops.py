"""Oracle: sparse 3-D convolution + SparseEncoder (row a4 of SURVEY §8).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The layer the config runs lives in un-vendored spconv>=2.0 /
mmcv-full 1.3.8-1.4.0 (mmdet3d/ops/spconv/__init__.py:4-12) and the reference's tests assert shapes only
(tests/test_models/test_common_modules/test_sparse_unet.py:53-137).  PINNED against the vendored spconv-1.x CPU
generator compiled from the reference header where it lies (oracle/_ref/libref_spconv.so, fixture
tests/golden/spconv_ref_cpp.npz: output sites and (offset, input, output) pairs identical, gather-mm-scatter over the
reference's pairs equal) and cross-checked against a dense F.conv3d on the scattered grid (tests/test_oracle_spconv.py);
the SparseEncoder assembly is pinned by the reference's own sparse_encoder.py / sparse_block.py driving a dense-grid
conv (tests/golden/ref_sparse.npz).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _triple(v):
    return list(v) if isinstance(v, (list, tuple)) else [v, v, v]


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    """ref: mmdet3d/ops/bevfusion-ops/spconv/ops.py:20-31."""
    return [(input_size[i] + 2 * padding[i] - dilation[i] * (kernel_size[i] - 1) - 1) // stride[i] + 1
            for i in range(len(input_size))]


class SparseConvTensor:
    """ref: mmdet3d/ops/bevfusion-ops/spconv/structure.py:20-63."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices          # [V,4] int32 (b,z,y,x)
        self.spatial_shape = list(spatial_shape)
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid

    def replace_feature(self, f):
        t = SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size)
        t.indice_dict = self.indice_dict
        return t

    def dense(self, channels_first=True):
        """ref: structure.py:49-59 (scatter_nd of features at (b,z,y,x))."""
        c = self.features.shape[1]
        out = torch.zeros([self.batch_size] + self.spatial_shape + [c], dtype=self.features.dtype)
        idx = self.indices.long()
        out[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = self.features
        return out.permute(0, 4, 1, 2, 3).contiguous() if channels_first else out


def _lin(idx, shape):
    idx = idx.long()
    return ((idx[:, 0] * shape[0] + idx[:, 1]) * shape[1] + idx[:, 2]) * shape[2] + idx[:, 3]


def get_indice_pairs(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, subm):
    """Rulebook.  Returns (out_indices [Vo,4] int32, out_shape, nbr [K, Vo] int64) where nbr[k, o] is the input
    row that reaches output row o through kernel offset k (or -1), k = (kz*KH + ky)*KW + kx.

    ref (spec): bevfusion-ops/spconv/include/spconv/geometry.h:24-85 (getValidOutPos: for input position i and
    kernel offset k the output is o = (i + pad - k*dil)/stride when divisible and in range; offset index
    accumulates with the LAST dim fastest, :62-73), :144-194 (getIndicePairsConv), :247-297 (SubM: outputs are
    the inputs, centre aligned), spconv_ops.h:104,129-137 (SubM keeps indices; strided conv on GPU takes the
    SORTED unique linear output index).  Output ROW ORDER is implementation-defined in the reference (§3.5);
    canonical here = ascending linear index, the GPU order.
    """
    ksize, stride, padding, dilation = map(_triple, (ksize, stride, padding, dilation))
    indices = indices.long()
    v = indices.shape[0]
    kvol = ksize[0] * ksize[1] * ksize[2]
    if subm:
        out_shape = list(spatial_shape)
        out_indices = indices
    else:
        out_shape = get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    in_key = _lin(indices, spatial_shape)
    in_sorted, in_order = torch.sort(in_key)
    offs = [(kz, ky, kx) for kz in range(ksize[0]) for ky in range(ksize[1]) for kx in range(ksize[2])]
    if not subm:
        cand = []
        for (kz, ky, kx) in offs:
            k3 = torch.tensor([kz, ky, kx])
            num = indices[:, 1:] + torch.tensor(padding) - k3 * torch.tensor(dilation)
            ok = (num % torch.tensor(stride) == 0).all(1)
            o = num // torch.tensor(stride)
            ok &= ((o >= 0) & (o < torch.tensor(out_shape))).all(1)
            cand.append(torch.cat([indices[ok, :1], o[ok]], 1))
        cand = torch.cat(cand, 0)
        okey = torch.unique(_lin(cand, out_shape))            # sorted ascending
        s2, s1, s0 = out_shape[2], out_shape[1], out_shape[0]
        out_indices = torch.stack([okey // (s0 * s1 * s2), (okey // (s1 * s2)) % s0, (okey // s2) % s1, okey % s2], 1)
    vo = out_indices.shape[0]
    nbr = torch.full((kvol, vo), -1, dtype=torch.long)
    for k, (kz, ky, kx) in enumerate(offs):
        k3 = torch.tensor([kz, ky, kx])
        if subm:
            pad = torch.tensor([(ksize[d] // 2) * dilation[d] for d in range(3)])
            ipos = out_indices[:, 1:] - pad + k3 * torch.tensor(dilation)
        else:
            ipos = out_indices[:, 1:] * torch.tensor(stride) - torch.tensor(padding) + k3 * torch.tensor(dilation)
        ok = ((ipos >= 0) & (ipos < torch.tensor(list(spatial_shape)))).all(1)
        key = _lin(torch.cat([out_indices[:, :1], ipos], 1), spatial_shape)
        pos = torch.searchsorted(in_sorted, key).clamp(max=max(v - 1, 0))
        hit = ok & (in_sorted[pos] == key) if v > 0 else torch.zeros_like(ok)
        nbr[k, hit] = in_order[pos[hit]]
    return out_indices.int(), out_shape, nbr


def indice_conv(features, weight, nbr):
    """out[o] = sum_k features[nbr[k,o]] @ W[k].  ref: spconv_ops.h:260-361 (per-offset gather -> mm ->
    scatter-add; SubM centre offset is a plain mm :300-303).  weight [kD,kH,kW,Cin,Cout] (conv.py:100)."""
    kvol, vo = nbr.shape
    w = weight.reshape(kvol, weight.shape[-2], weight.shape[-1])
    out = torch.zeros((vo, w.shape[-1]), dtype=features.dtype)
    for k in range(kvol):
        o = torch.nonzero(nbr[k] >= 0).squeeze(1)
        if o.numel():
            out.index_add_(0, o, features[nbr[k, o]] @ w[k])
    return out


class SparseConvolution(nn.Module):
    """ref: bevfusion-ops/spconv/conv.py:55-223 (weight (kD,kH,kW,Cin,Cout); rulebook cached by indice_key)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 subm=False, indice_key=None):
        super().__init__()
        assert groups == 1
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _triple(kernel_size), _triple(stride)
        self.padding, self.dilation = _triple(padding), _triple(dilation)
        self.subm, self.indice_key = subm, indice_key
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x):
        cached = x.indice_dict.get(self.indice_key) if self.indice_key is not None else None
        if cached is None:
            cached = get_indice_pairs(x.indices, x.batch_size, x.spatial_shape, self.kernel_size, self.stride,
                                      self.padding, self.dilation, self.subm)
            if self.indice_key is not None:
                x.indice_dict[self.indice_key] = cached
        out_indices, out_shape, nbr = cached
        f = indice_conv(x.features, self.weight, nbr)
        if self.bias is not None:
            f = f + self.bias
        out = SparseConvTensor(f, out_indices, out_shape, x.batch_size)
        out.indice_dict = x.indice_dict
        return out


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True, indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, False, indice_key)


class SparseSequential(nn.Sequential):
    """ref: bevfusion-ops/spconv/modules.py (dense modules act on .features)."""

    def forward(self, x):
        for m in self:
            if isinstance(m, (SparseConvolution, SparseSequential, SparseBasicBlock)):
                x = m(x)
            else:
                x = x.replace_feature(m(x.features))
        return x


class SparseBasicBlock(nn.Module):
    """ref: mmdet3d/ops/sparse_block.py:82-134 + mmdet BasicBlock (conv1->norm1->relu->conv2->norm2->+id->relu;
    parameter names conv1/bn1/conv2/bn2; SubMConv3d k3 p1 without bias and without indice_key)."""

    def __init__(self, inplanes, planes, eps=1e-3, momentum=0.01):
        super().__init__()
        self.conv1 = SubMConv3d(inplanes, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm1d(planes, eps=eps, momentum=momentum)
        self.conv2 = SubMConv3d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm1d(planes, eps=eps, momentum=momentum)

    def forward(self, x):
        identity = x.features
        out = self.conv1(x)
        out = out.replace_feature(F.relu(self.bn1(out.features)))
        out = self.conv2(out)
        out = out.replace_feature(self.bn2(out.features))
        out = out.replace_feature(F.relu(out.features + identity))
        return out


def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key, stride=1, padding=0,
                           conv_type='SubMConv3d', eps=1e-3, momentum=0.01):
    """ref: mmdet3d/ops/sparse_block.py:137-199 (conv(bias=False) -> BN1d -> ReLU)."""
    cls = SubMConv3d if conv_type == 'SubMConv3d' else SparseConv3d
    return SparseSequential(cls(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False,
                                indice_key=indice_key),
                            nn.BatchNorm1d(out_channels, eps=eps, momentum=momentum), nn.ReLU(inplace=True))


class SparseEncoder(nn.Module):
    """ref: mmdet3d/models/middle_encoders/sparse_encoder.py:18-216 (block_type='basicblock')."""

    def __init__(self, in_channels, sparse_shape, order=('conv', 'norm', 'act'),
                 norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), base_channels=16, output_channels=128,
                 encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                 encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)), block_type='conv_module', **kwargs):
        super().__init__()
        assert block_type == 'basicblock' and order[0] == 'conv'
        self.sparse_shape = sparse_shape
        eps, mom = norm_cfg.get('eps', 1e-5), norm_cfg.get('momentum', 0.1)
        self.conv_input = make_sparse_convmodule(in_channels, base_channels, 3, 'subm1', padding=1,
                                                 conv_type='SubMConv3d', eps=eps, momentum=mom)
        self.encoder_layers = SparseSequential()
        cin = base_channels
        for i, blocks in enumerate(encoder_channels):
            bl = []
            for j, cout in enumerate(tuple(blocks)):
                padding = tuple(encoder_paddings[i])[j]
                if j == len(blocks) - 1 and i != len(encoder_channels) - 1:
                    bl.append(make_sparse_convmodule(cin, cout, 3, f'spconv{i + 1}', stride=2, padding=padding,
                                                     conv_type='SparseConv3d', eps=eps, momentum=mom))
                else:
                    bl.append(SparseBasicBlock(cout, cout, eps, mom))
                cin = cout
            self.encoder_layers.add_module(f'encoder_layer{i + 1}', SparseSequential(*bl))
        self.conv_out = make_sparse_convmodule(cin, output_channels, (3, 1, 1), 'spconv_down2', stride=(2, 1, 1),
                                               padding=0, conv_type='SparseConv3d', eps=eps, momentum=mom)

    @torch.no_grad()
    def forward(self, voxel_features, coors, batch_size, **kwargs):
        x = SparseConvTensor(voxel_features.float(), coors.int(), self.sparse_shape, batch_size)
        x = self.conv_input(x)
        encode_features = [x]
        for layer in self.encoder_layers:
            x = layer(x)
            encode_features.append(x)
        out = self.conv_out(encode_features[-1])
        sp = out.dense()
        n, c, d, h, w = sp.shape
        return sp.view(n, c * d, h, w), encode_features, kwargs


def dense_conv_reference(x, conv):
    """Independent check: dense F.conv3d on the scattered grid (SubM: output masked to the active inputs)."""
    dense = x.dense()                                                       # [B,C,D,H,W]
    w = conv.weight.permute(4, 3, 0, 1, 2).contiguous()                     # [Cout,Cin,kD,kH,kW]
    pad = [(k // 2) for k in conv.kernel_size] if conv.subm else conv.padding
    y = F.conv3d(dense, w, None, stride=conv.stride if not conv.subm else 1, padding=pad, dilation=conv.dilation)
    return y
