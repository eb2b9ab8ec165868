"""SparseEncoder / SparseBasicBlock / make_sparse_convmodule on the B200 sparse-conv kernels.

ref: mmdet3d/models/middle_encoders/sparse_encoder.py:18-216, mmdet3d/ops/sparse_block.py:82-199.
Same constructor arguments, sub-module names and state_dict keys; in eval mode every conv + BN1d (+ residual)
(+ ReLU) is one fused implicit-GEMM kernel, SubM rulebooks are shared by all convs of a stage, and dense()
+ view(N, C*D, H, W) is written directly.
"""
import torch
from torch import nn

from . import _lib as L
from .registry import MIDDLE_ENCODERS, build_conv_layer, build_norm_layer
from .spconv import SparseConvTensor, SparseModule, SparseSequential, fold_bn


class SparseBasicBlock(SparseModule):
    """ref: sparse_block.py:82-134 (mmdet BasicBlock over SubMConv3d; names conv1/bn1/conv2/bn2)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, conv_cfg=None, norm_cfg=None):
        super().__init__()
        assert downsample is None and stride == 1
        conv_cfg = conv_cfg or dict(type='SubMConv3d')
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, 3, stride=stride, padding=1, dilation=1, bias=False)
        self.bn1 = build_norm_layer(norm_cfg, planes)[1]
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, 3, padding=1, bias=False)
        self.bn2 = build_norm_layer(norm_cfg, planes)[1]
        self.relu = nn.ReLU(inplace=True)

    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    def forward(self, x):
        if self.training:      # ref sparse_block.py:117-134: conv1 -> bn1 -> relu -> conv2 -> bn2 -> (+ identity) -> relu, batch statistics
            identity = x.features
            out = self.conv1(x)
            out = out.replace_feature(torch.relu(self.bn1(out.features)))
            out = self.conv2(out)
            return out.replace_feature(torch.relu(self.bn2(out.features) + identity))
        s1, b1 = fold_bn(self.bn1)
        s2, b2 = fold_bn(self.bn2)
        identity = x.features
        out = self.conv1(x, scale=s1, shift=b1, relu=True)
        out = self.conv2(out, scale=s2, shift=b2, residual=identity, relu=True)
        return out


def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key, stride=1, padding=0,
                           conv_type='SubMConv3d', norm_cfg=None, order=('conv', 'norm', 'act')):
    """ref: sparse_block.py:137-199."""
    assert isinstance(order, tuple) and len(order) <= 3
    conv_cfg = dict(type=conv_type, indice_key=indice_key)
    layers = []
    for layer in order:
        if layer == 'conv':
            layers.append(build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
                                           padding=padding, bias=False))
        elif layer == 'norm':
            layers.append(build_norm_layer(norm_cfg, out_channels)[1])
        elif layer == 'act':
            layers.append(nn.ReLU(inplace=True))
    return SparseSequential(*layers)


@MIDDLE_ENCODERS.register_module()
class SparseEncoder(nn.Module):

    def __init__(self, in_channels, sparse_shape, order=('conv', 'norm', 'act'),
                 norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), base_channels=16, output_channels=128,
                 encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                 encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)), block_type='conv_module', **kwargs):
        super().__init__()
        assert block_type in ['conv_module', 'basicblock']
        assert isinstance(order, tuple) and len(order) == 3 and set(order) == {'conv', 'norm', 'act'}
        self.sparse_shape = sparse_shape
        self.in_channels = in_channels
        self.order = order
        self.base_channels = base_channels
        self.output_channels = output_channels
        self.encoder_channels = encoder_channels
        self.encoder_paddings = encoder_paddings
        self.stage_num = len(self.encoder_channels)
        if self.order[0] != 'conv':
            self.conv_input = make_sparse_convmodule(in_channels, base_channels, 3, norm_cfg=norm_cfg, padding=1,
                                                     indice_key='subm1', conv_type='SubMConv3d', order=('conv',))
        else:
            self.conv_input = make_sparse_convmodule(in_channels, base_channels, 3, norm_cfg=norm_cfg, padding=1,
                                                     indice_key='subm1', conv_type='SubMConv3d')
        enc_out = self.make_encoder_layers(make_sparse_convmodule, norm_cfg, base_channels, block_type=block_type)
        self.conv_out = make_sparse_convmodule(enc_out, output_channels, kernel_size=(3, 1, 1), stride=(2, 1, 1),
                                               norm_cfg=norm_cfg, padding=0, indice_key='spconv_down2',
                                               conv_type='SparseConv3d')

    def make_encoder_layers(self, make_block, norm_cfg, in_channels, block_type='conv_module',
                            conv_cfg=dict(type='SubMConv3d')):
        self.encoder_layers = SparseSequential()
        for i, blocks in enumerate(self.encoder_channels):
            blocks_list = []
            for j, out_channels in enumerate(tuple(blocks)):
                padding = tuple(self.encoder_paddings[i])[j]
                if i != 0 and j == 0 and block_type == 'conv_module':
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg, stride=2,
                                                  padding=padding, indice_key=f'spconv{i + 1}', conv_type='SparseConv3d'))
                elif block_type == 'basicblock':
                    if j == len(blocks) - 1 and i != len(self.encoder_channels) - 1:
                        blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg, stride=2,
                                                      padding=padding, indice_key=f'spconv{i + 1}',
                                                      conv_type='SparseConv3d'))
                    else:
                        blocks_list.append(SparseBasicBlock(out_channels, out_channels, norm_cfg=norm_cfg, conv_cfg=conv_cfg))
                else:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg, padding=padding,
                                                  indice_key=f'subm{i + 1}', conv_type='SubMConv3d'))
                in_channels = out_channels
            self.encoder_layers.add_module(f'encoder_layer{i + 1}', SparseSequential(*blocks_list))
        return out_channels

    def strided_convs(self):
        """the four stride-2 SparseConv3d + conv_out, in execution order (their output-site capacity is tunable)."""
        out = []
        for stage in self.encoder_layers:
            for blk in stage:
                if isinstance(blk, SparseSequential) and not getattr(blk[0], 'subm', False):      # conv_module layouts wrap SubM convs too
                    out.append(blk[0])
        out.append(self.conv_out[0])
        return out

    def set_capacities(self, rows):
        """capacity mode: rows[i] = output-site buffer size of the i-th strided conv (None = analytic bound)."""
        for conv, r in zip(self.strided_convs(), rows):
            conv.out_capacity = r

    @torch.no_grad()
    def start_rulebooks(self, coors, batch_size, num_voxels_dev, stream, sorted_coors=True):
        """kick off every rulebook on `stream` as soon as the voxel coordinates exist (before the features do);
        returns the feature-less tensor to pass to forward(prebuilt=...)."""
        x = SparseConvTensor(None, coors.contiguous(), self.sparse_shape, batch_size, num_dev=num_voxels_dev)
        x.sorted = bool(sorted_coors)
        x.__dict__['_rb_stream'] = stream
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            self.prebuild_rulebooks(x)
        return x

    @torch.no_grad()
    def prebuild_rulebooks(self, x):
        """Build every rulebook of the encoder from the coordinates alone (they do not depend on features) and
        return [(event, ...)] in execution order.  Called on a side stream so that the index work of all stages
        overlaps the feature GEMMs; each rulebook is stored in the shared indice_dict under the key its conv will
        look up, with a CUDA event the consumer waits on."""
        events = []

        def index_only(indices, shape, num_dev):
            t = SparseConvTensor(None, indices, shape, x.batch_size, num_dev=num_dev)
            t.indice_dict = x.indice_dict
            t.sorted = x.sorted
            return t

        def build(conv, t):
            rb = conv.rulebook(t)
            ev = torch.cuda.Event()
            ev.record()
            x.indice_dict[('event', id(rb))] = ev
            events.append(ev)
            return rb

        cur = index_only(x.indices, x.spatial_shape, x.num_dev)
        build(self.conv_input[0], cur)
        for stage in self.encoder_layers:
            for blk in stage:
                if isinstance(blk, SparseBasicBlock):
                    build(blk.conv1, cur)
                    build(blk.conv2, cur)
                else:
                    rb = build(blk[0], cur)
                    cur = index_only(rb.out_indices, rb.out_shape, rb.num_dev)
                    cur.sorted = True
        build(self.conv_out[0], cur)
        return events

    def forward(self, voxel_features, coors, batch_size, *args, **kwargs):
        """eval: the fused no-grad CUDA path; train: the same kernels under autograd (spconv.ImplicitGemmFunction), BatchNorm with
        batch statistics over the active voxels, exact-size tensors (ref sparse_encoder.py:107-138)."""
        if self.training and torch.is_grad_enabled():
            return self.forward_train(voxel_features, coors, batch_size, sorted_coors=kwargs.get('sorted_coors', False))
        return self.forward_eval(voxel_features, coors, batch_size, *args, **kwargs)

    def forward_train(self, voxel_features, coors, batch_size, sorted_coors=False):
        x = SparseConvTensor(voxel_features.contiguous(), coors.int().contiguous(), self.sparse_shape, batch_size)
        x.sorted = bool(sorted_coors)
        x = self.conv_input(x)
        encode_features = [x]
        for encoder_layer in self.encoder_layers:
            x = encoder_layer(x)
            encode_features.append(x)
        out = self.conv_out(encode_features[-1])
        spatial_features = out.dense()
        n, c, d, h, w = spatial_features.shape
        return spatial_features.reshape(n, c * d, h, w), encode_features, {}

    @torch.no_grad()
    def forward_eval(self, voxel_features, coors, batch_size, swin_format=False, img_feats=None, dense=True,
                     num_voxels_dev=None, sorted_coors=False, rulebook_stream=None, prebuilt=None, **kwargs):
        """voxel_features [V,C], coors [V,4] int32 (b,z,y,x) -> (spatial_features [B, C*D, H, W], encode_features, kwargs)"""
        coors = coors.int()
        feats = voxel_features if voxel_features.dtype == torch.float16 else voxel_features.half()
        if prebuilt is not None:           # tensor whose rulebooks were started early (start_rulebooks); add the features
            x = prebuilt
            x.features = feats.contiguous()
            rulebook_stream_join = x.__dict__.get('_rb_stream')
        else:
            rulebook_stream_join = None
            x = SparseConvTensor(feats.contiguous(), coors.contiguous(), self.sparse_shape, batch_size, num_dev=num_voxels_dev)
            x.sorted = bool(sorted_coors)  # DynamicVFE output is in sorted-unique (b,z,y,x) order -> bitmap rulebooks
        if prebuilt is None and rulebook_stream is not None:
            # all index work on a side stream, overlapping the feature GEMMs of the main stream
            main = torch.cuda.current_stream()
            rulebook_stream.wait_stream(main)
            with torch.cuda.stream(rulebook_stream):
                self.prebuild_rulebooks(x)
        x = self.conv_input(x)
        encode_features = [x]
        for encoder_layer in self.encoder_layers:
            x = encoder_layer(x)
            encode_features.append(x)
        out = self.conv_out(encode_features[-1])
        self.last_out = out
        self.last_encode_features = encode_features
        if rulebook_stream_join is not None:
            torch.cuda.current_stream().wait_stream(rulebook_stream_join)
        elif rulebook_stream is not None:
            torch.cuda.current_stream().wait_stream(rulebook_stream)
        if not dense:      # fast path: the fusion encoder scatters `out` straight into its channels-last buffer
            return None, encode_features, kwargs
        spatial_features = out.dense()
        n, c, d, h, w = spatial_features.shape
        spatial_features = spatial_features.view(n, c * d, h, w)
        return spatial_features, encode_features, kwargs
