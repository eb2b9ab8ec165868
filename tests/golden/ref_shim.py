"""Import the REFERENCE's own hot-path Python files (unmodified, from /root/reference) in this container.

TEST INFRASTRUCTURE ONLY — used by tests/golden/make_golden_ref.py to generate golden fixtures from the
reference itself.  The reference cannot be imported as a package here: `mmdet3d/__init__.py` needs mmcv / mmdet /
mmseg, `ops/sst/sst_ops.py` needs ipdb and the compiled TorchEx op, `fusion_encoder.py` loads `mmcv._ext`, the
sparse layers need spconv.  This module installs SMALL STAND-INS for exactly those third-party
pieces in `sys.modules` and then executes the reference files where they lie under their real module names, so
every line of the reference's own code on the path (ISFusionEncoder, SSTInputLayerV2, SSTv2, SECONDV2, DynamicVFE,
SparseEncoder / SparseBasicBlock / make_sparse_convmodule, SECONDFPN, NaiveSyncBatchNorm) runs as written.

What is a stand-in (third-party code that is absent, restated from its published behaviour):
  * mmcv Registry / build_from_cfg, build_conv_layer, build_norm_layer (name = abbreviation + postfix, eps
    default 1e-5, mmcv/cnn/bricks/norm.py), build_upsample_layer, ConvModule (conv [bias = no norm] -> norm ->
    ReLU, sub-module names conv / bn / activate), BaseModule, auto_fp16 / force_fp32 (identity: fp16 is off);
  * mmdet `BasicBlock` (mmdet 2.14 resnet.py: conv1/bn1/relu/conv2/bn2, norm1/norm2 properties);
  * mmcv.ops sparse-conv layer (`SparseConvTensor`, `SubMConv3d`, `SparseConv3d`, `SparseSequential`): a DENSE
    formulation — scatter to a grid, `F.conv3d`, active-site mask by a ones-kernel convolution — i.e. an
    implementation independent of the oracle's rulebook formulation; weight layout (kD,kH,kW,Cin,Cout) and
    `.dense()` as mmdet3d/ops/bevfusion-ops/spconv/{conv.py:100, structure.py:49-59};
  * `mmcv._ext.ms_deform_attn_forward`: the reference's OWN PyTorch statement of the kernel,
    ops/functions/ms_deform_attn_func.py:41-61, executed from the reference file (the reference's ops/test.py
    pins the CUDA op against exactly this function);
  * TorchEx `ingroup_indices.forward`: rank of each element inside its group in index order — one of the valid
    executions of the atomic counter kernel (ingroup_inds_kernel.cu:17-30);
  * `DynamicScatter`: the brute-force definition the reference's test uses (tests/test_dynamic_scatter.py:57-66),
    rows in `torch.unique(dim=0)` order as scatter_points_cuda.cu:204-212; the module wrapper itself
    (mmdet3d/ops/voxel/scatter_points.py:53-104) is the reference's file.
"""
import collections
import collections.abc
import importlib.util
import inspect
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF = '/root/reference'
_installed = False


# ----------------------------------------------------------------------------- mmcv stand-ins
class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name, self._module_dict, self.parent, self.children = name, {}, parent, []
        if parent is not None:
            parent.children.append(self)

    def get(self, key):
        r = self
        while r is not None:
            if key in r._module_dict:
                return r._module_dict[key]
            r = r.parent
        # search the whole tree (mmcv scopes are not modelled)
        root = self
        while root.parent is not None:
            root = root.parent
        stack = [root]
        while stack:
            r = stack.pop()
            if key in r._module_dict:
                return r._module_dict[key]
            stack.extend(r.children)
        return None

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._module_dict[name or module.__name__] = module
            return module
        if isinstance(name, type):          # used as a bare decorator
            self._module_dict[name.__name__] = name
            return name

        def _reg(cls):
            names = [name] if isinstance(name, str) else (name or [cls.__name__])
            for n in names:
                self._module_dict[n] = cls
            return cls
        return _reg

    def build(self, cfg, **kw):
        return build_from_cfg(cfg, self, kw.get('default_args'))


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop('type')
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError('%s is not in the %s registry' % (t, registry.name))
    return cls(**args)


class ConfigDict(dict):
    __getattr__ = dict.get


MMCV_MODELS = Registry('model')
CONV_LAYERS = Registry('conv layer')
NORM_LAYERS = Registry('norm layer')
UPSAMPLE_LAYERS = Registry('upsample layer')
ATTENTION = Registry('attention')
TRANSFORMER_LAYER = Registry('transformerLayer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')
for _n, _c in [('Conv1d', nn.Conv1d), ('Conv2d', nn.Conv2d), ('Conv3d', nn.Conv3d), ('Conv', nn.Conv2d)]:
    CONV_LAYERS.register_module(_n, module=_c)
for _n, _c in [('BN', nn.BatchNorm2d), ('BN1d', nn.BatchNorm1d), ('BN2d', nn.BatchNorm2d), ('BN3d', nn.BatchNorm3d),
               ('LN', nn.LayerNorm)]:
    NORM_LAYERS.register_module(_n, module=_c)
UPSAMPLE_LAYERS.register_module('deconv', module=nn.ConvTranspose2d)


def build_conv_layer(cfg, *args, **kwargs):
    cfg_ = dict(type='Conv2d') if cfg is None else dict(cfg)
    cls = CONV_LAYERS.get(cfg_.pop('type'))
    return cls(*args, **kwargs, **cfg_)


def _norm_abbr(cls):
    if hasattr(cls, '_abbr_'):
        return cls._abbr_
    if issubclass(cls, nn.modules.batchnorm._BatchNorm):
        return 'bn'
    if issubclass(cls, nn.LayerNorm):
        return 'ln'
    return 'norm_layer'


def build_norm_layer(cfg, num_features, postfix=''):
    cfg_ = dict(cfg)
    cls = NORM_LAYERS.get(cfg_.pop('type'))
    name = _norm_abbr(cls) + str(postfix)
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    layer = cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return name, layer


def build_upsample_layer(cfg, *args, **kwargs):
    cfg_ = dict(cfg)
    cls = UPSAMPLE_LAYERS.get(cfg_.pop('type'))
    return cls(*args, **kwargs, **cfg_)


class ConvModule(nn.Module):
    """mmcv/cnn/bricks/conv_module.py: conv (bias='auto' -> not with_norm) -> norm -> act(ReLU, inplace)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias='auto',
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True, order=('conv', 'norm', 'act')):
        super().__init__()
        self.order = order
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                     dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            assert act_cfg['type'] == 'ReLU'
            self.activate = nn.ReLU(inplace=inplace)

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.with_norm else None

    def forward(self, x, activate=True, norm=True):
        for layer in self.order:
            if layer == 'conv':
                x = self.conv(x)
            elif layer == 'norm' and norm and self.with_norm:
                x = self.norm(x)
            elif layer == 'act' and activate and self.with_activation:
                x = self.activate(x)
        return x


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


def _identity_decorator(*dargs, **dkw):
    """auto_fp16 / force_fp32: with fp16_enabled unset they call the function unchanged."""
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return dargs[0]
    return lambda f: f


def deprecated_api_warning(name_dict, cls_name=None):
    return lambda f: f


def to_2tuple(x):
    """mmcv.utils.to_2tuple = _ntuple(2): any iterable (strings included, e.g. padding='corner') passes through."""
    return x if isinstance(x, collections.abc.Iterable) else (x, x)


# ----------------------------------------------------------------------------- mmdet BasicBlock (2.14)
class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch', with_cp=False,
                 conv_cfg=None, norm_cfg=dict(type='BN'), dcn=None, plugins=None, init_cfg=None):
        super().__init__()
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, 3, stride=stride, padding=dilation, dilation=dilation,
                                      bias=False)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, 3, padding=1, bias=False)
        self.add_module(self.norm2_name, norm2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride, self.dilation, self.with_cp = stride, dilation, with_cp

    @property
    def norm1(self):
        return getattr(self, self.norm1_name)

    @property
    def norm2(self):
        return getattr(self, self.norm2_name)


class Bottleneck(nn.Module):      # imported by sparse_block.py, never built on this path
    expansion = 4


# ----------------------------------------------------------------------------- sparse conv, dense formulation
def _triple(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features, self.indices = features, indices
        self.spatial_shape, self.batch_size = list(spatial_shape), batch_size
        self.indice_dict, self.grid = {}, grid

    @property
    def spatial_size(self):
        n = 1
        for s in self.spatial_shape:
            n *= s
        return n

    def dense(self, channels_first=True):
        c = self.features.shape[1]
        out = torch.zeros([self.batch_size] + self.spatial_shape + [c], dtype=self.features.dtype)
        i = self.indices.long()
        out[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] = self.features
        return out.permute(0, 4, 1, 2, 3).contiguous() if channels_first else out


class SparseModule(nn.Module):
    pass


class _SparseConv(SparseModule):
    SUBM = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__()
        assert groups == 1
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _triple(kernel_size), _triple(stride)
        self.padding, self.dilation = _triple(padding), _triple(dilation)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x):
        feats = x.features
        i = x.indices.long()
        grid = torch.zeros([x.batch_size, feats.shape[1]] + x.spatial_shape, dtype=torch.float64)
        grid[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]] = feats.double()
        w = self.weight.permute(4, 3, 0, 1, 2).double()
        if self.SUBM:
            # submanifold: stride 1, centre-aligned, outputs exactly at the input sites (conv.py:129-131,152-164)
            pad = tuple((k // 2) * d for k, d in zip(self.kernel_size, self.dilation))
            out = F.conv3d(grid, w, None, 1, pad, self.dilation)
            oi = x.indices
            shape = x.spatial_shape
        else:
            out = F.conv3d(grid, w, None, self.stride, self.padding, self.dilation)
            mask = torch.zeros([x.batch_size, 1] + x.spatial_shape, dtype=torch.float64)
            mask[i[:, 0], 0, i[:, 1], i[:, 2], i[:, 3]] = 1
            act = F.conv3d(mask, torch.ones((1, 1) + self.kernel_size, dtype=torch.float64), None, self.stride,
                           self.padding, self.dilation)[:, 0] > 0.5
            oi = torch.nonzero(act).int()          # ascending (b,z,y,x): the order of the sorted-unique GPU rulebook
            shape = list(out.shape[2:])
        o = oi.long()
        f = out[o[:, 0], :, o[:, 1], o[:, 2], o[:, 3]].to(feats.dtype)
        if self.bias is not None:
            f = f + self.bias
        y = SparseConvTensor(f, oi, shape, x.batch_size)
        y.indice_dict = x.indice_dict
        return y


class SubMConv3d(_SparseConv):
    SUBM = True


class SparseConv3d(_SparseConv):
    pass


class SparseSequential(SparseModule):
    """mmcv/ops/sparse_modules.py: plain modules act on `.features`, sparse modules on the tensor."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], collections.OrderedDict):
            for k, m in args[0].items():
                self.add_module(k, m)
        else:
            for idx, m in enumerate(args):
                self.add_module(str(idx), m)
        for k, m in kwargs.items():
            self.add_module(k, m)

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, idx):
        return list(self._modules.values())[idx]

    def add(self, module, name=None):
        self.add_module(name or str(len(self._modules)), module)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x.features = m(x.features)
            else:
                x = m(x)
        return x


CONV_LAYERS.register_module('SubMConv3d', module=SubMConv3d)
CONV_LAYERS.register_module('SparseConv3d', module=SparseConv3d)


# ----------------------------------------------------------------------------- compiled ops
def _ingroup_forward(group_inds, out_inds):
    order = torch.argsort(group_inds, stable=True)
    g = group_inds[order]
    start = torch.ones_like(g, dtype=torch.bool)
    start[1:] = g[1:] != g[:-1]
    first = torch.nonzero(start).squeeze(1)
    seg = torch.cumsum(start.long(), 0) - 1
    rank = torch.arange(len(g)) - first[seg]
    out_inds[order] = rank


def _msda_core():
    src = open(os.path.join(REF, 'ops/functions/ms_deform_attn_func.py')).read().split('\n')
    ns = {'torch': torch, 'F': F}
    exec('\n'.join(src[40:61]), ns)
    return ns['ms_deform_attn_core_pytorch']


class _Ext:
    def __init__(self):
        self._core = _msda_core()

    def ms_deform_attn_forward(self, value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                               im2col_step=64):
        return self._core(value, spatial_shapes, sampling_locations, attention_weights)

    def ms_deform_attn_backward(self, *a, **k):
        raise NotImplementedError


def _dynamic_point_to_voxel_forward(feats, coors, reduce_type):
    """scatter_points_cuda.cu:183-239: drop rows with a negative coordinate, unique_dim(0) order, reduce."""
    ok = (coors >= 0).all(1)
    out_coors, inv, cnt = torch.unique(coors[ok], dim=0, return_inverse=True, return_counts=True)
    coors_map = torch.full((coors.shape[0],), -1, dtype=torch.int32)
    coors_map[ok] = inv.int()
    v, c = out_coors.shape[0], feats.shape[1]
    f = feats[ok]
    if reduce_type == 'max':
        red = torch.full((v, c), float('-inf'), dtype=feats.dtype)
        red = red.scatter_reduce(0, inv[:, None].expand(-1, c), f, 'amax', include_self=True)
    else:
        red = torch.zeros((v, c), dtype=feats.dtype).index_add_(0, inv, f)
        if reduce_type == 'mean':
            red = red / cnt[:, None].to(feats.dtype)
    return red, out_coors, coors_map, cnt.int()


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition('.')
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


def _pkg(name, **attrs):
    m = _mod(name, **attrs)
    m.__path__ = []
    return m


def _load_ref(name, rel):
    path = os.path.join(REF, rel)
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition('.')
    spec.loader.exec_module(m)
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


def install():
    """Install the stand-ins and load the reference files.  Returns a namespace of the reference classes."""
    global _installed
    if not os.path.isdir(REF):
        raise RuntimeError('the reference tree is only present in the build container')
    if not _installed:
        for n in ('mmcv', 'mmdet', 'mmseg', 'mmdet3d', 'spconv', 'ipdb'):
            assert n not in sys.modules, n + ' is importable: use the real reference instead of the shim'
        _mod('ipdb', set_trace=lambda *a, **k: None)
        _mod('ingroup_indices', forward=_ingroup_forward)
        ext = _Ext()
        _pkg('mmcv', __version__='1.4.0')
        _pkg('mmcv.cnn', ConvModule=ConvModule, build_conv_layer=build_conv_layer, build_norm_layer=build_norm_layer,
             build_upsample_layer=build_upsample_layer, MODELS=MMCV_MODELS, NORM_LAYERS=NORM_LAYERS,
             CONV_LAYERS=CONV_LAYERS, UPSAMPLE_LAYERS=UPSAMPLE_LAYERS)
        _pkg('mmcv.cnn.bricks')
        _mod('mmcv.cnn.bricks.registry', ATTENTION=ATTENTION, TRANSFORMER_LAYER=TRANSFORMER_LAYER,
             TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE, NORM_LAYERS=NORM_LAYERS, CONV_LAYERS=CONV_LAYERS)
        _pkg('mmcv.runner', force_fp32=_identity_decorator, auto_fp16=_identity_decorator, BaseModule=BaseModule,
             ModuleList=ModuleList, Sequential=Sequential)
        _mod('mmcv.runner.base_module', BaseModule=BaseModule, ModuleList=ModuleList, Sequential=Sequential)
        _pkg('mmcv.utils', Registry=Registry, build_from_cfg=build_from_cfg, ConfigDict=ConfigDict,
             deprecated_api_warning=deprecated_api_warning, to_2tuple=to_2tuple,
             ext_loader=types.SimpleNamespace(load_ext=lambda name, funcs: ext))
        _mod('mmcv.utils.ext_loader', load_ext=lambda name, funcs: ext)
        _pkg('mmcv.ops', SparseConvTensor=SparseConvTensor, SparseSequential=SparseSequential, SparseModule=SparseModule,
             SubMConv3d=SubMConv3d, SparseConv3d=SparseConv3d)
        _mod('mmcv.ops.multi_scale_deform_attn', multi_scale_deformable_attn_pytorch=ext._core)
        regs = {k: Registry(k.lower(), parent=MMCV_MODELS) for k in
                ('BACKBONES', 'DETECTORS', 'HEADS', 'LOSSES', 'NECKS', 'ROI_EXTRACTORS', 'SHARED_HEADS')}
        _pkg('mmdet')
        _pkg('mmdet.models', **regs)
        _mod('mmdet.models.builder', **regs)
        _pkg('mmdet.models.backbones')
        _mod('mmdet.models.backbones.resnet', BasicBlock=BasicBlock, Bottleneck=Bottleneck)
        _pkg('mmseg')
        _pkg('mmseg.models')
        _mod('mmseg.models.builder', SEGMENTORS=Registry('segmentors', parent=MMCV_MODELS))
        _mod('cv2') if 'cv2' not in sys.modules and importlib.util.find_spec('cv2') is None else None

        _pkg('mmdet3d')
        _pkg('mmdet3d.core', draw_heatmap_gaussian=None, gaussian_radius=None)
        ops = _pkg('mmdet3d.ops')
        _mod('mmdet3d.ops.spconv', IS_SPCONV2_AVAILABLE=False)
        _load_ref('mmdet3d.ops.norm', 'mmdet3d/ops/norm.py')
        sb = _load_ref('mmdet3d.ops.sparse_block', 'mmdet3d/ops/sparse_block.py')
        ops.SparseBasicBlock, ops.make_sparse_convmodule = sb.SparseBasicBlock, sb.make_sparse_convmodule
        _pkg('mmdet3d.ops.sst')
        sst = _load_ref('mmdet3d.ops.sst.sst_ops', 'mmdet3d/ops/sst/sst_ops.py')
        for k in ('flat2window_v2', 'window2flat_v2', 'make_continuous_inds', 'get_flat2win_inds_v2', 'get_window_coors',
                  'get_inner_win_inds', 'flat2window', 'window2flat'):
            setattr(ops, k, getattr(sst, k))
        # voxel ops: the Python wrappers are the reference's, the compiled layer is the stand-in above
        _pkg('mmdet3d.ops.voxel')
        _mod('mmdet3d.ops.voxel.voxel_layer', dynamic_point_to_voxel_forward=_dynamic_point_to_voxel_forward,
             dynamic_point_to_voxel_backward=None, dynamic_voxelize=None, hard_voxelize=None)
        sp = _load_ref('mmdet3d.ops.voxel.scatter_points', 'mmdet3d/ops/voxel/scatter_points.py')
        ops.DynamicScatter, ops.dynamic_scatter = sp.DynamicScatter, sp.dynamic_scatter

        _pkg('mmdet3d.models')
        _load_ref('mmdet3d.models.builder', 'mmdet3d/models/builder.py')
        _pkg('mmdet3d.models.sst')
        _load_ref('mmdet3d.models.sst.sst_basic_block_v2', 'mmdet3d/models/sst/sst_basic_block_v2.py')
        _load_ref('mmdet3d.models.sst.sst_input_layer_v2', 'mmdet3d/models/sst/sst_input_layer_v2.py')
        _pkg('mmdet3d.models.backbones')
        _load_ref('mmdet3d.models.backbones.sst_v2', 'mmdet3d/models/backbones/sst_v2.py')
        _load_ref('mmdet3d.models.backbones.second', 'mmdet3d/models/backbones/second.py')
        _pkg('mmdet3d.models.necks')
        _load_ref('mmdet3d.models.necks.second_fpn', 'mmdet3d/models/necks/second_fpn.py')
        _load_ref('mmdet3d.models.necks.generalized_lss', 'mmdet3d/models/necks/generalized_lss.py')
        _pkg('mmdet3d.models.middle_encoders')
        _load_ref('mmdet3d.models.middle_encoders.multi_scale_deformable_attn_function',
                  'mmdet3d/models/middle_encoders/multi_scale_deformable_attn_function.py')
        _load_ref('mmdet3d.models.middle_encoders.sparse_encoder', 'mmdet3d/models/middle_encoders/sparse_encoder.py')
        _load_ref('mmdet3d.models.middle_encoders.fusion_encoder', 'mmdet3d/models/middle_encoders/fusion_encoder.py')
        _pkg('mmdet3d.models.voxel_encoders')
        _load_ref('mmdet3d.models.voxel_encoders.utils', 'mmdet3d/models/voxel_encoders/utils.py')
        _load_ref('mmdet3d.models.voxel_encoders.voxel_encoder', 'mmdet3d/models/voxel_encoders/voxel_encoder.py')
        _installed = True
    m = sys.modules
    return types.SimpleNamespace(
        ISFusionEncoder=m['mmdet3d.models.middle_encoders.fusion_encoder'].ISFusionEncoder,
        fusion_encoder=m['mmdet3d.models.middle_encoders.fusion_encoder'],
        SSTInputLayerV2=m['mmdet3d.models.sst.sst_input_layer_v2'].SSTInputLayerV2,
        SSTv2=m['mmdet3d.models.backbones.sst_v2'].SSTv2,
        SECONDV2=m['mmdet3d.models.backbones.second'].SECONDV2,
        SECONDFPN=m['mmdet3d.models.necks.second_fpn'].SECONDFPN,
        GeneralizedLSSFPN=m['mmdet3d.models.necks.generalized_lss'].GeneralizedLSSFPN,
        SparseEncoder=m['mmdet3d.models.middle_encoders.sparse_encoder'].SparseEncoder,
        DynamicVFE=m['mmdet3d.models.voxel_encoders.voxel_encoder'].DynamicVFE,
        DynamicScatter=m['mmdet3d.ops'].DynamicScatter,
        sst_ops=m['mmdet3d.ops.sst.sst_ops'],
        SparseConvTensor=SparseConvTensor)


_swin_installed = False


class _MmcvFFN(BaseModule):
    """mmcv/cnn/bricks/transformer.py FFN (mmcv 1.4): [Linear -> act -> Dropout] x (num_fcs-1) -> Linear -> Dropout,
    + identity; parameter names layers.{i}.0.* and layers.{num_fcs-1}.*."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        act = {'GELU': nn.GELU, 'ReLU': nn.ReLU}[act_cfg['type']]
        layers, cin = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(cin, feedforward_channels), act(), nn.Dropout(ffn_drop)))
            cin = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.dropout_layer = nn.Identity()            # DropPath is the identity in eval mode
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        return (x if identity is None else identity) + self.dropout_layer(out)


def install_swin():
    """additionally load the reference's SwinTransformer (backbones/swin.py) with its PatchEmbed /
    PatchMerging (models/utils/transformer.py).  Stand-ins: mmcv FFN (above), DropPath = identity (eval), initialisers."""
    global _swin_installed
    ns = install()
    m = sys.modules
    if not _swin_installed:
        class _Inert(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
        cnn = m['mmcv.cnn']
        cnn.constant_init = lambda *a, **k: None
        cnn.trunc_normal_init = lambda *a, **k: None
        cnn.xavier_init = lambda *a, **k: None
        cnn.build_activation_layer = lambda cfg: {'GELU': nn.GELU, 'ReLU': nn.ReLU}[cfg['type']]()
        _mod('mmcv.cnn.bricks.transformer', FFN=_MmcvFFN, build_dropout=lambda cfg: nn.Identity(),
             BaseTransformerLayer=_Inert, TransformerLayerSequence=_Inert, build_transformer_layer_sequence=lambda *a, **k: None,
             MultiScaleDeformableAttention=_Inert)
        m['mmcv.ops.multi_scale_deform_attn'].MultiScaleDeformableAttention = _Inert
        m['mmcv.runner']._load_checkpoint = lambda *a, **k: None
        _pkg('mmdet.models.utils')
        _mod('mmdet.models.utils.builder', TRANSFORMER=Registry('Transformer'))
        _mod('mmdet3d.utils', get_root_logger=lambda *a, **k: None)
        utils = sys.modules.get('mmdet3d.models.utils')
        if utils is None or not hasattr(utils, '__path__'):
            old = utils
            utils = _pkg('mmdet3d.models.utils')
            if old is not None:
                utils.__dict__.update({k: v for k, v in old.__dict__.items() if not k.startswith('__')})
        _load_ref('mmdet3d.models.utils.ckpt_convert', 'mmdet3d/models/utils/ckpt_convert.py')
        _load_ref('mmdet3d.models.utils.transformer', 'mmdet3d/models/utils/transformer.py')
        _load_ref('mmdet3d.models.backbones.swin', 'mmdet3d/models/backbones/swin.py')
        _swin_installed = True
    ns.SwinTransformer = m['mmdet3d.models.backbones.swin'].SwinTransformer
    return ns


_head_installed = False


def install_head():
    """additionally load the reference's TransFusionHeadV2 (dense_heads/transfusion_head_v2.py) and
    TransFusionBBoxCoder (core/bbox/coders/transfusion_bbox_coder.py).  Only the inference forward / decode run; the
    training-side imports (losses, assigners, samplers, box structures, NMS ops) are inert placeholders."""
    global _head_installed
    ns = install()
    m = sys.modules
    if not _head_installed:
        class _Inert:
            def __init__(self, *a, **k):
                pass
        losses = m['mmdet.models.builder'].LOSSES
        for n in ('FocalLoss', 'L1Loss', 'GaussianFocalLoss', 'VarifocalLoss', 'CrossEntropyLoss'):
            losses.register_module(n, module=type(n, (_Inert,), {}))
        m['mmcv.cnn'].kaiming_init = lambda *a, **k: None
        core = m['mmdet3d.core']
        for n in ('circle_nms', 'xywhr2xyxyr', 'limit_period'):
            setattr(core, n, None)
        for n in ('PseudoSampler', 'Box3DMode', 'LiDARInstance3DBoxes'):
            setattr(core, n, type(n, (_Inert,), {}))
        _pkg('mmdet3d.core.bbox')
        _mod('mmdet3d.core.bbox.structures', rotation_3d_in_axis=None)
        _pkg('mmdet3d.core.bbox.coders')
        coders = Registry('bbox_coder')

        class BaseBBoxCoder:
            def __init__(self, **kwargs):
                pass

        def multi_apply(func, *args, **kwargs):                       # mmdet/core/utils/misc.py
            from functools import partial
            pfunc = partial(func, **kwargs) if kwargs else func
            return tuple(map(list, zip(*map(pfunc, *args))))
        _pkg('mmdet.core', build_bbox_coder=lambda cfg, **k: coders.build(cfg), multi_apply=multi_apply,
             build_assigner=lambda *a, **k: None, build_sampler=lambda *a, **k: None, AssignResult=_Inert)
        _pkg('mmdet.core.bbox', BaseBBoxCoder=BaseBBoxCoder)
        _mod('mmdet.core.bbox.builder', BBOX_CODERS=coders)
        _load_ref('mmdet3d.core.bbox.coders.transfusion_bbox_coder', 'mmdet3d/core/bbox/coders/transfusion_bbox_coder.py')
        _mod('mmdet3d.models.utils', clip_sigmoid=None)
        _mod('mmdet3d.models.fusion_layers', apply_3d_transformation=None)
        _pkg('mmdet3d.ops.iou3d')
        _mod('mmdet3d.ops.iou3d.iou3d_utils', nms_gpu=None)
        _pkg('mmdet3d.models.dense_heads')
        _load_ref('mmdet3d.models.dense_heads.transfusion_head_v2', 'mmdet3d/models/dense_heads/transfusion_head_v2.py')
        _head_installed = True
    ns.TransFusionHeadV2 = m['mmdet3d.models.dense_heads.transfusion_head_v2'].TransFusionHeadV2
    ns.TransFusionBBoxCoder = m['mmdet3d.core.bbox.coders.transfusion_bbox_coder'].TransFusionBBoxCoder
    return ns


if __name__ == '__main__':
    ns = install_head()
    print('reference classes loaded:', [k for k in vars(ns)])
    print(inspect.getsourcefile(ns.ISFusionEncoder))


_loading_installed = False


def install_loading():
    """SURVEY §8f rank 3 (input contract): load the reference's point loaders (datasets/pipelines/loading.py) with its own
    point classes (core/points/*.py).  Stand-ins: mmcv.FileClient = plain disk reads, the mmdet PIPELINES registry, and the
    two mmdet loader base classes the file imports but this path never builds."""
    global _loading_installed
    ns = install()
    m = sys.modules
    if not _loading_installed:
        class _FileClient:
            def __init__(self, backend='disk', **kw):
                assert backend == 'disk'

            def get(self, path):
                with open(path, 'rb') as f:
                    return f.read()

        mm = m['mmcv']
        mm.FileClient = _FileClient
        mm.check_file_exist = lambda p: None
        _pkg('mmdet.datasets')
        _mod('mmdet.datasets.builder', PIPELINES=Registry('pipeline'))
        _mod('mmdet.datasets.pipelines', LoadAnnotations=type('LoadAnnotations', (), {}),
             LoadImageFromFile=type('LoadImageFromFile', (), {}))
        core = m.get('mmdet3d.core') or _pkg('mmdet3d.core')
        if not hasattr(core, '__path__'):
            core.__path__ = []
        pts = _pkg('mmdet3d.core.points')
        for leaf in ('base_points', 'cam_points', 'depth_points', 'lidar_points'):
            _load_ref('mmdet3d.core.points.' + leaf, 'mmdet3d/core/points/%s.py' % leaf)
        # what core/points/__init__.py exports (it is 30 lines of imports + this lookup)
        pts.BasePoints = m['mmdet3d.core.points.base_points'].BasePoints
        pts.LiDARPoints = m['mmdet3d.core.points.lidar_points'].LiDARPoints
        pts.CameraPoints = m['mmdet3d.core.points.cam_points'].CameraPoints
        pts.DepthPoints = m['mmdet3d.core.points.depth_points'].DepthPoints
        pts.get_points_type = lambda t: {'CAMERA': pts.CameraPoints, 'LIDAR': pts.LiDARPoints, 'DEPTH': pts.DepthPoints}[t.upper()]
        _pkg('mmdet3d.datasets')
        _pkg('mmdet3d.datasets.pipelines')
        _load_ref('mmdet3d.datasets.pipelines.loading', 'mmdet3d/datasets/pipelines/loading.py')
        _loading_installed = True
    mod = m['mmdet3d.datasets.pipelines.loading']
    ns.LoadPointsFromFile = mod.LoadPointsFromFile
    ns.LoadPointsFromMultiSweeps = mod.LoadPointsFromMultiSweeps
    return ns
