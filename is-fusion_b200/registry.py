"""Tiny registry / builder shim so that configs/isfusion/isfusion_0075voxel.py can be instantiated by
name without mmcv (ref: mmdet3d/models/builder.py:10-63).  When mmcv/mmdet ARE importable the same
classes are also registered into their registries with force=True — the pattern the reference uses
for spconv2 (mmdet3d/ops/spconv/overwrite_spconv/write_spconv2.py:20-36) — so the reference's own
`build_model(cfg.model)` picks up the B200 modules."""
import inspect

from torch import nn

from .norm import NaiveSyncBatchNorm1d, NaiveSyncBatchNorm2d


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=True, module=None):
        def _reg(cls):
            self._modules[name or cls.__name__] = cls
            return cls
        return _reg(module) if module is not None else _reg

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg, **default_args):
        cfg = dict(cfg)
        typ = cfg.pop('type')
        cls = typ if inspect.isclass(typ) else self._modules.get(typ)
        if cls is None:
            raise KeyError(f'{typ} is not in the {self.name} registry')
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        return cls(**cfg)


MODELS = Registry('models')
# mmdet3d aliases all of these to one registry (builder.py:10-15)
BACKBONES = NECKS = VOXEL_ENCODERS = MIDDLE_ENCODERS = FUSION_LAYERS = DETECTORS = MODELS
NORM_LAYERS = Registry('norm')
CONV_LAYERS = Registry('conv')
PIPELINES = Registry('pipeline')        # mmdet.datasets.builder.PIPELINES (data-loading transforms of the input contract)

NORM_LAYERS.register_module('BN', module=nn.BatchNorm2d)
NORM_LAYERS.register_module('BN1d', module=nn.BatchNorm1d)
NORM_LAYERS.register_module('BN2d', module=nn.BatchNorm2d)
# plain BatchNorm in eval / single process, cross-rank moments when training on >1 rank (ops/norm.py:136-271)
NORM_LAYERS.register_module('naiveSyncBN1d', module=NaiveSyncBatchNorm1d)
NORM_LAYERS.register_module('naiveSyncBN2d', module=NaiveSyncBatchNorm2d)
CONV_LAYERS.register_module('Conv2d', module=nn.Conv2d)


def build_norm_layer(cfg, num_features):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    cfg.pop('requires_grad', None)
    layer = NORM_LAYERS.get(typ)(num_features, **cfg)
    return typ.lower(), layer


def build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg) if cfg is not None else dict(type='Conv2d')
    typ = cfg.pop('type')
    return CONV_LAYERS.get(typ)(*args, **kwargs, **cfg)


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_middle_encoder(cfg):
    return MIDDLE_ENCODERS.build(cfg)


def build_voxel_encoder(cfg):
    return VOXEL_ENCODERS.build(cfg)


def build_fusion_layer(cfg):
    return FUSION_LAYERS.build(cfg)


def build_detector(cfg, **kw):
    return DETECTORS.build(cfg, **kw)


class ConvModule(nn.Sequential):
    """mmcv.cnn.ConvModule subset used by the path (fusion_encoder.py:862-958): conv (bias=False when a
    norm follows) -> BN -> ReLU with sub-module names conv / bn / activate."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type='ReLU'), bias='auto'):
        super().__init__()
        with_norm = norm_cfg is not None
        if bias == 'auto':
            bias = not with_norm
        self.add_module('conv', nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                          bias=bias))
        if with_norm:
            self.add_module('bn', build_norm_layer(norm_cfg, out_channels)[1])
        if act_cfg is not None:
            self.add_module('activate', nn.ReLU(inplace=True))


def register_into_mmcv():
    """Best effort: expose the B200 modules under the reference's registry names."""
    try:
        from mmdet.models.builder import MODELS as MM
    except Exception:
        return False
    for k, v in MODELS._modules.items():
        MM.register_module(name=k, force=True, module=v)
    try:
        from mmcv.cnn import CONV_LAYERS as MC
        for k, v in CONV_LAYERS._modules.items():
            if k != 'Conv2d':
                MC.register_module(name=k, force=True, module=v)
    except Exception:
        pass
    return True
