"""GeneralizedLSSFPN (the image neck, SURVEY §8f rank 1: the producer of `img_feats`) on the B200 kernels.

ref: mmdet3d/models/necks/generalized_lss.py:12-103, configs/isfusion/isfusion_0075voxel.py:50-55.  Same registry name,
constructor kwargs, forward signature and state_dict keys.  Per level, top-down: bilinear upsample (align_corners=True)
of the level above -> concat -> 1x1 ConvModule -> 3x3 ConvModule.  On a channels-last map the concat + 1x1 conv is ONE
tcgen05 GEMM over the pixel rows ([N*H*W, C_lat + C_up] x [256, C]^T) with the folded BN scale / shift and the ReLU in
its epilogue; the 3x3 conv runs like the neck's dense convs (cuDNN, or the library's implicit GEMM with ISF_OWN_CONV=1);
the upsample is torch's bilinear kernel on the fp16 channels-last map.  Eval mode only.  The Swin backbone that feeds it
stays the reference's module."""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from .fusion_encoder import _conv, _fold_conv_bn, _Packable
from .gemm import gemm, pack_weight
from .registry import NECKS, ConvModule


@NECKS.register_module()
class GeneralizedLSSFPN(_Packable):

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, no_norm_on_lateral=False,
                 conv_cfg=None, norm_cfg=dict(type='BN2d'), act_cfg=dict(type='ReLU'),
                 upsample_cfg=dict(mode='bilinear', align_corners=True), init_cfg=None):
        super().__init__()
        if start_level != 0 or end_level != -1 or no_norm_on_lateral or act_cfg != dict(type='ReLU'):
            raise L.IsfError('GeneralizedLSSFPN: built for the isfusion_0075voxel.py configuration (config:50-55)')
        if out_channels not in (32, 64, 128, 256):
            raise L.IsfError('GeneralizedLSSFPN: out_channels must be one column tile of the GEMM (32/64/128/256)')
        self.in_channels, self.out_channels, self.num_outs = list(in_channels), out_channels, num_outs
        self.upsample_cfg = dict(upsample_cfg)
        last = len(in_channels) - 1
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(last):
            cin = in_channels[i] + (in_channels[i + 1] if i == last - 1 else out_channels)
            self.lateral_convs.append(ConvModule(cin, out_channels, 1, norm_cfg=norm_cfg))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1, norm_cfg=norm_cfg))
        self._pk = None

    def _pack(self):
        if self.__dict__.get('_pk') is None:
            pk = []
            for lat, fpn in zip(self.lateral_convs, self.fpn_convs):
                bn = lat.bn
                scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
                shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
                w = pack_weight(lat.conv.weight.detach()[:, :, 0, 0])                 # [Cout, Cin] -> K-major fp16
                pk.append(dict(w=w, scale=scale, shift=shift, k=lat.conv.in_channels, fpn=_fold_conv_bn(fpn.conv, fpn.bn)))
            self.__dict__['_pk'] = pk
        return self.__dict__['_pk']

    @torch.no_grad()
    def forward(self, inputs):
        """inputs: the backbone's stage outputs [N, C_i, H_i, W_i] (any float dtype / memory format) ->
        tuple of [N, out_channels, H_i, W_i] fp16 maps (channels-last memory) for every level but the last."""
        if self.training:
            raise L.IsfError('GeneralizedLSSFPN: eval-mode CUDA path only')
        if len(inputs) != len(self.in_channels):
            raise L.IsfError('GeneralizedLSSFPN: expected %d input levels' % len(self.in_channels))
        pk = self._pack()
        lat = []
        for t in inputs:
            if not L.is_device(t):
                raise L.IsfError('expected a CUDA tensor (the product path has no CPU fallback)')
            lat.append(t.half().contiguous(memory_format=torch.channels_last))
        for i in range(len(lat) - 2, -1, -1):
            n, _, h, w = lat[i].shape
            up = F.interpolate(lat[i + 1], size=(h, w), **self.upsample_cfg)
            rows = torch.cat([lat[i].permute(0, 2, 3, 1), up.permute(0, 2, 3, 1)], dim=-1).reshape(n * h * w, -1)
            p = pk[i]
            y = gemm(rows.contiguous(), p['w'], k=p['k'], scale=p['scale'], shift=p['shift'], act='relu')   # 1x1 conv + BN + ReLU
            y = _conv(y.view(n, h, w, self.out_channels), p['fpn'])                                       # 3x3 conv + BN + ReLU
            lat[i] = y.permute(0, 3, 1, 2)
        return tuple(lat[:-1])
