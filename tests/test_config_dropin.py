"""Drop-in check of the module/registry boundary: the reference's own config file instantiates the B200 modules by
name (no mmcv), and the parameter names of the hot-path sub-modules are the reference's.
Needs the reference checkout (build container only); skipped elsewhere."""
import os

import pytest

CFG = '/root/reference/configs/isfusion/isfusion_0075voxel.py'


@pytest.mark.skipif(not os.path.exists(CFG), reason='reference checkout not present')
def test_reference_config_builds_b200_modules():
    import isfusion_b200.detector  # noqa: F401  (registers the modules)
    from isfusion_b200.registry import build_detector
    ns = {}
    exec(compile(open(CFG).read(), CFG, 'exec'), ns)       # the config is self-contained (no _base_)
    model = dict(ns['model'])
    assert model['type'] == 'ISFusionDetector'
    det = build_detector(model)                            # every sub-module of the config is built by name
    sd = det.state_dict()
    expected = [
        'pts_voxel_encoder.vfe_layers.0.linear.weight', 'pts_voxel_encoder.vfe_layers.1.norm.running_var',
        'pts_middle_encoder.conv_input.0.weight', 'pts_middle_encoder.encoder_layers.encoder_layer1.0.conv1.weight',
        'pts_middle_encoder.encoder_layers.encoder_layer3.2.0.weight', 'pts_middle_encoder.conv_out.1.running_mean',
        'fusion_encoder.conv_fusion.conv.weight', 'fusion_encoder.conv_fusion.bn.running_mean',
        'fusion_encoder.grid2region_att.0.linear0.weight',
        'fusion_encoder.grid2region_att.1.block_list.0.encoder_list.1.win_attn.self_attn.in_proj_weight',
        'fusion_encoder.grid2region_att.0.block_list.0.encoder_list.0.norm2.bias',
        'fusion_encoder.heatmap_head_3.bias', 'fusion_encoder.instance_att.layers.1.cross_attn.sampling_offsets.bias',
        'fusion_encoder.instance_att.query_pos_embed.position_embedding_head.3.weight',
        'fusion_encoder.instance_to_scene_att.multihead_attn.in_proj_weight', 'fusion_encoder.instance_to_scene_att.norm.weight',
        'pts_backbone.blocks.0.15.weight', 'pts_backbone.ds_layer.1.running_var', 'pts_backbone.blocks.1.12.weight',
        'pts_neck.deblocks.0.0.weight', 'pts_neck.deblocks.1.1.running_mean',
        'img_backbone.patch_embed.projection.weight', 'img_backbone.stages.2.blocks.5.attn.w_msa.relative_position_bias_table',
        'img_backbone.stages.1.downsample.reduction.weight', 'img_backbone.stages.3.blocks.1.ffn.layers.1.bias', 'img_backbone.norm3.weight',
        'img_neck.lateral_convs.1.conv.weight', 'img_neck.fpn_convs.0.bn.running_mean',
        'pts_bbox_head.shared_conv.bias', 'pts_bbox_head.heatmap_head.0.bn.running_var', 'pts_bbox_head.class_encoding.weight',
        'pts_bbox_head.decoder.0.multihead_attn.in_proj_weight', 'pts_bbox_head.decoder.0.cross_posembed.position_embedding_head.0.weight',
        'pts_bbox_head.prediction_heads.0.vel.0.conv.weight', 'pts_bbox_head.prediction_heads.0.heatmap.1.bias',
    ]
    missing = [k for k in expected if k not in sd]
    assert not missing, missing
    # shapes fixed by the config
    assert tuple(sd['pts_middle_encoder.conv_input.0.weight'].shape) == (3, 3, 3, 64, 32)       # spconv1/mmcv layout
    assert tuple(sd['fusion_encoder.conv_fusion.conv.weight'].shape) == (128, 768, 3, 3)
    assert tuple(sd['pts_neck.deblocks.1.0.weight'].shape) == (256, 256, 2, 2)
    assert det.fusion_encoder.bev_size == 180 and det.fusion_encoder.instance_num == 200
    assert det.pts_pillar_layer.max_num_points == 12 and det.pillar_size[:2] == [0.6, 0.6]
    assert det.pts_middle_encoder.sparse_shape == [41, 1440, 1440]
    assert tuple(sd['img_neck.lateral_convs.1.conv.weight'].shape) == (256, 1152, 1, 1)
    assert tuple(sd['pts_bbox_head.shared_conv.weight'].shape) == (128, 512, 3, 3)
    assert det.pts_bbox_head.x_size == 180 and det.pts_bbox_head.num_proposals == 200 and det.pts_bbox_head.num_classes == 10
    assert det.pts_bbox_head.bbox_coder_cfg['post_center_range'] == [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]


def test_head_module_has_the_reference_state_dict_keys():
    """the B200 TransFusionHeadV2 builds from the reference's head config and exposes exactly the
    parameter names of the reference module (tests/golden/ref_head.npz holds the reference module's keys)."""
    import os
    import sys
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, 'golden'))
    import golden_common as gc
    from isfusion_b200.head import TransFusionHeadV2
    from isfusion_b200.registry import MODELS
    assert MODELS.get('TransFusionHeadV2') is TransFusionHeadV2
    cfg = gc.head_cfg()
    net = MODELS.build(dict(type='TransFusionHeadV2', **cfg))
    keys = list(np.load(os.path.join(here, 'golden', 'ref_head.npz'))['keys'])
    assert sorted(net.state_dict().keys()) == keys
    from oracle import head_ops
    ref = head_ops.TransFusionHeadV2(**cfg)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}


def test_img_neck_module_has_the_reference_state_dict_keys():
    import os
    import sys
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, 'golden'))
    import golden_common as gc
    from isfusion_b200.img_neck import GeneralizedLSSFPN
    net = GeneralizedLSSFPN(**gc.lss_cfg())
    assert sorted(net.state_dict().keys()) == list(np.load(os.path.join(here, 'golden', 'ref_lss.npz'))['keys'])


@pytest.mark.skipif(not os.path.exists(CFG), reason='reference checkout not present')
def test_synthetic_full_model_config_equals_the_reference_config():
    """synth.full_model_config() — what the GPU box builds the whole detector from (no reference checkout there) — holds
    the reference config's values for every key it carries (losses / train_cfg are the only keys left out)."""
    import synth
    ns = {}
    exec(compile(open(CFG).read(), CFG, 'exec'), ns)
    ref, mine = ns['model'], synth.full_model_config()

    def norm(v):
        if isinstance(v, dict):
            return {k: norm(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [norm(x) for x in v]
        return v

    def check(a, b, path):
        assert isinstance(b, dict), path
        for k, v in a.items():
            assert k in b, path + '.' + k
            if isinstance(v, dict):
                check(v, b[k], path + '.' + k)
            else:
                assert norm(v) == norm(b[k]), (path + '.' + k, v, b[k])

    check(mine, ref, 'model')
    left_out = set(ref) - set(mine)
    assert left_out == {'train_cfg'}, left_out
    assert set(ref['pts_bbox_head']) - set(mine['pts_bbox_head']) == {'loss_cls', 'loss_bbox', 'loss_heatmap'}


def test_load_checkpoint_helper(tmp_path, capsys):
    """a checkpoint saved the way mmcv does (`{'meta', 'state_dict'}` with DataParallel's `module.` prefix, sparse-conv weights
    in the spconv2 layout) loads into the B200 modules; key problems are reported, `strict` raises."""
    import pytest as _pytest
    import torch
    import isfusion_b200.spconv as sp
    from isfusion_b200.checkpoint import load_checkpoint
    src = torch.nn.Sequential(sp.SubMConv3d(16, 32, 3, padding=1, bias=False), torch.nn.BatchNorm1d(32))
    torch.nn.init.normal_(src[0].weight)
    sd = {'module.' + k: v.clone() for k, v in src.state_dict().items()}
    sd['module.0.weight'] = sd['module.0.weight'].permute(4, 0, 1, 2, 3).contiguous()          # spconv2 layout
    sd['module.extra.bias'] = torch.zeros(3)
    path = str(tmp_path / 'ckpt.pth')
    torch.save({'meta': {'epoch': 10}, 'state_dict': sd}, path)
    dst = torch.nn.Sequential(sp.SubMConv3d(16, 32, 3, padding=1, bias=False), torch.nn.BatchNorm1d(32))
    ckpt = load_checkpoint(dst, path)
    assert ckpt['meta']['epoch'] == 10 and torch.equal(dst[0].weight, src[0].weight) and torch.equal(dst[1].running_var, src[1].running_var)
    assert 'unexpected key' in capsys.readouterr().out
    with _pytest.raises(RuntimeError):
        load_checkpoint(dst, path, strict=True)
