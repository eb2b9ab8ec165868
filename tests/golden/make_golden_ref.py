"""Golden fixtures produced BY THE REFERENCE'S OWN PYTHON FILES, executed unmodified from /root/reference.

Build container only:  python tests/golden/make_golden_ref.py      (writes tests/golden/ref_*.npz)

tests/golden/ref_shim.py explains what is the reference's code (every module on the path) and what is a stand-in
(absent third-party pieces: mmcv builders, a dense-formulation sparse conv, the PyTorch statement of the MSDA
kernel from the reference's ops/functions/ms_deform_attn_func.py, the TorchEx in-group rank).  Weights and the
large inputs are regenerated on both sides by tests/golden/golden_common.py, so the fixtures hold outputs only.

Fixtures (checked by tests/test_oracle_vs_reference.py against oracle/):
  ref_fusion.npz  : ISFusionEncoder.forward + SECONDV2 stages (fusion_encoder.py:1152-1189, second.py:200-230) on
                    a 24x24 BEV, B = 2: both output maps, the instance heat-map, img_fv_to_bev (every 8th channel),
                    conv_fusion output (every 8th channel); SECONDFPN on the two maps (every 8th channel).
  ref_fusion_full.npz : the same at the config's own sizes (180x180 / 90x90, B = 1): sampled values + moments.
  ref_sst.npz     : SSTInputLayerV2 (shuffle off) window/slot indices, positional embedding, key mask for both
                    shifts, and SSTv2 level-0 output on those tokens.
  ref_vfe.npz     : DynamicVFE.forward (voxel_encoder.py:452-547) features + coordinates; DynamicScatter mean.
  ref_sparse.npz  : SparseEncoder.forward (sparse_encoder.py:107-138) dense output [B,512,H/8,W/8] and the
                    active-site count after every stage.
  ref_swin.npz    : SwinTransformer.forward (backbones/swin.py:744-763, Swin-T of the config) on two 62x94 images.
  ref_lss.npz     : GeneralizedLSSFPN.forward (necks/generalized_lss.py:79-103) on Swin-T-shaped stage outputs of 3 images.
  ref_head.npz    : TransFusionHeadV2.forward (transfusion_head_v2.py:771-908) prediction dict on a [2,512,24,24] map,
                    query labels, and TransFusionBBoxCoder.decode (transfusion_bbox_coder.py:41-125) of it.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings('ignore')

import golden_common as gc  # noqa: E402
import ref_shim  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


@torch.no_grad()
def fusion(R):
    enc = gc.fill_state(R.ISFusionEncoder(**gc.fusion_cfg()), 1).eval()
    bb = gc.fill_state(R.SECONDV2(**gc.backbone_cfg()), 2).eval()
    fpn = gc.fill_state(R.SECONDFPN(in_channels=[128, 256], out_channels=[256, 256], upsample_strides=[1, 2],
                                    norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01),
                                    upsample_cfg=dict(type='deconv', bias=False), use_conv_for_no_stride=True), 3).eval()
    inp = gc.fusion_inputs()
    bs = 2
    pts_metas = dict(pillars=inp['pillars'], pillars_num_points=None, pillar_coors=inp['pillar_coors'], pts=None, pillar_size=None)
    kw = dict(pts_metas=pts_metas, img_metas=inp['img_metas'], lidar2img=inp['lidar2img'],
              img_aug_matrix=inp['img_aug_matrix'], lidar_aug_matrix=inp['lidar_aug_matrix'])
    img_bev = enc.img_fv_to_bev([inp['img_feats'][1]], bs, **kw)
    bev = enc.conv_fusion(torch.cat([img_bev, inp['lidar_feats']], 1))
    torch.manual_seed(0)                                       # the reference shuffles tokens (randperm) in eval too
    feats, hm = enc(inp['img_feats'], inp['lidar_feats'], bs, pts_backbone=bb, **kw)
    neck = fpn(feats)[0]
    np.savez_compressed(os.path.join(HERE, 'ref_fusion.npz'), feat0=_np(feats[0]), feat1=_np(feats[1]), heatmap=_np(hm),
                        img_bev_c8=_np(img_bev[:, ::8]), bev_c8=_np(bev[:, ::8]), neck_c8=_np(neck[:, ::8]),
                        keys_enc=np.array(sorted(enc.state_dict().keys())), keys_bb=np.array(sorted(bb.state_dict().keys())),
                        keys_fpn=np.array(sorted(fpn.state_dict().keys())))
    print('fusion:', [tuple(f.shape) for f in feats], tuple(hm.shape), 'neck', tuple(neck.shape),
          'nonzero img_bev cells', int((img_bev.abs().sum(1) > 0).sum()))


@torch.no_grad()
def fusion_full(R):
    """ref_fusion_full.npz: the same reference run at the CONFIG's sizes (180x180 BEV, B = 1, 6000 pillars): 4096 sampled
    values + first / second moments of every output (the maps themselves are 20 MB)."""
    enc = gc.fill_state(R.ISFusionEncoder(**gc.fusion_cfg_full()), 1).eval()
    bb = gc.fill_state(R.SECONDV2(**gc.backbone_cfg()), 2).eval()
    inp = gc.fusion_inputs(bs=1, bev=180, pillars_per_frame=6000)
    pts_metas = dict(pillars=inp['pillars'], pillars_num_points=None, pillar_coors=inp['pillar_coors'], pts=None, pillar_size=None)
    torch.manual_seed(0)
    feats, hm = enc(inp['img_feats'], inp['lidar_feats'], 1, pts_backbone=bb, pts_metas=pts_metas, img_metas=inp['img_metas'],
                    lidar2img=inp['lidar2img'], img_aug_matrix=inp['img_aug_matrix'], lidar_aug_matrix=inp['lidar_aug_matrix'])
    out = {}
    for name, t in (('feat0', feats[0]), ('feat1', feats[1]), ('heatmap', hm)):
        flat = _np(t).reshape(-1)
        out[name + '_samples'] = flat[gc.sample_positions(t.shape)]
        out[name + '_moments'] = np.array([flat.mean(), np.abs(flat).mean(), (flat.astype(np.float64) ** 2).mean(), np.abs(flat).max()])
        out[name + '_shape'] = np.array(t.shape)
    np.savez_compressed(os.path.join(HERE, 'ref_fusion_full.npz'), **out)
    print('fusion_full:', {k: v.tolist() for k, v in out.items() if k.endswith('_shape')},
          {k: np.round(v, 4).tolist() for k, v in out.items() if k.endswith('_moments')})


@torch.no_grad()
def sst(R):
    cfg = gc.fusion_cfg()
    d = 128
    layer = R.SSTInputLayerV2(window_shape=cfg['region_shape'][0], sparse_shape=cfg['grid_size'][0], shuffle_voxels=False,
                              drop_info=cfg['region_drop_info'][0], pos_temperature=1000, normalize_pos=False, pos_embed=d).eval()
    net = gc.fill_state(R.SSTv2(d_model=[d] * 4, nhead=[8] * 4, num_blocks=1, dim_feedforward=[d] * 4,
                                output_shape=cfg['grid_size'][0][:2], in_channel=d), 4).eval()
    bs, g = 2, gc.BEV
    feats = gc.rand('sst_tokens', (bs * g * g, d), 0)
    b, y, x = torch.meshgrid(torch.arange(bs), torch.arange(g), torch.arange(g), indexing='ij')
    # the encoder's grid coordinates are (b, 0, x-index, y-index) of a meshgrid: any (b,z,y,x) rows do
    coors = torch.stack([b, torch.zeros_like(b), y, x], -1).reshape(-1, 4).int()
    info = layer(feats, coors, bs)
    out = {}
    for i in range(2):
        out['flat2win_%d' % i] = _np(info['flat2win_inds_shift%d' % i][0][0])
        out['pos_%d_c4' % i] = _np(info['pos_dict_shift%d' % i][0][..., ::4])
        out['mask_%d' % i] = _np(info['key_mask_shift%d' % i][0])
        out['batch_win_inds_%d' % i] = _np(info['batch_win_inds_shift%d' % i])
        out['coors_in_win_%d' % i] = _np(info['coors_in_win_shift%d' % i])
    y0 = net(info)[0]
    out['sst_out'] = _np(y0)
    out['keys'] = np.array(sorted(net.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, 'ref_sst.npz'), **out)
    print('sst:', {k: v.shape for k, v in out.items() if k != 'keys'})


@torch.no_grad()
def vfe(R):
    net = gc.fill_state(R.DynamicVFE(**gc.vfe_cfg()), 5).eval()
    pts, coors = gc.vfe_inputs()
    f, c = net(pts, coors)
    sc = R.DynamicScatter(gc.VFE_VOXEL, gc.VFE_RANGE, True)
    fm, cm = sc(pts, coors)
    np.savez_compressed(os.path.join(HERE, 'ref_vfe.npz'), feats=_np(f), coors=_np(c), mean_feats=_np(fm), mean_coors=_np(cm),
                        keys=np.array(sorted(net.state_dict().keys())))
    print('vfe:', tuple(f.shape), tuple(c.shape), 'in-range points', int((coors[:, 1] >= 0).sum()))


@torch.no_grad()
def sparse(R):
    net = gc.fill_state(R.SparseEncoder(**gc.sparse_cfg()), 6).eval()
    feats, coors = gc.sparse_inputs()
    out, enc_feats, _ = net(feats, coors, 2)
    counts = np.array([e.indices.shape[0] for e in enc_feats])
    np.savez_compressed(os.path.join(HERE, 'ref_sparse.npz'), out=_np(out), stage_voxels=counts,
                        stage1_coors=_np(enc_feats[1].indices), keys=np.array(sorted(net.state_dict().keys())))
    print('sparse:', tuple(out.shape), 'active sites per stage', counts.tolist(), 'nonzero', int((out != 0).sum()))


@torch.no_grad()
def swin(R):
    R = ref_shim.install_swin()
    net = gc.fill_state(R.SwinTransformer(**gc.swin_cfg()), 9)
    net.eval()                                       # the reference's train() returns None
    outs = net(gc.swin_input())
    np.savez_compressed(os.path.join(HERE, 'ref_swin.npz'), out0=_np(outs[0]), out1=_np(outs[1]), out2=_np(outs[2]),
                        rel_index=_np(net.stages[0].blocks[0].attn.w_msa.relative_position_index),
                        keys=np.array(sorted(net.state_dict().keys())))
    print('swin:', [tuple(o.shape) for o in outs])


@torch.no_grad()
def lss(R):
    net = gc.fill_state(R.GeneralizedLSSFPN(**gc.lss_cfg()), 8).eval()
    outs = net(gc.lss_inputs())
    np.savez_compressed(os.path.join(HERE, 'ref_lss.npz'), out0_c8=_np(outs[0][:, ::8]), out1_c4=_np(outs[1][:, ::4]),
                        keys=np.array(sorted(net.state_dict().keys())))
    print('lss:', [tuple(o.shape) for o in outs])


@torch.no_grad()
def head(R):
    R = ref_shim.install_head()
    net = gc.fill_state(R.TransFusionHeadV2(**gc.head_cfg()), 7).eval()
    x = gc.head_input()
    res = net([x], None, None)                                  # tuple(list[dict]) as the detector receives it
    p = res[0][0]
    out = {k: _np(v) for k, v in p.items()}
    out['query_labels'] = _np(net.query_labels)
    # get_bboxes (:1278-1409) wraps the result in metas[0]['box_type_3d'] (a box class of mmdet3d.core, absent here) and
    # asserts batch 1; its arithmetic up to that point is the coder's decode over the last-layer predictions:
    n = net.num_proposals
    score = p['heatmap'][..., -n:].sigmoid()
    one_hot = torch.nn.functional.one_hot(net.query_labels, num_classes=net.num_classes).permute(0, 2, 1)
    score = score * p['query_heatmap_score'] * one_hot
    dec = net.bbox_coder.decode(score, p['rot'][..., -n:].clone(), p['dim'][..., -n:].clone(), p['center'][..., -n:].clone(),
                                p['height'][..., -n:].clone(), p['vel'][..., -n:].clone(), filter=True)
    for i, d in enumerate(dec):
        out['box%d_bboxes' % i], out['box%d_scores' % i], out['box%d_labels' % i] = _np(d['bboxes']), _np(d['scores']), _np(d['labels'])
    out['keys'] = np.array(sorted(net.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, 'ref_head.npz'), **out)
    print('head:', {k: v.shape for k, v in out.items() if k != 'keys'})


def sweeps(R):
    """the reference's LoadPointsFromFile + LoadPointsFromMultiSweeps on synthetic sweep files written to a temp dir."""
    import tempfile
    R = ref_shim.install_loading()
    key, ts, sw = gc.sweep_inputs()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        key.tofile(os.path.join(d, 'key.bin'))
        infos = []
        for i, s in enumerate(sw):
            path = os.path.join(d, 'sweep%d.bin' % i)
            s['points'].tofile(path)
            infos.append(dict(data_path=path, timestamp=s['timestamp'], sensor2lidar_rotation=s['sensor2lidar_rotation'],
                              sensor2lidar_translation=s['sensor2lidar_translation']))
        for tag, kw, n_sw in (('plain', dict(), 10), ('close', dict(remove_close=True), 10), ('test3', dict(test_mode=True, sweeps_num=3), 10),
                              ('pad', dict(pad_empty_sweeps=True, remove_close=True, sweeps_num=2), 0)):
            res = dict(pts_filename=os.path.join(d, 'key.bin'), timestamp=ts, sweeps=infos[:n_sw])
            res = R.LoadPointsFromFile(coord_type='LIDAR', load_dim=5, use_dim=5)(res)
            res = R.LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], **kw)(res)
            out[tag] = res['points'].tensor.numpy()
            print('sweeps', tag, out[tag].shape)
            if tag == 'plain':      # PointsRangeFilter's arithmetic: the reference's own BasePoints.in_range_3d on a 40 m range
                out['plain_in_range40'] = res['points'].in_range_3d(np.array([-40, -40, -3, 40, 40, 1], dtype=np.float32)).numpy()
    np.savez_compressed(os.path.join(HERE, 'ref_sweeps.npz'), **out)


def images(R):
    """the reference's ImageAug3D + ImageNormalize (transforms_3d.py:25-145, those lines executed unmodified; the rest of that
    file needs numba / cv2 / box classes) on synthetic uint8 frames: test branch, and the train branch with seeded draws
    (rot_lim = 0: PIL's rotate is outside the device path)."""
    from PIL import Image
    import torchvision
    from typing import Any, Dict
    src = open(os.path.join(ref_shim.REF, 'mmdet3d/datasets/pipelines/transforms_3d.py')).read().split('\n')
    ns = dict(np=np, torch=torch, torchvision=torchvision, Image=Image, Any=Any, Dict=Dict, PIPELINES=ref_shim.Registry('pipeline'))
    exec('\n'.join(src[23:145]), ns)
    imgs = gc.camera_images()
    out = {}
    h, w = imgs[0].shape[:2]
    for tag, cfg, seed in (('test', gc.IMG_AUG_TEST, None),
                           ('train', dict(gc.IMG_AUG_TEST, resize_lim=[0.57, 0.825], rand_flip=True, is_train=True), 4),
                           ('train2', dict(gc.IMG_AUG_TEST, resize_lim=[0.57, 0.825], rand_flip=True, is_train=True), 9)):
        if seed is not None:
            np.random.seed(seed)
        data = dict(img=[Image.fromarray(i) for i in imgs], ori_shape=(w, h))
        data = ns['ImageAug3D'](**cfg)(data)
        data = ns['ImageNormalize'](**gc.IMG_NORM)(data)
        out[tag + '_img'] = np.stack([t.numpy() for t in data['img']])
        out[tag + '_aug'] = np.stack(data['img_aug_matrix'])
        print('images', tag, out[tag + '_img'].shape, out[tag + '_aug'][0][:2, [0, 1, 3]].tolist())
    np.savez_compressed(os.path.join(HERE, 'ref_images.npz'), **out)


if __name__ == '__main__':
    R = ref_shim.install()
    which = sys.argv[1:] or ['fusion', 'fusion_full', 'sst', 'vfe', 'sparse', 'head', 'lss', 'swin', 'sweeps', 'images']
    for w in which:
        globals()[w](R)
    for f in sorted(os.listdir(HERE)):
        if f.startswith('ref_') and f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')
