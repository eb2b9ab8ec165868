"""GPU parity of the fusion-encoder kernels and of the whole neck against the oracle."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import fusion_ops as fo
from oracle import model as omodel
from oracle import sst_ops as ss

pytestmark = pytest.mark.gpu
# north_star's bar is 1e-3 relative: measured on B200 (profiles/call34_r02.log) SST level 5.2-6.1e-4, instance->scene block 4.5e-4,
# SECONDV2 stage 1 (6 convs + the stride-2 conv) 7.7e-4 / 8.4e-4; stage 2 stacks 5 more fp16 conv outputs on stage 1's (9.6e-4)
SST_BOUND, I2S_BOUND, SECOND_BOUND = 1e-3, 1e-3, (1e-3, 1.5e-3)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp(min=1e-12)).item(), (a - b).abs().max().item()


def chk(name, value, bound):
    """asserted bound next to the measured value (pytest -s shows the line; profiles/call*_r02.log keep them)."""
    print('parity %-46s %.2e  (bound %.1e)' % (name, value, bound))
    assert value < bound, (name, value, bound)


def build_pair(seed=0, shuffle=False):
    import isfusion_b200.detector as det
    torch.manual_seed(seed)
    cfg = omodel.neck_config()
    ref = synth.randomize_bn(omodel.ISFusionNeck(cfg, shuffle_voxels=shuffle)).eval()
    net = det.ISFusionDetector(
        pts_voxel_layer=dict(point_cloud_range=cfg['pc_range'], max_num_points=-1, voxel_size=cfg['voxel_size'], max_voxels=(-1, -1)),
        pts_voxel_encoder=dict(type='DynamicVFE', **cfg['pts_voxel_encoder']),
        pts_middle_encoder=dict(type='SparseEncoder', **cfg['pts_middle_encoder']),
        fusion_encoder=dict(type='ISFusionEncoder', **cfg['fusion_encoder']),
        pts_backbone=dict(type='SECONDV2', **cfg['pts_backbone']),
        detach=True, pc_range=cfg['pc_range'], voxel_size=cfg['voxel_size'], out_size_factor=8)
    net.load_state_dict(ref.state_dict())
    return ref, net.cuda().eval()


def test_msda_golden_and_oracle(golden_dir):
    from isfusion_b200.fusion_encoder import ms_deform_attn
    g = np.load(os.path.join(golden_dir, 'msda_ref.npz'))
    # case A: ops/test.py shapes (2 levels), fp32
    shapes = torch.from_numpy(g['a_shapes']).cuda()
    lvl = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    out = ms_deform_attn(torch.from_numpy(g['a_value']).cuda(), shapes, lvl, torch.from_numpy(g['a_loc']).cuda(),
                         torch.from_numpy(g['a_aw']).cuda())
    assert torch.allclose(out.cpu(), torch.from_numpy(g['a_out']), rtol=1e-2, atol=1e-3)   # tolerance of ops/test.py:58
    assert (out.cpu() - torch.from_numpy(g['a_out'])).abs().max() < 1e-6
    # case B: IGF head layout, fp16 value
    shapes = torch.tensor([[40, 44]]).cuda()
    value = torch.from_numpy(g['b_value_f16']).cuda()
    out = ms_deform_attn(value, shapes, torch.zeros(1, dtype=torch.long).cuda(), torch.from_numpy(g['b_loc']).cuda(),
                         torch.from_numpy(g['b_aw']).cuda())
    assert (out.cpu() - torch.from_numpy(g['b_out'])).abs().max() < 1e-4
    # full IGF shape against the oracle restatement
    torch.manual_seed(0)
    value = torch.randn(2, 32400, 8, 16)
    loc = torch.rand(2, 200, 8, 1, 16, 2) * 1.1 - 0.05
    aw = torch.softmax(torch.randn(2, 200, 8, 16), -1).view(2, 200, 8, 1, 16)
    shapes = torch.tensor([[180, 180]])
    ref = fo.ms_deform_attn_core_pytorch(value, shapes, loc, aw)
    out = ms_deform_attn(value.cuda(), shapes.cuda(), torch.zeros(1, dtype=torch.long).cuda(), loc.cuda(), aw.cuda())
    assert (out.cpu() - ref).abs().max() < 1e-4


@pytest.mark.parametrize('g,c,b', [(18, 128, 2), (21, 256, 1), (180, 128, 1), (90, 256, 2)])
def test_sst_level_parity(g, c, b):
    """SSTInputLayerV2 + SSTv2 (one shifted-window block) on a dense grid; partial windows when 6 does not divide g."""
    from isfusion_b200.fusion_encoder import SSTInputLayerV2, SSTv2
    torch.manual_seed(g + c)
    di = {0: {'max_tokens': 36, 'drop_range': (0, 100000)}}
    rl = ss.SSTInputLayerV2(di, (6, 6, 1), [g, g, 1], shuffle_voxels=True, pos_temperature=1000, pos_embed=c)
    rn = ss.SSTv2([c] * 4, [8] * 4, 1, [c] * 4, output_shape=[g, g], in_channel=128 if c == 128 else None).eval()
    for n_, p in rn.named_parameters():
        if 'bias' in n_:
            p.data.normal_(0, 0.05)
    ys, xs = torch.meshgrid(torch.arange(g), torch.arange(g), indexing='ij')
    coords = torch.cat([torch.stack([torch.full((g * g,), i), torch.zeros(g * g, dtype=torch.long), ys.flatten(), xs.flatten()], 1)
                        for i in range(b)]).int()
    feats = torch.randn(b * g * g, 128 if c == 128 else c).half().float()
    ref = rn(rl(feats, coords, b))[0]
    gl = SSTInputLayerV2(di, (6, 6, 1), [g, g, 1], pos_temperature=1000, pos_embed=c)
    gn = SSTv2([c] * 4, [8] * 4, 1, [c] * 4, output_shape=[g, g], in_channel=128 if c == 128 else None)
    gn.load_state_dict(rn.state_dict())
    gn = gn.cuda().eval()
    out = gn(gl(feats.half().cuda(), None, b))[0]
    assert out.shape == ref.shape
    r, m = rel(out, ref)
    chk('SST level g=%d c=%d rel-L2' % (g, c), r, SST_BOUND)       # 2 encoder layers: 10 GEMMs, 2 attention cores, 4 LayerNorms


def test_cam2bev_parity():
    from isfusion_b200.fusion_encoder import ISFusionEncoder
    from isfusion_b200.voxel_layer import Voxelization
    from oracle import voxel_ops as vo
    cfg = omodel.neck_config()['fusion_encoder']
    B = 2
    pts = [synth.make_points(20000, 'S', s) for s in range(B)]
    l2i, aug, laug = synth.make_calib(B)
    laug[1, :3, :3] = torch.tensor([[0.98, -0.199, 0.], [0.199, 0.98, 0.], [0., 0., 1.]]) * 1.02    # exercise the inverse aug
    laug[1, :3, 3] = torch.tensor([0.5, -0.3, 0.1])
    feats = synth.make_img_feats(B)
    feats[1] = feats[1].half().float()
    ref_enc = fo.ISFusionEncoder(**cfg)
    pv, pc, pn = [], [], []
    for i, p in enumerate(pts):
        v, c, n = vo.hard_voxelize(p, [0.6, 0.6, 8.0], synth.PC_RANGE, 12, 60000)
        pv.append(v); pc.append(torch.nn.functional.pad(c, (1, 0), value=i)); pn.append(n)
    pm = dict(pillars=torch.cat(pv), pillar_coors=torch.cat(pc), pillars_num_points=torch.cat(pn))
    kw = dict(img_metas=[dict(input_shape=(256, 704))] * B, lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    ref = ref_enc.img_fv_to_bev([feats[1]], B, pts_metas=pm, **kw)
    enc = ISFusionEncoder(**cfg).cuda()
    pmg = {k: v.cuda() for k, v in pm.items()}
    out = enc.img_fv_to_bev([feats[1].cuda()], B, pts_metas=pmg, **kw)
    out = out.permute(0, 3, 1, 2)
    r, m = rel(out, ref)
    print('cam2bev rel %.2e max %.2e (ref max %.2f, nonzero cells %d)' % (r, m, ref.abs().max(), (ref.abs().sum(1) > 0).sum()))
    assert (ref.abs().sum(1) > 0).sum() > 1000
    assert r < 1e-3 and m < 2e-3 * max(1.0, ref.abs().max().item())       # measured rel-L2 2.1e-4


def test_heatmap_topk():
    import ctypes
    from isfusion_b200 import _lib as L
    torch.manual_seed(0)
    B, C, G, K = 2, 10, 180, 200
    logits = torch.randn(B, C, G, G)
    hm = logits.sigmoid()
    lm = torch.zeros_like(hm)
    lm[:, :, 1:-1, 1:-1] = torch.nn.functional.max_pool2d(hm, 3, stride=1, padding=0)
    lm[:, 8] = hm[:, 8]; lm[:, 9] = hm[:, 9]
    sc = (hm * (hm == lm)).view(B, -1)
    ref_scores, ref_idx = sc.sort(dim=-1, descending=True, stable=True)
    ws_bytes = ctypes.c_size_t(0)
    L.check(L.lib().isf_heatmap_topk_workspace_size(B, C, G, G, ctypes.byref(ws_bytes)))
    ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device='cuda')
    top = torch.empty((B, K), dtype=torch.int32, device='cuda')
    score = torch.empty((B, K), dtype=torch.float32, device='cuda')
    lg = logits.cuda()
    L.check(L.lib().isf_heatmap_topk(L.ptr(lg), B, C, G, G, 8, K, L.ptr(top), L.ptr(score), L.ptr(ws),
                                     ctypes.c_size_t(ws.numel()), L.stream()))
    assert torch.equal(top.cpu().long(), ref_idx[:, :K] % (G * G))
    assert torch.allclose(score.cpu(), ref_scores[:, :K], atol=1e-6)


def test_instance_to_scene_parity():
    from isfusion_b200 import _lib as L
    from isfusion_b200.gemm import gemm, pack_weight
    torch.manual_seed(0)
    B, C, G, K = 2, 128, 180, 200
    ref = fo.Instane2SceneAtt(C).eval()
    ref.multihead_attn.in_proj_bias.data.normal_(0, 0.05)
    q = (torch.randn(B, C, G * G) * 0.5).half().float()
    key = torch.randn(B, C, K).half().float()
    scene = (torch.randn(B, C, G, G) * 0.3).half().float()
    want = ref(q, key, scene, B, G)
    mha = ref.multihead_attn
    qt = q.permute(0, 2, 1).reshape(B * G * G, C).half().cuda().contiguous()
    qq = gemm(qt, pack_weight(mha.in_proj_weight.data[:C].cuda()), shift=mha.in_proj_bias.data[:C].cuda().contiguous())
    kv = torch.nn.functional.linear(key.permute(0, 2, 1), mha.in_proj_weight.data[C:], mha.in_proj_bias.data[C:]).half().cuda().contiguous()
    att = torch.empty_like(qq)
    L.check(L.lib().isf_i2s_attention(L.ptr(qq), L.ptr(kv), B, G * G, K, C, 8, L.ptr(att), L.stream()))
    qi = gemm(att, pack_weight(mha.out_proj.weight.data.cuda()), shift=mha.out_proj.bias.data.cuda().contiguous(), residual=qt,
              ln=(ref.norm.weight.data.cuda().contiguous(), ref.norm.bias.data.cuda().contiguous(), ref.norm.eps))
    sc = scene.permute(0, 2, 3, 1).half().cuda().contiguous()
    out = torch.empty_like(sc)
    import ctypes
    nb = ctypes.c_size_t(0)
    L.check(L.lib().isf_channel_attention_workspace_size(B, G, C, ctypes.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device='cuda')
    L.check(L.lib().isf_channel_attention(L.ptr(sc), C, L.ptr(qi), C, B, G, C, L.ptr(out), C, L.ptr(ws),
                                          ctypes.c_size_t(ws.numel()), L.stream()))
    r, m = rel(out.permute(0, 3, 1, 2), want)
    chk('instance->scene attention block rel-L2', r, I2S_BOUND)


@pytest.mark.parametrize('B,T,K', [(1, 32400, 200), (2, 32400, 200), (1, 8100, 200), (3, 1000, 150), (1, 127, 8), (1, 129, 208)])
def test_i2s_attention_tcgen05_kernel(B, T, K):
    """isf_i2s_attention on tcgen05 tiles (csrc/i2s_tc.cu: TMA-staged Q / K, S and O in TMEM, P through shared memory) against the
    fp32 attention of the same fp16 operands and against the mma.sync kernel (the default; the tcgen05 kernel is opt-in,
    isf_set_option("i2s_tc", 1): measured 3x slower at head dimension 16); ragged last tile, fewer than 208 keys, several frames.
    Prints both kernels' times."""
    from isfusion_b200 import _lib as L
    torch.manual_seed(T + K)
    C, H = 128, 8
    q = (torch.randn(B * T, C) * 0.8).half().cuda()
    kv = (torch.randn(B, K, 2 * C) * 0.8).half().cuda()
    qf = q.float().view(B, T, H, 16).permute(0, 2, 1, 3)
    kf = kv[..., :C].float().view(B, K, H, 16).permute(0, 2, 1, 3)
    vf = kv[..., C:].float().view(B, K, H, 16).permute(0, 2, 1, 3)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) / 4.0, -1) @ vf          # [B, H, T, 16]
    ref = ref.permute(0, 2, 1, 3).reshape(B * T, C)
    outs, times = {}, {}
    try:
        for mode in (1, 0):
            L.set_option('i2s_tc', mode)
            o = torch.full((B * T, C), 7.0, dtype=torch.float16, device='cuda')
            L.check(L.lib().isf_i2s_attention(L.ptr(q), L.ptr(kv), B, T, K, C, H, L.ptr(o), L.stream()))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                L.check(L.lib().isf_i2s_attention(L.ptr(q), L.ptr(kv), B, T, K, C, H, L.ptr(o), L.stream()))
            e1.record()
            torch.cuda.synchronize()
            outs[mode], times[mode] = o, e0.elapsed_time(e1) / 10
    finally:
        L.set_option('i2s_tc', 0)
    print('i2s B=%d T=%d K=%d: tcgen05 %.1f us, mma.sync %.1f us' % (B, T, K, times[1] * 1e3, times[0] * 1e3))
    for mode in (1, 0):
        r, m = rel(outs[mode], ref)
        assert r < 1e-3 and m < 2e-3 * max(1.0, ref.abs().max().item()), (mode, r, m)
    assert rel(outs[1], outs[0])[0] < 1e-3


def test_secondv2_parity():
    from isfusion_b200.fusion_encoder import SECONDV2
    torch.manual_seed(0)
    cfg = omodel.neck_config()['pts_backbone']
    ref = synth.randomize_bn(fo.SECONDV2(**cfg)).eval()
    net = SECONDV2(**cfg)
    net.load_state_dict(ref.state_dict())
    net = net.cuda().eval()
    x = torch.randn(1, 128, 60, 60).half().float()
    rt, rc, rf = ref([x], 'stage1')
    xt = x.half().cuda().contiguous(memory_format=torch.channels_last)
    gt, _, gf = net([xt], 'stage1')
    chk('SECONDV2 stage1 block0 (6 convs) rel-L2', rel(gf, rf)[0], SECOND_BOUND[0])
    chk('SECONDV2 stage1 downsample rel-L2', rel(gt, rt)[0], SECOND_BOUND[0])
    r2 = ref([rt.view(1, 30, 30, 256).permute(0, 3, 1, 2)], 'stage2')[2]
    g2 = net([gt.view(1, 30, 30, 256).permute(0, 3, 1, 2)], 'stage2')[2]
    chk('SECONDV2 stage2 (5 convs more) rel-L2', rel(g2, r2)[0], SECOND_BOUND[1])


@pytest.mark.parametrize('npts,batch', [(20000, 1), (12000, 2)])
def test_neck_end_to_end(npts, batch):
    """BASELINE configs[0]: synthetic frame(s), 6x256x704 cameras, whole neck vs the CPU oracle."""
    ref, net = build_pair(0)
    pts = [synth.make_points(npts, 'S', s) for s in range(batch)]
    l2i, aug, laug = synth.make_calib(batch)
    feats = synth.make_img_feats(batch)
    metas = [dict(input_shape=(256, 704))] * batch
    rf, rhm = ref(pts, feats, metas, lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    gf, ghm = net([p.cuda() for p in pts], [f.cuda() for f in feats], metas, lidar2img=l2i, img_aug_matrix=aug,
                  lidar_aug_matrix=laug)
    dbg_r, dbg_g = ref.fusion_encoder._dbg, net.fusion_encoder._dbg
    r_hm, m_hm = rel(ghm, rhm)
    overlap = np.mean([len(set(dbg_r['top_idx'][b].tolist()) & set(dbg_g['top_idx'][b].cpu().tolist())) / 200.0
                       for b in range(batch)])
    r0, m0 = rel(gf[0], rf[0])
    r1, m1 = rel(gf[1], rf[1])
    print('neck e2e: heatmap rel %.2e | top-200 overlap %.3f | feat0 rel %.2e max %.2e | feat1 rel %.2e max %.2e'
          % (r_hm, overlap, r0, m0, r1, m1))
    assert gf[0].shape == rf[0].shape and gf[1].shape == rf[1].shape and ghm.shape == rhm.shape
    assert r_hm < 1e-3 and overlap > 0.95
    assert r0 < 2e-3 and r1 < 2.5e-3       # measured 9.0e-4 / 1.4e-3 (fp16 storage through ~20 / ~35 layers)


def test_sync_free_and_graph_paths_match_exact_path():
    """capacity buffers + device-side counts (nosync) and the one-graph replay give the same bits as the exact path."""
    from isfusion_b200.detector import GraphedNeck
    import isfusion_b200.spconv as spmod
    _, net = build_pair(0)
    B, N = 2, 15000
    pts = [synth.make_points(N, 'S', 10 + s).cuda() for s in range(B)]
    l2i, aug, laug = synth.make_calib(B)
    feats = [f.cuda() for f in synth.make_img_feats(B)]
    metas = [dict(input_shape=(256, 704))] * B
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    ef, ehm = net(pts, feats, metas, **kw)
    sf, shm = net(pts, feats, metas, nosync=True, **kw)
    assert not spmod.check_overflow()
    for a, b in zip(ef, sf):
        assert torch.equal(a, b)
    assert torch.equal(ehm, shm)
    caps = GraphedNeck.calibrate(net, [pts], feats, metas, **kw)
    gn = GraphedNeck(net, B, max(p.shape[0] for p in pts) + 500, calib=kw, capacities=caps)   # padded points
    gn.load(pts, feats[1])
    gn.capture()
    gf, ghm = gn(pts, feats[1])
    torch.cuda.synchronize()
    assert not gn.check()
    for a, b in zip(ef, gf):
        assert torch.equal(a, b)
    # a different frame through the same graph
    pts2 = [synth.make_points(N - 1000, 'S', 20 + s).cuda() for s in range(B)]
    gf2, _ = gn(pts2, feats[1])
    ef2, _ = net(pts2, feats, metas, **kw)
    for a, b in zip(ef2, gf2):
        assert torch.equal(a, b)

    # pipelined host interface (H2D / D2H on a copy stream) returns the same bits
    hp = [[p.cpu().pin_memory() for p in pts], [p.cpu().pin_memory() for p in pts2]]
    hf = feats[1].cpu().pin_memory()
    oh = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in ef] for _ in range(2)]
    t0 = gn.prefetch(hp[0], hf)
    t1 = gn.prefetch(hp[1], hf)
    gn.run_prefetched(t0, oh[0])
    gn.run_prefetched(t1, oh[1])
    torch.cuda.synchronize()
    for a, b in zip(ef, oh[0]):
        assert torch.equal(a.cpu(), b)
    for a, b in zip(ef2, oh[1]):
        assert torch.equal(a.cpu(), b)


def test_two_frames_in_flight_match_exact_path():
    """FramePipeline: two graph lanes over one set of weights replayed concurrently on their own streams give the same
    bits as the exact path, for resident frames and through the pipelined host interface."""
    from isfusion_b200.detector import FramePipeline, GraphedNeck
    _, net = build_pair(0)
    B, N = 1, 15000
    l2i, aug, laug = synth.make_calib(B)
    feats = [f.cuda() for f in synth.make_img_feats(B)]
    metas = [dict(input_shape=(256, 704))] * B
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    frames = [[synth.make_points(N - 700 * f, 'S', 30 + f).cuda()] for f in range(3)]
    want = [net(pts, feats, metas, **kw)[0] for pts in frames]
    caps = GraphedNeck.calibrate(net, frames, feats, metas, **kw)
    pipe = FramePipeline(net, 2, B, N + 100, calib=kw, capacities=caps).capture(frames[0], feats[1])
    got = []
    for i in range(9):                                   # no host sync between submissions: lanes overlap
        k, (out, _) = pipe.submit(frames[i % 3], feats[1])
        with torch.cuda.stream(pipe.streams[k]):
            got.append([t.clone() for t in out])
    pipe.join()
    torch.cuda.synchronize()
    assert not pipe.check()
    for i, g in enumerate(got):
        for a, b in zip(want[i % 3], g):
            assert torch.equal(a, b), i
    # host interface: each lane prefetches its own next frame
    hp = [[p.cpu().pin_memory() for p in pts] for pts in frames]
    hf = feats[1].cpu().pin_memory()
    oh = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in want[0]] for _ in range(4)]
    t = [pipe.prefetch(i % 2, hp[i % 3], hf) for i in range(2)]
    for i in range(4):
        k = i % 2
        ticket = t[k]
        if i + 2 < 4:
            t[k] = pipe.prefetch(k, hp[(i + 2) % 3], hf)
        pipe.submit_prefetched(k, ticket, oh[i])
    pipe.join()
    torch.cuda.synchronize()
    for i in range(4):
        for a, b in zip(want[i % 3], oh[i]):
            assert torch.equal(a.cpu(), b), i


def test_secondfpn_parity():
    """pts_neck: 1x1 conv + 2x2 deconv + BN + ReLU + concat + [X,Y] transpose on the tcgen05 GEMM."""
    from isfusion_b200.fusion_encoder import SECONDFPN
    torch.manual_seed(0)
    cfg = dict(in_channels=[128, 256], out_channels=[256, 256], upsample_strides=[1, 2],
               norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), upsample_cfg=dict(type='deconv', bias=False),
               use_conv_for_no_stride=True)
    ref = synth.randomize_bn(fo.SECONDFPN(**cfg)).eval()
    net = SECONDFPN(**cfg)
    net.load_state_dict(ref.state_dict())
    net = net.cuda().eval()
    x0 = torch.randn(2, 128, 180, 180).half().float()
    x1 = torch.randn(2, 256, 90, 90).half().float()
    want = ref([x0, x1])[0]
    got = net([x0.half().cuda().contiguous(memory_format=torch.channels_last),
               x1.half().cuda().contiguous(memory_format=torch.channels_last)])[0]
    assert got.shape == want.shape == (2, 512, 180, 180)
    r, m = rel(got, want)
    assert r < 2e-3, (r, m)


def test_bench_workload_parity_through_frame_pipeline():
    """THE BENCHMARKED WORKLOAD (bench.py defaults: BASELINE configs[1], 300 k-point 'S' cloud, 6 x 384 x 1056 cameras ->
    24 x 66 stride-16 map, 3 graph lanes, capacity mode) against the CPU oracle on the same frame, through the same
    FramePipeline.submit() the bench times.  Voxel-level index work is covered bit-exactly elsewhere; here the heat-map, the
    top-200 selection and the two output maps of the whole neck.  The selection is discrete: one near-tie among 200 picks
    (fp16 vs fp32 heat-map) swaps an instance and moves the maps by ~1e-2 around it, so the oracle is run on the product's
    selection (test hook `force_top_idx`) and the natural selections are compared separately.  Error budget of the maps:
    ~20 (feat0) / ~35 (feat1) fp16 roundings on the activation path at ~2e-4 relative RMS each -> 1-2e-3 end to end
    (DESIGN.md section 4); per-op bounds are 1e-3."""
    from isfusion_b200.detector import FramePipeline, GraphedNeck
    ref, net = build_pair(0)
    img = (384, 1056)
    frames = [[synth.make_points(300000, 'S', 8 * s)] for s in range(3)]
    l2i, aug, laug = synth.make_calib(1, img)
    feats = synth.make_img_feats(1, img, seed=0)
    metas = [dict(input_shape=img)]
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    dev_frames = [[p.cuda() for p in f] for f in frames]
    dfeats = [t.cuda() for t in feats]
    caps = GraphedNeck.calibrate(net, dev_frames[:2], dfeats, metas, margin=1.35, **kw)
    pipe = FramePipeline(net, 3, 1, 300000, img_hw=img, calib=kw, capacities=caps, device=torch.device('cuda:0'))
    pipe.capture(dev_frames[0], dfeats[1])
    outs = []
    for i in (1, 2, 0, 1, 0):                         # frame 0 goes through two different lanes, other frames in flight
        lane, (o, hm) = pipe.submit(dev_frames[i], dfeats[1])
        if i == 0:
            pipe.join()
            torch.cuda.synchronize()
            outs.append(([t.clone() for t in o], hm.clone()))
    pipe.join()
    torch.cuda.synchronize()
    assert not pipe.check()
    (gf, ghm), (gf2, ghm2) = outs
    for a, b in zip(gf + [ghm], gf2 + [ghm2]):
        assert torch.equal(a, b)                     # lane-independent bits
    ef, ehm = net(dev_frames[0], dfeats, metas, **kw)
    for a, b in zip(list(ef) + [ehm], gf + [ghm]):
        assert torch.equal(a, b)                     # graph lanes == exact eager path
    top_g = net.fusion_encoder._dbg['top_idx']       # reference indexing t = x*G + y, score order
    ref.fusion_encoder.force_top_idx = top_g.cpu()
    try:
        rf, rhm = ref(frames[0], feats, metas, **kw)
    finally:
        ref.fusion_encoder.force_top_idx = None
    dbg_r = ref.fusion_encoder._dbg
    overlap = len(set(dbg_r['top_idx'][0].tolist()) & set(top_g[0].cpu().tolist())) / 200.0
    r_hm, _ = rel(ghm, rhm)
    r0, m0 = rel(gf[0], rf[0])
    r1, m1 = rel(gf[1], rf[1])
    print('bench workload parity: heatmap rel %.2e | top-200 overlap %.3f | feat0 rel %.2e max %.2e | feat1 rel %.2e max %.2e'
          % (r_hm, overlap, r0, m0, r1, m1))
    assert r_hm < 1e-3 and overlap >= 0.98
    assert r0 < 2e-3 and r1 < 2.5e-3


@pytest.mark.parametrize('batch', [1, 2])
def test_fused_instance_decoder_matches_module_path_and_oracle(batch):
    """a11 in one launch (isf_ins_context_decoder: position embedding + 2 x (self-attention, deformable cross-attention, FFN,
    3 LayerNorms) + the k | v projection of the i2s attention; fp16 operands on mma.sync, fp32 accumulation / residual stream)
    against (a) the module-by-module fp32 torch path over the SAME selection and value maps and (b) the CPU oracle."""
    import isfusion_b200.fusion_encoder as FE
    ref, net = build_pair(0)
    pts = [synth.make_points(15000, 'S', 30 + s) for s in range(batch)]
    l2i, aug, laug = synth.make_calib(batch)
    feats = synth.make_img_feats(batch)
    metas = [dict(input_shape=(256, 704))] * batch
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    dpts, dfeats = [p.cuda() for p in pts], [f.cuda() for f in feats]
    old = FE.FUSED_DECODER
    try:
        FE.FUSED_DECODER = False
        mf, mhm = net(dpts, dfeats, metas, **kw)
        x_mod = net.fusion_encoder._dbg['x_ins'].clone()
        FE.FUSED_DECODER = True
        ff, fhm = net(dpts, dfeats, metas, **kw)
        x_fus = net.fusion_encoder._dbg['x_ins'].clone()
        top = net.fusion_encoder._dbg['top_idx'].clone()
        torch.cuda.synchronize()
    finally:
        FE.FUSED_DECODER = old
    assert torch.equal(mhm, fhm)                                  # same heat-map -> same selection on both paths
    r_mod, m_mod = rel(x_fus, x_mod)
    ref.fusion_encoder.force_top_idx = top.cpu()
    try:
        rf, _ = ref(pts, feats, metas, **kw)
    finally:
        ref.fusion_encoder.force_top_idx = None
    x_ref = ref.fusion_encoder._dbg['x_ins'].transpose(1, 2)      # [B,C,K] -> [B,K,C]
    r_ref, m_ref = rel(x_fus, x_ref)
    r0, _ = rel(ff[0], rf[0])
    r1, _ = rel(ff[1], rf[1])
    print('fused decoder: vs module path rel %.2e max %.2e | vs oracle rel %.2e max %.2e | feat0 %.2e feat1 %.2e' % (r_mod, m_mod, r_ref, m_ref, r0, r1))
    assert x_fus.shape == x_mod.shape == x_ref.shape
    assert torch.isfinite(x_fus).all()
    assert r_mod < 2e-3 and r_ref < 2e-3
    assert r0 < 2e-3 and r1 < 2.5e-3
    for a, b in zip(ff, mf):
        assert rel(a, b)[0] < 2e-3
