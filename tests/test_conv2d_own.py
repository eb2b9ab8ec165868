"""Dense 3x3 convolutions on the library's own implicit-GEMM kernel (is-fusion_b200/conv2d.py, opt-in ISF_OWN_CONV=1).

CPU: the host-side logic — closed-form neighbour table, K-major weight packing, channel-group split (768 = 3 x 256),
output-width padding — against F.conv2d, through a pure-torch statement of what the kernel computes.
GPU (bring-up, ISF_EXPERIMENTAL=1): the kernel itself against the cuDNN path, layer by layer and for the whole neck."""
import os

import pytest
import torch
import torch.nn.functional as F

experimental = pytest.mark.skipif(os.environ.get('ISF_EXPERIMENTAL') != '1', reason='opt-in path: set ISF_EXPERIMENTAL=1')


@pytest.mark.parametrize('cin,cout,stride,h,w,relu', [(768, 128, 1, 9, 7, True), (128, 128, 1, 12, 12, True), (128, 256, 2, 12, 12, True),
                                                     (128, 256, 2, 11, 9, True), (64, 10, 1, 8, 10, False), (16, 32, 1, 5, 5, True)])
def test_dense_rulebook_and_packing_reproduce_conv2d(cin, cout, stride, h, w, relu):
    from isfusion_b200 import conv2d
    torch.manual_seed(0)
    bs = 2
    x = torch.randn(bs, h, w, cin)
    wt = torch.randn(cout, cin, 3, 3) / (9 * cin) ** 0.5
    bias = torch.randn(cout) * 0.1
    groups = conv2d.channel_groups(cin)
    assert groups == (3 if cin == 768 else 1)
    nbr, oh, ow = conv2d.dense_rulebook(bs, h, w, stride, groups, x.device)
    wp, shift, cpad = conv2d.pack_dense_weight(wt, bias)
    assert nbr.shape[0] == 9 * groups <= 27 and nbr.shape[1] % 128 == 0 and cpad in (32, 64, 128, 256)
    assert wp.shape[1] % 64 == 0 and wp.shape[1] >= 9 * cin
    got = conv2d.emulate(x, nbr, wp, shift, groups, cout, relu)[:bs * oh * ow].view(bs, oh, ow, cout)
    want = F.conv2d(x.permute(0, 3, 1, 2), wt.half().float(), bias, stride=stride, padding=1)
    want = (torch.relu(want) if relu else want).permute(0, 2, 3, 1)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)
    assert (nbr[:, bs * oh * ow:] == -1).all()


def test_unsupported_shapes_fall_back():
    from isfusion_b200 import conv2d
    assert conv2d.channel_groups(10) is None and conv2d.channel_groups(96) == 3 and conv2d.channel_groups(512) == 1
    assert conv2d.pack_dense_weight(torch.zeros(300, 16, 3, 3), None) is None          # wider than one 256-column tile


@pytest.mark.gpu
@experimental
@pytest.mark.parametrize('cin,cout,stride,g,relu', [(768, 128, 1, 180, True), (128, 128, 1, 180, True), (128, 256, 2, 180, True),
                                                   (256, 256, 1, 90, True), (64, 10, 1, 180, False)])
def test_own_conv_matches_cudnn(cin, cout, stride, g, relu):
    from isfusion_b200 import conv2d
    from isfusion_b200.fusion_encoder import _conv
    torch.manual_seed(1)
    x = (torch.randn(2, g, g, cin, device='cuda') * 0.5).half()
    w = (torch.randn(cout, cin, 3, 3, device='cuda') / (9 * cin) ** 0.5).half().contiguous(memory_format=torch.channels_last)
    b = (torch.randn(cout, device='cuda') * 0.1).half()
    try:
        conv2d.ENABLED = False
        ref = _conv(x, (w, b), stride=stride, relu=relu).float()
        conv2d.ENABLED = True
        out = _conv(x, (w, b), stride=stride, relu=relu).float()
        torch.cuda.synchronize()
    finally:
        conv2d.ENABLED = False
    assert out.shape == ref.shape
    assert ((out - ref).norm() / ref.norm()).item() < 2e-3


@pytest.mark.gpu
@experimental
def test_neck_with_own_convs_matches_cudnn_neck():
    import synth
    from __graft_entry__ import make_pair
    from isfusion_b200 import conv2d
    _, net = make_pair(0)
    net = net.cuda().eval()
    pts = [synth.make_points(20000, 'S', 5).cuda()]
    l2i, aug, laug = synth.make_calib(1)
    feats = [t.cuda() for t in synth.make_img_feats(1)]
    metas = [dict(input_shape=(256, 704))]
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    try:
        conv2d.ENABLED = False
        rf, rhm = net(pts, feats, metas, **kw)
        conv2d.ENABLED = True
        gf, ghm = net(pts, feats, metas, **kw)
        torch.cuda.synchronize()
    finally:
        conv2d.ENABLED = False
    assert ((ghm.float() - rhm.float()).norm() / rhm.float().norm()).item() < 5e-3
    for a, b in zip(gf, rf):
        assert ((a.float() - b.float()).norm() / b.float().norm()).item() < 1e-2
