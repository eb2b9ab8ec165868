"""Dense 3x3 convolutions on the library's own kernels (is-fusion_b200/conv2d.py: TMA-halo tcgen05 conv, 9-offset gather
implicit GEMM).

CPU: the host-side logic — closed-form neighbour table, K-major weight packing, channel-group split (768 = 3 x 256),
output-width padding — against F.conv2d, through a pure-torch statement of what the kernel computes.
GPU: both kernels against the fp32 convolution of the same fp16 operands at the neck's layer sizes, ragged maps, batch > 1,
both strides; determinism and zero padding of the TMA kernel; the whole neck on either kernel."""
import os

import pytest
import torch
import torch.nn.functional as F



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


def _ref_conv(x, w, b, stride, relu):
    """fp32 reference of the op on the fp16-rounded operands (cuDNN fp32, TF32 off)."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float(), stride=stride, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return (torch.relu(y) if relu else y).permute(0, 2, 3, 1)


# the neck's dense convs at their own sizes (conv_fusion, 128->128 @180, SECONDV2 down-sampling, 256->256 @90, the 10-class
# heat-map conv) + maps that do not divide into the kernel's 4x32 / 8x32 pixel tiles + batch > 1
CASES = [(768, 128, 1, 180, 180, 1, True), (128, 128, 1, 180, 180, 2, True), (128, 256, 2, 180, 180, 1, True),
         (256, 256, 1, 90, 90, 2, True), (128, 10, 1, 180, 180, 1, False), (64, 64, 1, 37, 45, 3, True), (64, 32, 2, 10, 6, 1, False),
         (128, 256, 1, 256, 256, 1, True), (256, 128, 2, 256, 256, 1, True)]


@pytest.mark.gpu
@pytest.mark.parametrize('backend', ['tma', 'gather'])
@pytest.mark.parametrize('cin,cout,stride,h,w,bs,relu', CASES)
def test_own_conv_matches_fp32_reference(backend, cin, cout, stride, h, w, bs, relu):
    """both own kernels against the fp32 convolution of the same fp16 operands: <= 1e-3 relative (fp16 output rounding)."""
    from isfusion_b200 import conv2d
    torch.manual_seed(1)
    x = (torch.randn(bs, h, w, cin, device='cuda') * 0.5).half()
    wt = (torch.randn(cout, cin, 3, 3, device='cuda') / (9 * cin) ** 0.5).half().contiguous(memory_format=torch.channels_last)
    b = torch.randn(cout, device='cuda') * 0.1
    packed = (wt, b.half(), b.float())
    fn = conv2d.conv3x3_tma if backend == 'tma' else conv2d.conv3x3_gather
    out = fn(x, packed, stride, relu)
    assert out is not None
    torch.cuda.synchronize()
    ref = _ref_conv(x, wt, b.half() if backend == 'gather' else b, stride, relu)
    assert out.shape == ref.shape
    err = (out.float() - ref).abs().max().item()
    assert ((out.float() - ref).norm() / ref.norm()).item() < 1e-3
    assert err <= 1e-3 * ref.abs().max().item() + 1e-3, err


@pytest.mark.gpu
def test_tma_conv_is_deterministic_and_padding_is_zero():
    """same bits on every run; a one-hot input reproduces the (flipped) kernel around it, including at the map border."""
    from isfusion_b200 import conv2d
    torch.manual_seed(2)
    cin, cout, g = 64, 64, 45
    wt = (torch.randn(cout, cin, 3, 3, device='cuda') / 8).half().contiguous(memory_format=torch.channels_last)
    packed = (wt, torch.zeros(cout, device='cuda').half(), torch.zeros(cout, device='cuda'))
    x = (torch.randn(2, g, g, cin, device='cuda')).half()
    a = conv2d.conv3x3_tma(x, packed, 1, False)
    b = conv2d.conv3x3_tma(x, packed, 1, False)
    assert torch.equal(a, b)
    for (py, px) in ((0, 0), (0, g - 1), (g - 1, 17), (20, 31), (20, 32), (3, 4)):
        x = torch.zeros(1, g, g, cin, device='cuda', dtype=torch.float16)
        x[0, py, px, 5] = 1.0
        y = conv2d.conv3x3_tma(x, packed, 1, False)[0].float()
        want = torch.zeros(g, g, cout, device='cuda')
        for dy in range(3):
            for dx in range(3):
                oy, ox = py - dy + 1, px - dx + 1
                if 0 <= oy < g and 0 <= ox < g:
                    want[oy, ox] = wt[:, 5, dy, dx].float()
        assert torch.equal(y, want), (py, px)


@pytest.mark.gpu
def test_neck_with_gather_convs_matches_tma_neck():
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
    old = conv2d.BACKEND
    try:
        conv2d.BACKEND = 'tma'
        rf, rhm = net(pts, feats, metas, **kw)
        conv2d.BACKEND = 'gather'
        gf, ghm = net(pts, feats, metas, **kw)
        torch.cuda.synchronize()
    finally:
        conv2d.BACKEND = old
    assert ((ghm.float() - rhm.float()).norm() / rhm.float().norm()).item() < 2e-3
    for a, b in zip(gf, rf):
        assert ((a.float() - b.float()).norm() / b.float().norm()).item() < 2e-3
