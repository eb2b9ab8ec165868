"""GPU parity of the bf16 tensor-core path (BASELINE configs[2]: "HSF-only ablation ... bf16 tensor-core path"): the `_bf16` entry
points run the kernels of their f16 siblings with bf16 operand / output encodings.  Each test compares against an fp32 reference of
the SAME bf16-rounded operands (so that only the path's own arithmetic and its bf16 OUTPUT rounding, 2^-9 relative, are measured)
and, for the whole Grid-to-Region level, against the CPU oracle.  Tolerances are bf16's, written out per test; the fp16 path is the
one that meets north_star's 1e-3."""
import pytest
import torch
import torch.nn.functional as F

import synth
from oracle import fusion_ops as fo
from oracle import model as omodel
from oracle import sst_ops as ss

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
BF_OUT = 2.0 ** -8          # a bf16 store rounds by at most 2^-9 relative; 2^-8 leaves room for the fp32 summation order


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp(min=1e-12)).item(), (a - b).abs().max().item()


@pytest.fixture
def bf16_mode():
    from isfusion_b200 import _lib as L
    L.set_dtype(BF)
    yield L
    L.set_dtype(torch.float16)


@pytest.mark.parametrize('m,k,n,mode', [(4096, 128, 128, 'ln'), (32400, 128, 384, 'plain'), (8100, 256, 256, 'gelu'), (300, 128, 128, 'ln'),
                                        (32400, 256, 128, 'res')])
def test_gemm_bf16(bf16_mode, m, k, n, mode):
    from isfusion_b200.gemm import gemm, pack_weight
    torch.manual_seed(m + n)
    a = torch.randn(m, k).to(BF)
    w = (torch.randn(n, k) / k ** 0.5).to(BF)
    shift = torch.randn(n) * 0.1
    res = torch.randn(m, n).to(BF) if mode in ('ln', 'res') else None
    ref = a.float() @ w.float().t() + shift
    if res is not None:
        ref = ref + res.float()
    ln = None
    if mode == 'ln':
        g, b = torch.rand(n) + 0.5, torch.randn(n) * 0.1
        ref = F.layer_norm(ref, (n,), g, b, 1e-5)
        ln = (g.cuda(), b.cuda(), 1e-5)
    if mode == 'gelu':
        ref = F.gelu(ref)
    out = gemm(a.cuda(), pack_weight(w.cuda()), shift=shift.cuda(), residual=None if res is None else res.cuda(),
               act='gelu' if mode == 'gelu' else None, ln=ln)
    assert out.dtype == BF
    r, mx = _rel(out, ref)
    assert mx <= BF_OUT * max(1.0, ref.abs().max().item()) and r <= BF_OUT, (r, mx)
    # fp32 output of the same product: no output rounding left, fp16-path accuracy
    if mode == 'plain':
        o32 = gemm(a.cuda(), pack_weight(w.cuda()), shift=shift.cuda(), out_dtype=torch.float32)
        assert _rel(o32, ref)[0] < 1e-5


@pytest.mark.parametrize('cin,cout,stride,h,w', [(768, 128, 1, 180, 180), (128, 128, 1, 180, 180), (128, 256, 2, 180, 180), (256, 256, 1, 90, 90)])
def test_conv3x3_bf16(bf16_mode, cin, cout, stride, h, w):
    from isfusion_b200 import conv2d
    torch.manual_seed(cin + cout)
    x = torch.randn(1, cin, h, w).to(BF)
    wt = (torch.randn(cout, cin, 3, 3) / (9 * cin) ** 0.5).to(BF)
    b = torch.randn(cout) * 0.1
    ref = F.relu(F.conv2d(x.float(), wt.float(), b, stride=stride, padding=1))
    xg = x.cuda().permute(0, 2, 3, 1).contiguous()
    out = conv2d.conv3x3_tma(xg, (wt.cuda(), b.to(BF).cuda(), b.cuda()), stride, True)
    assert out is not None and out.dtype == BF
    r, mx = _rel(out.permute(0, 3, 1, 2), ref)
    assert mx <= BF_OUT * max(1.0, ref.abs().max().item()) and r <= BF_OUT, (r, mx)


@pytest.mark.parametrize('g,c', [(180, 128), (90, 256), (21, 128)])
def test_window_attention_bf16_vs_f16_kernel(g, c):
    """the bf16 kernel against the fp16 kernel on the same values (every bf16 number is an fp16 number in this range)."""
    from isfusion_b200 import _lib as L
    torch.manual_seed(g)
    qkv = (torch.randn(g * g, 3 * c) * 0.7).to(BF)
    posq, posk = (torch.randn(36, c) * 0.3).cuda(), (torch.randn(36, c) * 0.3).cuda()
    for shifted in (0, 1):
        o16 = torch.empty(g * g, c, dtype=torch.float16, device='cuda')
        obf = torch.empty(g * g, c, dtype=BF, device='cuda')
        q16, qbf = qkv.half().cuda(), qkv.cuda()
        L.check(L.lib().isf_window_attention(L.ptr(q16), L.ptr(posq), L.ptr(posk), 1, g, c, 8, shifted, L.ptr(o16), L.stream()))
        L.check(L.lib().isf_window_attention_bf16(L.ptr(qbf), L.ptr(posq), L.ptr(posk), 1, g, c, 8, shifted, L.ptr(obf), L.stream()))
        torch.cuda.synchronize()
        r, mx = _rel(obf, o16)
        # K + pos and the probabilities are re-rounded to bf16 inside the kernel (2^-9 each), the output once more
        assert r <= 2 * BF_OUT and mx <= 4 * BF_OUT * max(1.0, o16.float().abs().max().item()), (shifted, r, mx)


def test_cam2bev_bf16(bf16_mode):
    from isfusion_b200.fusion_encoder import ISFusionEncoder
    from oracle import voxel_ops as vo
    cfg = omodel.neck_config()['fusion_encoder']
    pts = synth.make_points(20000, 'S', 3)
    l2i, aug, laug = synth.make_calib(1)
    feats = synth.make_img_feats(1)
    feats[1] = feats[1].to(BF).float()
    v, c, n = vo.hard_voxelize(pts, [0.6, 0.6, 8.0], synth.PC_RANGE, 12, 60000)
    pm = dict(pillars=v, pillar_coors=torch.nn.functional.pad(c, (1, 0), value=0), pillars_num_points=n)
    kw = dict(img_metas=[dict(input_shape=(256, 704))], lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    ref = fo.ISFusionEncoder(**cfg).img_fv_to_bev([feats[1]], 1, pts_metas=pm, **kw)
    enc = ISFusionEncoder(**cfg).cuda()
    out = enc.img_fv_to_bev([feats[1].cuda()], 1, pts_metas={k: t.cuda() for k, t in pm.items()}, **kw)
    assert out.dtype == BF
    r, mx = _rel(out.permute(0, 3, 1, 2), ref)
    assert r <= BF_OUT and mx <= BF_OUT * max(1.0, ref.abs().max().item()), (r, mx)


@pytest.mark.parametrize('g,c', [(180, 128), (90, 256)])
def test_grid_to_region_level_bf16_against_the_oracle(bf16_mode, g, c):
    """SSTInputLayerV2 + SSTv2 (one shifted-window block = 2 encoder layers, 10 GEMMs with bf16 outputs, 2 attention cores, 4
    LayerNorms) in bf16 against the fp32 oracle: ~10 stored bf16 tensors on the path, 2^-9 each -> a few 1e-3 of relative L2."""
    from isfusion_b200.fusion_encoder import SSTInputLayerV2, SSTv2
    torch.manual_seed(g + c)
    di = {0: {'max_tokens': 36, 'drop_range': (0, 100000)}}
    rl = ss.SSTInputLayerV2(di, (6, 6, 1), [g, g, 1], shuffle_voxels=True, pos_temperature=1000, pos_embed=c)
    rn = ss.SSTv2([c] * 4, [8] * 4, 1, [c] * 4, output_shape=[g, g], in_channel=128 if c == 128 else None).eval()
    ys, xs = torch.meshgrid(torch.arange(g), torch.arange(g), indexing='ij')
    coords = torch.stack([torch.zeros(g * g, dtype=torch.long), torch.zeros(g * g, dtype=torch.long), ys.flatten(), xs.flatten()], 1).int()
    feats = torch.randn(g * g, 128 if c == 128 else c).to(BF).float()
    ref = rn(rl(feats, coords, 1))[0]
    gl = SSTInputLayerV2(di, (6, 6, 1), [g, g, 1], pos_temperature=1000, pos_embed=c)
    gn = SSTv2([c] * 4, [8] * 4, 1, [c] * 4, output_shape=[g, g], in_channel=128 if c == 128 else None)
    gn.load_state_dict(rn.state_dict())
    gn = gn.cuda().eval()
    out = gn(gl(feats.to(BF).cuda(), None, 1))[0]
    assert out.dtype == BF and out.shape == ref.shape
    r, mx = _rel(out, ref)
    print('bf16 grid-to-region level %d x %d: rel-L2 %.2e max %.2e' % (g, g, r, mx))
    assert r < 1.2e-2, (r, mx)       # weights (2^-9), 10 bf16 activations, bf16 probabilities; the fp16 path asserts 3e-3 here
