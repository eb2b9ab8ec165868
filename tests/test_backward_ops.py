"""Backward kernels of the two custom operators with a gradient (SURVEY §8b operator table, §8f rank 4: K3b dynamic
scatter backward, K9b multi-scale deformable attention backward).

CPU (default suite):
  * the kernel's own per-sample arithmetic — shared __host__ __device__ source, compiled for the host
    (isf_debug_msda_backward_host) — against gradients of the REFERENCE's ms_deform_attn_core_pytorch taken by autograd in
    float64 (tests/golden/msda_bwd_ref.npz, made by make_golden.py from the reference file; what the reference's own
    gradient test ops/test.py:63-86 holds its CUDA backward against), and the oracle's restatement against the same fixture;
  * the oracle's scatter backward (loop restatement of scatter_points_cuda.cu:105-179) against autograd of the brute-force
    definition (sum / mean) and the lowest-index tie rule (max);
  * the product's autograd wiring (voxel_layer._dynamic_scatter, fusion_encoder.MultiScaleDeformableAttnFunction) on the
    ABI emulator.
GPU: the kernels against the same."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle import fusion_ops as fo     # noqa: E402
from oracle import voxel_ops as vo      # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'msda_bwd_ref.npz'))
CASES = ['a', 'b', 'c']


def _case(k):
    t = lambda n: torch.from_numpy(G[k + '_' + n])
    return t('value'), t('shapes'), t('loc'), t('aw'), t('grad_out'), t('grad_value'), t('grad_loc'), t('grad_aw')


def _lvl_start(shapes):
    return torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])


def _close(a, b, tol):
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize('k', CASES)
def test_oracle_msda_gradients_match_the_reference_function(k):
    value, shapes, loc, aw, go, gv, gl, ga = _case(k)
    v, l, a = (t.clone().double().requires_grad_(True) for t in (value, loc, aw))
    y = fo.ms_deform_attn_core_pytorch(v, shapes, l, a)
    assert _close(y.detach().float(), torch.from_numpy(G[k + '_out']), 1e-6)
    g = torch.autograd.grad(y, (v, l, a), go.double())
    for got, want in zip(g, (gv, gl, ga)):
        assert _close(got.float(), want, 1e-6)


@pytest.mark.parametrize('k', CASES)
def test_msda_backward_kernel_arithmetic_on_host(k):
    """msda_corners / msda_bwd_channel (the device functions of msda_backward_kernel) compiled for the host."""
    from isfusion_b200 import _lib as L
    value, shapes, loc, aw, go, gv, gl, ga = _case(k)
    b, s, h, d = value.shape
    _, q, _, lv, p, _ = loc.shape
    arrs = [value.numpy().copy(), shapes.numpy().astype(np.int64), _lvl_start(shapes).numpy().astype(np.int64), loc.numpy().copy(),
            aw.numpy().copy(), go.numpy().copy()]
    grad_value = np.zeros(value.shape, np.float32)
    grad_loc = np.full(loc.shape, 7.0, np.float32)         # overwritten, not accumulated
    grad_aw = np.full(aw.shape, 7.0, np.float32)
    ptr = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    L.check(L.lib().isf_debug_msda_backward_host(*[ptr(x) for x in arrs], b, s, h, d, lv, q, p, ptr(grad_value), ptr(grad_loc),
                                                 ptr(grad_aw)))
    assert _close(torch.from_numpy(grad_value), gv, 2e-5)
    assert _close(torch.from_numpy(grad_loc), gl, 2e-5)
    assert _close(torch.from_numpy(grad_aw), ga, 2e-5)
    # grad_value is accumulated into (the caller zeroes it): a second call doubles it
    L.check(L.lib().isf_debug_msda_backward_host(*[ptr(x) for x in arrs], b, s, h, d, lv, q, p, ptr(grad_value), ptr(grad_loc),
                                                 ptr(grad_aw)))
    assert _close(torch.from_numpy(grad_value), 2 * gv, 4e-5) and _close(torch.from_numpy(grad_loc), gl, 2e-5)


def _scatter_case(seed, n=400, c=3, ties=True):
    g = torch.Generator().manual_seed(seed)
    feats = torch.rand(n, c, generator=g) * 10 - 5
    if ties:                                                  # repeated values inside a voxel: the max has several arg-maxes
        feats = (feats * 2).round() / 2
    coors = torch.randint(-1, 4, (n, 3), generator=g, dtype=torch.int32)
    return feats, coors


@pytest.mark.parametrize('red', ['sum', 'mean', 'max'])
def test_oracle_scatter_backward_against_the_brute_force_definition(red):
    feats, coors = _scatter_case(0)
    rf, rc, cmap, cnt = vo.dynamic_scatter(feats, coors, red)
    g = torch.randn(rf.shape, generator=torch.Generator().manual_seed(1))
    got = vo.dynamic_scatter_backward(g, feats, rf, cmap, cnt, red)
    # brute force over voxels (tests/.../test_dynamic_scatter.py:57-66 is the forward definition)
    want = torch.zeros_like(feats)
    for v in range(rf.shape[0]):
        members = torch.nonzero(cmap == v).flatten()
        if red == 'sum':
            want[members] = g[v]
        elif red == 'mean':
            want[members] = g[v] / len(members)
        else:
            for i in range(feats.shape[1]):
                first = members[(feats[members, i] == rf[v, i])].min()      # lowest index among the arg-maxes
                want[first, i] = g[v, i]
    assert torch.equal(got, want)
    assert (got[cmap < 0] == 0).all()
    if red != 'max':        # and autograd of the definition
        f = feats.clone().requires_grad_(True)
        valid = cmap >= 0
        y = torch.zeros(rf.shape).index_add(0, cmap[valid].long(), f[valid])
        if red == 'mean':
            y = y / cnt.unsqueeze(-1)
        y.backward(g)
        assert torch.allclose(f.grad, got, atol=1e-6)


@pytest.mark.parametrize('red', ['sum', 'mean', 'max'])
def test_product_scatter_autograd_wiring_on_the_emulator(red, monkeypatch):
    import abi_emulator
    import isfusion_b200.voxel_layer as VL
    abi_emulator.install(monkeypatch)
    feats, coors = _scatter_case(2)
    outs = {}
    for own in (False, True):          # device-side torch index ops (default) and the K3b entry point give the same gradient
        monkeypatch.setattr(VL, 'OWN_SCATTER_BACKWARD', own)
        f = feats.clone().requires_grad_(True)
        vf, vc = VL.dynamic_scatter(f, coors, red)
        g = torch.randn(vf.shape, generator=torch.Generator().manual_seed(3))
        vf.backward(g)
        outs[own] = f.grad.clone()
    rf, rc, cmap, cnt = vo.dynamic_scatter(feats, coors, red)
    want = vo.dynamic_scatter_backward(g, feats, rf, cmap, cnt, red)
    assert torch.allclose(outs[True], want, atol=1e-6) and torch.allclose(outs[False], want, atol=1e-6)


def test_product_msda_autograd_wiring_on_the_emulator(monkeypatch):
    import abi_emulator
    import isfusion_b200.fusion_encoder as FE
    abi_emulator.install(monkeypatch)
    value, shapes, loc, aw, go, gv, gl, ga = _case('c')
    v, l, a = (t.clone().requires_grad_(True) for t in (value, loc, aw))
    y = FE.ms_deform_attn(v, shapes, _lvl_start(shapes), l, a)
    assert y.grad_fn is not None
    y.backward(go)
    assert _close(v.grad, gv, 1e-5) and _close(l.grad, gl, 1e-5) and _close(a.grad, ga, 1e-5)
    with torch.no_grad():                     # inference path: plain forward call, no autograd node
        y2 = FE.ms_deform_attn(v, shapes, _lvl_start(shapes), l, a)
    assert y2.grad_fn is None and torch.equal(y2, y.detach())


# ---------------------------------------------------------------------------------------------- GPU (bring-up)
@pytest.mark.gpu
@pytest.mark.parametrize('k', CASES)
@pytest.mark.parametrize('half', [False, True])
def test_msda_backward_kernel(k, half):
    import isfusion_b200.fusion_encoder as FE
    value, shapes, loc, aw, go, gv, gl, ga = _case(k)
    if half:
        value = value.half().float()
        v64, l64, a64 = (t.double().requires_grad_(True) for t in (value, loc, aw))
        gv, gl, ga = (t.float() for t in torch.autograd.grad(fo.ms_deform_attn_core_pytorch(v64, shapes, l64, a64), (v64, l64, a64), go.double()))
    v = (value.half() if half else value).cuda().requires_grad_(True)
    l, a = loc.cuda().requires_grad_(True), aw.cuda().requires_grad_(True)
    y = FE.ms_deform_attn(v, shapes.cuda(), _lvl_start(shapes).cuda(), l, a)
    y.backward(go.cuda())
    assert _close(v.grad.float().cpu(), gv, 2e-3 if half else 2e-5)
    assert _close(l.grad.cpu(), gl, 2e-5) and _close(a.grad.cpu(), ga, 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('red', ['sum', 'mean', 'max'])
@pytest.mark.parametrize('n,c', [(400, 3), (200000, 5), (50000, 64)])
def test_scatter_backward_kernels(red, n, c, monkeypatch):
    import isfusion_b200.voxel_layer as VL
    feats, coors = _scatter_case(4, n, c)
    g = None
    grads = {}
    for own in (False, True):
        monkeypatch.setattr(VL, 'OWN_SCATTER_BACKWARD', own)
        f = feats.cuda().requires_grad_(True)
        vf, _ = VL.dynamic_scatter(f, coors.cuda(), red)
        if g is None:
            g = torch.randn(vf.shape, generator=torch.Generator().manual_seed(5)).cuda()
        vf.backward(g)
        grads[own] = f.grad.cpu()
    assert torch.equal(grads[True], grads[False])             # same arithmetic (one division), same tie rule
    if n <= 1000:
        rf, rc, cmap, cnt = vo.dynamic_scatter(feats, coors, red)
        assert torch.allclose(grads[True], vo.dynamic_scatter_backward(g.cpu(), feats, rf, cmap, cnt, red), atol=1e-6)
