"""tcgen05 GEMM vs a plain PyTorch fp32 reference of the same op (floating-point kernel => torch reference)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, scale, shift, res, ln, act):
    y = a.float() @ w.float().t()
    if scale is not None:
        y = y * scale
    if shift is not None:
        y = y + shift
    if res is not None:
        y = y + res.float()
    if ln is not None:
        y = torch.nn.functional.layer_norm(y, (y.shape[1],), ln[0], ln[1], ln[2])
    if act == 'relu':
        y = torch.relu(y)
    elif act == 'gelu':
        y = torch.nn.functional.gelu(y)
    return y


@pytest.mark.parametrize('m,k,n', [(1, 64, 32), (127, 64, 64), (128, 128, 128), (1000, 256, 256), (32400, 128, 384),
                                   (8100, 256, 768), (300, 32, 64), (5000, 72, 128)])
def test_gemm_plain(m, k, n):
    from isfusion_b200.gemm import gemm, pack_weight
    torch.manual_seed(m + k + n)
    a = (torch.randn(m, k, device='cuda') * 0.5).half()
    w = (torch.randn(n, k, device='cuda') / k ** 0.5).half()
    out = gemm(a, pack_weight(w), k=k)
    ref = _ref(a, w, None, None, None, None, None)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err     # fp16 output rounding of an fp32 accumulation


@pytest.mark.parametrize('n', [128, 256])
@pytest.mark.parametrize('act', [None, 'relu', 'gelu'])
def test_gemm_epilogues(n, act):
    from isfusion_b200.gemm import gemm, pack_weight
    torch.manual_seed(n)
    m, k = 3333, n
    a = (torch.randn(m, k, device='cuda') * 0.5).half()
    w = (torch.randn(n, k, device='cuda') / k ** 0.5).half()
    scale = torch.rand(n, device='cuda') + 0.5
    shift = torch.randn(n, device='cuda') * 0.1
    res = torch.randn(m, n, device='cuda').half()
    gamma, beta = torch.rand(n, device='cuda') + 0.5, torch.randn(n, device='cuda') * 0.1
    wp = pack_weight(w)
    for use_ln in (False, True):
        ln = (gamma, beta, 1e-5) if use_ln else None
        out = gemm(a, wp, scale=scale, shift=shift, residual=res, act=act, ln=ln)
        ref = _ref(a, w, scale, shift, res, ln, act)
        err = (out.float() - ref).abs().max().item()
        assert err <= 2e-3 * max(1.0, ref.abs().max().item()), (use_ln, err)
    out32 = gemm(a, wp, shift=shift, out_dtype=torch.float32)
    assert (out32 - _ref(a, w, None, shift, None, None, None)).abs().max().item() < 1e-3


def test_gemm_scatter_rows():
    from isfusion_b200.gemm import gemm, pack_weight
    torch.manual_seed(0)
    m, k, n = 2000, 128, 128
    a = torch.randn(m, k, device='cuda').half()
    w = (torch.randn(n, k, device='cuda') / k ** 0.5).half()
    perm = torch.randperm(m, device='cuda').int()
    perm[::7] = -1
    out = torch.zeros(m, n, device='cuda', dtype=torch.float16)
    gemm(a, pack_weight(w), out=out, out_rows=perm)
    ref = _ref(a, w, None, None, None, None, None)
    keep = perm >= 0
    exp = torch.zeros_like(ref)
    exp[perm[keep].long()] = ref[keep]
    assert (out.float() - exp).abs().max().item() < 2e-2


@pytest.mark.parametrize('cs', [2])
def test_gemm_cluster_multicast(cs):
    """thread-block-cluster variant (TMA-multicast weights) gives the same bits as the single-CTA variant."""
    from isfusion_b200 import _lib as L
    from isfusion_b200.gemm import gemm, pack_weight
    torch.manual_seed(cs)
    try:
        for (m, k, n) in [(5000, 128, 128), (129, 256, 256), (32400, 128, 384), (700, 64, 64), (1000, 64, 32)]:
            a = (torch.randn(m, k, device='cuda') * 0.5).half()
            w = (torch.randn(n, k, device='cuda') / k ** 0.5).half()
            res = torch.randn(m, n, device='cuda').half()
            shift = torch.randn(n, device='cuda') * 0.1
            L.set_option('gemm_cluster', 1)
            ref = gemm(a, pack_weight(w), shift=shift, residual=res, act='relu')
            L.set_option('gemm_cluster', cs)
            out = gemm(a, pack_weight(w), shift=shift, residual=res, act='relu')
            torch.cuda.synchronize()
            assert torch.equal(out, ref), (m, k, n)
    finally:
        L.set_option('gemm_cluster', 1)
