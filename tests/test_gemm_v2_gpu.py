"""Bring-up tests of the OPT-IN warp-specialised GEMM schedule (csrc/tc_gemm2.cuh, isf_set_option("gemm_v2")).

Not part of the default GPU suite: the kernel was written at the end of round 1 without GPU time left to run it.
Enable with ISF_EXPERIMENTAL=1 (run each under `timeout`: the kernel's barrier waits are bounded and trap, but a
first run of new tcgen05 code deserves a belt as well).  The schedule changes nothing in the arithmetic order
(same chunk order into the same accumulator, same epilogue), so the results must be bit-identical to the default
kernel."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('ISF_EXPERIMENTAL') != '1', reason='opt-in kernel: set ISF_EXPERIMENTAL=1')]


@pytest.mark.parametrize('m,k,n', [(1, 64, 32), (700, 64, 64), (5000, 128, 128), (129, 256, 256), (40000, 128, 384)])
def test_v2_dense_gemm_bit_identical(m, k, n):
    from isfusion_b200 import _lib as L
    from isfusion_b200.gemm import gemm, pack_weight
    torch.manual_seed(m)
    a = (torch.randn(m, k, device='cuda') * 0.5).half()
    w = (torch.randn(n, k, device='cuda') / k ** 0.5).half()
    res = torch.randn(m, n, device='cuda').half()
    shift = torch.randn(n, device='cuda') * 0.1
    gamma, beta = torch.rand(n, device='cuda') + 0.5, torch.randn(n, device='cuda') * 0.1
    wp = pack_weight(w)
    try:
        for ln in (None, (gamma, beta, 1e-5)) if n in (128, 256) else (None,):
            L.set_option('gemm_v2', 0)
            ref = gemm(a, wp, shift=shift, residual=res, act='relu', ln=ln)
            L.set_option('gemm_v2', 2)
            for io in (0, 1):           # 16-byte / 32-byte (LDG/STG.E.ENL2.256) epilogue accesses
                L.set_option('gemm_v2_io256', io)
                out = gemm(a, wp, shift=shift, residual=res, act='relu', ln=ln)
                torch.cuda.synchronize()
                assert torch.equal(out, ref), (m, k, n, ln is not None, io)
            # a column-offset view whose rows are NOT 32-byte aligned must fall back to 16-byte stores by itself
            wide = torch.zeros(m, n + 8, device='cuda', dtype=torch.float16)
            gemm(a, wp, shift=shift, residual=res, act='relu', ln=ln, out=wide[:, 8:])
            torch.cuda.synchronize()
            assert torch.equal(wide[:, 8:], ref) and not wide[:, :8].any()
    finally:
        L.set_option('gemm_v2', 0)
        L.set_option('gemm_v2_io256', 1)


@pytest.mark.parametrize('zskip', [0, 1])
def test_v2_sparse_encoder_bit_identical(zskip):
    from isfusion_b200 import _lib as L
    from isfusion_b200.sparse_encoder import SparseEncoder
    from test_spconv_gpu import _rand_tensor
    torch.manual_seed(0)
    cfg = dict(in_channels=64, sparse_shape=[41, 128, 128], base_channels=32, output_channels=256,
               encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')
    mod = SparseEncoder(**cfg).cuda().eval()
    x = _rand_tensor([41, 128, 128], 20000, 64, 2, 11, clustered=True)
    f, c = torch.relu(x.features).half().cuda(), x.indices.cuda()
    try:
        L.set_option('gemm_v2', 0)
        ref, _, _ = mod(f, c, 2)
        L.set_option('gemm_v2', 1)
        L.set_option('gemm_v2_zskip', zskip)
        out, _, _ = mod(f, c, 2)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
    finally:
        L.set_option('gemm_v2', 0)
        L.set_option('gemm_v2_zskip', 1)
