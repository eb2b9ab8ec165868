"""GPU parity: CUDA voxel ops (through the C-ABI) vs the oracle and the reference golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import voxel_ops as vo

pytestmark = pytest.mark.gpu

MAIN = ([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3])
PILLAR = ([0.6, 0.6, 8.0], [-54, -54, -5, 54, 54, 3])


def _cloud(n, seed, margin=1.0):
    g = torch.Generator().manual_seed(seed)
    span = torch.tensor([108. + 2 * margin, 108. + 2 * margin, 8. + 0.2 * margin, 1., 0.5])
    lo = torch.tensor([-54. - margin, -54. - margin, -5. - 0.1 * margin, 0., 0.])
    return torch.rand(n, 5, generator=g) * span + lo


def test_dynamic_voxelize_bit_exact_golden(golden_dir):
    from isfusion_b200.voxel_layer import Voxelization
    g = np.load(os.path.join(golden_dir, 'voxel_ref_cpp.npz'))
    pts = torch.from_numpy(g['points']).cuda()
    for name, (v, r) in {'main': MAIN, 'pillar': PILLAR}.items():
        out = Voxelization(v, r, -1, (-1, -1))(pts).cpu().numpy()
        ref = g['dyn_' + name]
        ok = (ref >= 0).all(1)
        assert np.array_equal(out[ok], ref[ok])                 # voxel indices bit-exact (incl. 1-ulp faces)
        assert np.array_equal(out, vo.dynamic_voxelize(pts.cpu(), v, r, mode='cuda').numpy())


@pytest.mark.parametrize('n', [0, 1, 300000, 1000000])
def test_dynamic_voxelize_sizes(n):
    from isfusion_b200.voxel_layer import Voxelization
    pts = _cloud(n, 1)
    out = Voxelization(*MAIN, -1, (-1, -1))(pts.cuda())
    assert torch.equal(out.cpu(), vo.dynamic_voxelize(pts, *MAIN, mode='cuda'))


@pytest.mark.parametrize('tag,mp,mv', [('p12_v60000', 12, 60000), ('p12_v300', 12, 300), ('p3_v60000', 3, 60000)])
def test_hard_voxelize_golden(golden_dir, tag, mp, mv):
    from isfusion_b200.voxel_layer import voxelization
    g = np.load(os.path.join(golden_dir, 'voxel_ref_cpp.npz'))
    v, c, n = voxelization(torch.from_numpy(g['points']).cuda(), *PILLAR, mp, mv)
    assert np.array_equal(v.cpu().numpy(), g['hard_%s_voxels' % tag])
    assert np.array_equal(c.cpu().numpy(), g['hard_%s_coors' % tag])
    assert np.array_equal(n.cpu().numpy(), g['hard_%s_npv' % tag])


def test_hard_voxelize_full_size_and_known_answer(golden_dir):
    from isfusion_b200.voxel_layer import voxelization
    pts = _cloud(300000, 2)
    for mv in (60000, 30000):
        a = voxelization(pts.cuda(), *PILLAR, 12, mv)
        b = vo.hard_voxelize(pts, *PILLAR, 12, mv)
        assert all(torch.equal(x.cpu(), y) for x, y in zip(a, b))
    g = np.load(os.path.join(golden_dir, 'voxel_generator_kat.npz'))
    v, c, n = voxelization(torch.from_numpy(g['points']).cuda(), [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    assert np.array_equal(c.cpu().numpy(), g['coors']) and np.array_equal(n.cpu().numpy(), g['npv'])
    assert np.array_equal(v.cpu().numpy(), g['voxels'])


@pytest.mark.parametrize('red', ['mean', 'max', 'sum'])
def test_dynamic_scatter_contract(red):
    """ref contract: tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:9-94."""
    from isfusion_b200.voxel_layer import dynamic_point_to_voxel_forward
    g = torch.Generator().manual_seed(0)
    feats = torch.rand(200000, 3, generator=g) * 100 - 50
    coors = torch.randint(-1, 20, (200000, 3), generator=g, dtype=torch.int32)
    f, c, m, cnt = dynamic_point_to_voxel_forward(feats.cuda(), coors.cuda(), red)
    rf, rc, rm, rcnt = vo.dynamic_scatter(feats, coors, red)
    assert torch.equal(c.cpu(), rc) and torch.equal(m.cpu(), rm) and torch.equal(cnt.cpu(), rcnt)
    if red == 'max':
        assert torch.equal(f.cpu(), rf)
    else:
        assert torch.allclose(f.cpu(), rf, rtol=1e-5, atol=1e-2)


def test_dynamic_scatter_edge_cases():
    from isfusion_b200.voxel_layer import DynamicScatter, dynamic_point_to_voxel_forward
    dsmax = DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], False)
    e = dsmax(torch.rand(0, 3).cuda(), torch.randint(-1, 20, (0, 3), dtype=torch.int32).cuda())
    assert e[0].shape == (0, 3) and e[1].shape == (0, 3)
    f, c, m, cnt = dynamic_point_to_voxel_forward(torch.rand(100, 4).cuda(), -torch.ones(100, 3, dtype=torch.int32).cuda(), 'max')
    assert f.shape[0] == 0 and c.shape[0] == 0 and (m == -1).all()
    # gradients: all-invalid -> zero grad; max routes to arg-max
    feats = torch.rand(500, 4, device='cuda', requires_grad=True)
    coors = torch.randint(-1, 4, (500, 3), dtype=torch.int32, device='cuda')
    out, _ = dsmax(feats, coors)
    out.sum().backward()
    valid = (coors >= 0).all(1)
    assert (feats.grad[~valid] == 0).all()
    assert abs(feats.grad.sum().item() - out.numel()) < 1e-3


def test_ingroup_indices():
    from isfusion_b200.voxel_layer import ingroup_indices
    g = torch.Generator().manual_seed(0)
    grp = torch.randint(0, 1861, (64800,), generator=g)
    out = ingroup_indices(grp.cuda()).cpu()
    # bijection onto 0..count-1 inside every group (sst_ops.py:150-175 sanity checks)
    cnt = torch.bincount(grp)
    assert (out >= 0).all() and (out < cnt[grp]).all()
    key = grp * 100000 + out
    assert key.unique().numel() == grp.numel()


def _make_vfe(seed=0):
    torch.manual_seed(seed)
    vfe = vo.DynamicVFE(5, [64, 64], False, True, True, MAIN[0], MAIN[1], dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01))
    for m in vfe.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    return vfe.eval()


@pytest.mark.parametrize('tensor_core', [True, False])
@pytest.mark.parametrize('n,batch', [(20000, 1), (300000, 2)])
def test_dynamic_vfe_parity(n, batch, tensor_core):
    from isfusion_b200.voxel_encoder import DynamicVFE
    ref = _make_vfe()
    pts = _cloud(n, 3, margin=-0.01)                       # strictly in range (Appendix C #5)
    c3 = vo.dynamic_voxelize(pts, *MAIN)
    coors = torch.nn.functional.pad(c3, (1, 0), value=0)
    coors[n // 2:, 0] = batch - 1
    rf, rc = ref(pts, coors)
    mod = DynamicVFE(5, [64, 64], False, True, True, MAIN[0], MAIN[1], dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01))
    mod.load_state_dict(ref.state_dict())
    mod = mod.cuda().eval()
    mod.tensor_core = tensor_core      # layer 2 on the tcgen05 GEMM (default) or the all-fp32 SIMT kernel
    f, c = mod(pts.cuda(), coors.cuda())
    assert torch.equal(c.cpu(), rc)                        # voxel coordinates bit-exact, same order
    err = (f.float().cpu() - rf).abs().max().item()
    rel_l2 = ((f.float().cpu() - rf).norm() / rf.norm()).item()
    print('vfe tensor_core=%s: max abs err %.3e (max ref %.2f), rel l2 %.2e' % (tensor_core, err, rf.abs().max().item(), rel_l2))
    assert rel_l2 <= 1e-3, rel_l2
    assert err <= 2e-3 * max(1.0, rf.abs().max().item()), err   # 1e-3 rel (fp16 output of an fp32 pipeline)


@pytest.mark.parametrize('own_bwd', [False, True])
def test_dynamic_vfe_train_mode_forward_backward(own_bwd):
    """TRAIN mode of the product DynamicVFE (batch-statistics naiveSyncBN1d, differentiable: scatter forward K1-K3 and backward
    K3b from the library, Linear / BN / ReLU through autograd) against the oracle module trained the same way on the CPU:
    outputs, voxel order, gradients of every parameter and the updated running statistics."""
    import isfusion_b200.voxel_layer as VL
    from isfusion_b200.voxel_encoder import DynamicVFE
    ref = _make_vfe().train()
    n = 6000
    pts = _cloud(n, 9, margin=-0.01)
    c3 = vo.dynamic_voxelize(pts, *MAIN)
    coors = torch.nn.functional.pad(c3, (1, 0), value=0)
    coors[n // 2:, 0] = 1
    mod = DynamicVFE(5, [64, 64], False, True, True, MAIN[0], MAIN[1], dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01))
    mod.load_state_dict(ref.state_dict())
    mod = mod.cuda().train()
    rf, rc = vo.dynamic_vfe_train_forward(ref, pts, coors)
    with torch.no_grad():
        ef, ec = ref.eval()(pts, coors)        # the differentiable restatement == the oracle's forward (eval statistics)
    ref.train()
    assert torch.equal(ec, rc)
    torch.manual_seed(3)
    wgt = torch.randn(rf.shape)
    (rf * wgt).sum().backward()
    old = VL.OWN_SCATTER_BACKWARD
    try:
        VL.OWN_SCATTER_BACKWARD = own_bwd
        f, c = mod(pts.cuda(), coors.cuda())
        assert f.requires_grad and f.dtype == torch.float32
        (f * wgt.cuda()).sum().backward()
        torch.cuda.synchronize()
    finally:
        VL.OWN_SCATTER_BACKWARD = old
    assert torch.equal(c.cpu(), rc)
    assert torch.allclose(f.detach().cpu(), rf.detach(), rtol=1e-4, atol=1e-4)
    for (name, p), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, name
        r = ((p.grad.cpu() - q.grad).norm() / q.grad.norm().clamp(min=1e-12)).item()
        assert r < 1e-3, (name, r)
    for (name, b), (_, q) in zip(mod.named_buffers(), ref.named_buffers()):
        assert torch.allclose(b.cpu().float(), q.float(), rtol=1e-4, atol=1e-5), name
    # eval after train still takes the fused kernels (packed weights were invalidated by train())
    mod.eval()
    fe, ce = mod(pts.cuda(), coors.cuda())
    assert fe.dtype == torch.float16 and torch.equal(ce.cpu(), rc)
