"""CPU tests: the oracle's voxel ops against the reference's own golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import voxel_ops as vo

MAIN = ([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3])
PILLAR = ([0.6, 0.6, 8.0], [-54, -54, -5, 54, 54, 3])


def test_voxel_generator_known_answer(golden_dir):
    """ref: tests/test_models/test_voxel_encoder/test_voxel_generator.py:7-23 (expected values verbatim)."""
    g = np.load(os.path.join(golden_dir, 'voxel_generator_kat.npz'))
    expected_coors = np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1], [7, 81, 0], [6, 80, 1],
                               [7, 80, 0], [6, 80, 0]])
    expected_npv = np.array([120, 121, 127, 134, 115, 127, 125, 131])
    assert np.all(g['coors'] == expected_coors) and np.all(g['npv'] == expected_npv)
    v, c, n = vo.hard_voxelize(torch.from_numpy(g['points']), [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    assert np.all(c.numpy() == expected_coors)
    assert np.all(n.numpy() == expected_npv)
    assert v.shape == (8, 1000, 4)
    assert np.array_equal(v.numpy(), g['voxels'])


def test_dynamic_voxelize_vs_reference_cpp(golden_dir):
    g = np.load(os.path.join(golden_dir, 'voxel_ref_cpp.npz'))
    pts = torch.from_numpy(g['points'])
    for name, (v, r) in {'main': MAIN, 'pillar': PILLAR}.items():
        ours = vo.dynamic_voxelize(pts, v, r, mode='cpu').numpy()
        assert np.array_equal(ours, g['dyn_' + name]), name
        # the CUDA encoding differs only on invalid rows and agrees on validity
        cu = vo.dynamic_voxelize(pts, v, r, mode='cuda').numpy()
        ok = (ours >= 0).all(1)
        assert np.array_equal(cu[ok], ours[ok])
        assert ((cu[~ok] < 0).any(1)).all()


@pytest.mark.parametrize('tag,mp,mv', [('p12_v60000', 12, 60000), ('p12_v300', 12, 300), ('p3_v60000', 3, 60000)])
def test_hard_voxelize_vs_reference_cpp(golden_dir, tag, mp, mv):
    g = np.load(os.path.join(golden_dir, 'voxel_ref_cpp.npz'))
    pts = torch.from_numpy(g['points'])
    for fn in (vo.hard_voxelize, vo.hard_voxelize_loop):
        v, c, n = fn(pts, *PILLAR, mp, mv)
        assert np.array_equal(v.numpy(), g['hard_%s_voxels' % tag])
        assert np.array_equal(c.numpy(), g['hard_%s_coors' % tag])
        assert np.array_equal(n.numpy(), g['hard_%s_npv' % tag])


def test_live_reference_cpp_if_built():
    """When oracle/_ref is present, compare on a fresh larger cloud too."""
    from oracle import ref_voxel
    if not ref_voxel.available():
        pytest.skip('oracle/_ref not built')
    g = torch.Generator().manual_seed(7)
    pts = torch.rand(50000, 5, generator=g) * torch.tensor([110., 110., 8.2, 1., 1.]) + torch.tensor([-55., -55., -5.1, 0., 0.])
    assert torch.equal(vo.dynamic_voxelize(pts, *MAIN, mode='cpu'), ref_voxel.dynamic_voxelize(pts, *MAIN))
    a = vo.hard_voxelize(pts, *PILLAR, 12, 30000)
    b = ref_voxel.hard_voxelize(pts, *PILLAR, 12, 30000)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_dynamic_scatter_bruteforce_contract():
    """ref: tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:9-72 (brute-force definition)."""
    g = torch.Generator().manual_seed(0)
    feats = torch.rand(20000, 3, generator=g) * 100 - 50
    coors = torch.randint(-1, 20, (20000, 3), generator=g, dtype=torch.int32)
    for red in ('mean', 'max', 'sum'):
        f, c, m, cnt = vo.dynamic_scatter(feats, coors, red)
        valid = (coors >= 0).all(1)
        uc = torch.unique(coors[valid], dim=0)
        assert torch.equal(c, uc)
        for k in [0, 17, len(uc) // 2, len(uc) - 1]:
            sel = (coors == uc[k]).all(1)
            ref = {'mean': feats[sel].mean(0), 'max': feats[sel].max(0).values, 'sum': feats[sel].sum(0)}[red]
            assert torch.allclose(f[k], ref, atol=1e-2)
            assert cnt[k] == sel.sum()
            assert (m[sel] == k).all()
        assert (m[~valid] == -1).all()
    # empty input and all-invalid input
    e = vo.dynamic_scatter(torch.zeros(0, 3), torch.zeros(0, 3, dtype=torch.int32), 'max')
    assert e[0].shape[0] == 0
    f, c, m, cnt = vo.dynamic_scatter(torch.rand(10, 3), -torch.ones(10, 3, dtype=torch.int32), 'max')
    assert f.shape[0] == 0 and (m == -1).all()


def test_vfe_canvas_equivalence():
    """map_voxel_center_to_point: sorted-search restatement == the reference's literal dense canvas
    (voxel_encoder.py:410-450) on a small grid."""
    torch.manual_seed(0)
    vs, pr = [0.5, 0.5, 1.0], [0, 0, 0, 8, 8, 2]
    vfe = vo.DynamicVFE(5, [64, 64], False, True, True, vs, pr, dict(type='BN1d', eps=1e-3, momentum=0.01)).eval()
    pts = torch.rand(500, 5) * torch.tensor([8., 8., 2., 1., 1.])
    c3 = vo.dynamic_voxelize(pts, vs, pr)
    coors = torch.nn.functional.pad(c3, (1, 0), value=0)
    coors[250:, 0] = 1
    vm, vc = vfe.cluster_scatter(pts, coors)
    a = vfe.map_voxel_center_to_point(coors, vm, vc)
    cz, cy, cx = 2, 16, 16
    canvas = torch.zeros(cz * cy * cx * 2, dtype=torch.long)
    lin = lambda c: (c[:, 0] * cz * cy * cx + c[:, 1] * cy * cx + c[:, 2] * cx + c[:, 3]).long()
    canvas[lin(vc)] = torch.arange(vm.size(0))
    b = vm[canvas[lin(coors)]]
    assert torch.equal(a, b)
    out, oc = vfe(pts, coors)
    assert out.shape == (vc.shape[0], 64) and torch.equal(oc, vc)
