"""Size-independent properties of the hot-path operators (no oracle in the assertion): what holds for ANY input, so it
can be checked at BASELINE's full sizes where the CPU oracle would take minutes.

  * sparse conv: the result does not depend on the storage order of the voxels (sorted rows -> bitmap-rank rulebook,
    shuffled rows -> hash rulebook: the same features row for row, bit-exact); an identity kernel returns its input;
    scaling the input by a power of two scales the output exactly; the output sites of a strided conv are the sorted
    unique set of candidate sites computed independently with torch ops.
  * dynamic scatter: counts add up to the valid points, out_coors strictly ascending, mean * count == sum, max >= mean.
  * heat-map top-K: scores descending, indices inside the map.
  * deformable attention: a constant value map sampled strictly inside it with weights summing to one returns the constant.

The same property functions run (a) on the CPU at small sizes against the ABI emulator — which proves the TEST code, and the
product's host code, before any GPU time is spent — and (b) on the GPU at the full-frame sizes."""
import ctypes
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _voxels(n_points, shape, voxel, lo, seed, dev):
    """sorted unique (b, z, y, x) voxel coordinates of a LiDAR-shaped synthetic cloud, by torch ops only."""
    import synth
    pts = synth.make_points(n_points, 'S', seed)[:, :3]
    c = ((pts - torch.tensor(lo)) / torch.tensor(voxel)).floor().long()
    ok = ((c >= 0) & (c < torch.tensor([shape[2], shape[1], shape[0]]))).all(1)
    c = c[ok]
    key = torch.unique((c[:, 2] * shape[1] + c[:, 1]) * shape[2] + c[:, 0])           # sorted
    idx = torch.stack([torch.zeros_like(key), key // (shape[1] * shape[2]), (key // shape[2]) % shape[1], key % shape[2]], 1)
    return idx.int().to(dev)


def _tensor(idx, feats, shape, sorted_rows):
    from isfusion_b200.spconv import SparseConvTensor
    t = SparseConvTensor(feats, idx, shape, 1)
    t.sorted = sorted_rows
    return t


def prop_sparse_conv(dev, idx, shape, cin, cout, seed=0):
    import isfusion_b200.spconv as sp
    g = torch.Generator().manual_seed(seed)
    v = idx.shape[0]
    feats = torch.randn(v, cin, generator=g).half().to(dev)
    perm = torch.randperm(v, generator=g).to(dev)
    # --- SubM: storage-order invariance, identity kernel, exact scaling
    conv = sp.SubMConv3d(cin, cout, 3, padding=1, bias=False)
    conv.weight.data = (torch.randn(conv.weight.shape, generator=g) / (27 * cin) ** 0.5 * 3).half().float()
    conv = conv.to(dev).eval()
    with torch.no_grad():
        ys = conv(_tensor(idx, feats, shape, True)).features
        yu = conv(_tensor(idx[perm].contiguous(), feats[perm].contiguous(), shape, False)).features
        assert torch.equal(yu, ys[perm]), 'SubM conv depends on the storage order of the voxels'
        y4 = conv(_tensor(idx, (feats * 4).contiguous(), shape, True)).features
        big = ys.float().abs() >= 1e-3                          # away from fp16 subnormals the scaling is exact
        assert torch.equal(y4[big].float(), ys[big].float() * 4)
        if cin == cout:
            ident = sp.SubMConv3d(cin, cout, 3, padding=1, bias=False)
            ident.weight.data.zero_()
            ident.weight.data[1, 1, 1] = torch.eye(cin)
            ident = ident.to(dev).eval()
            assert torch.equal(ident(_tensor(idx, feats, shape, True)).features, feats)
        # --- strided conv: output sites = sorted unique candidates (independent torch computation), order invariance
        down = sp.SparseConv3d(cin, cout, 3, stride=2, padding=1, bias=False)
        down.weight.data = (torch.randn(down.weight.shape, generator=g) / (27 * cin) ** 0.5 * 3).half().float()
        down = down.to(dev).eval()
        ds = down(_tensor(idx, feats, shape, True))
        du = down(_tensor(idx[perm].contiguous(), feats[perm].contiguous(), shape, False))
        oshape = [(s + 2 - 3) // 2 + 1 for s in shape]
        assert list(ds.spatial_shape) == oshape
        keys = []
        zyx = idx[:, 1:].long()
        for kz in range(3):
            for ky in range(3):
                for kx in range(3):
                    num = zyx + 1 - torch.tensor([kz, ky, kx], device=zyx.device)
                    o = torch.div(num, 2, rounding_mode='floor')
                    ok = ((num >= 0) & (num % 2 == 0) & (o < torch.tensor(oshape, device=zyx.device))).all(1)
                    o = o[ok]
                    keys.append((o[:, 0] * oshape[1] + o[:, 1]) * oshape[2] + o[:, 2])
        want = torch.unique(torch.cat(keys))
        n_out = ds.features.shape[0]
        got = (ds.indices[:n_out, 1].long() * oshape[1] + ds.indices[:n_out, 2].long()) * oshape[2] + ds.indices[:n_out, 3].long()
        assert torch.equal(got, want), 'strided conv: output sites differ from the sorted unique candidate set'
        assert torch.equal(du.indices[:n_out], ds.indices[:n_out]) and torch.equal(du.features, ds.features)
    return v, n_out


def prop_dynamic_scatter(dev, n, c, seed=0):
    from isfusion_b200.voxel_layer import dynamic_point_to_voxel_forward
    g = torch.Generator().manual_seed(seed)
    feats = (torch.rand(n, c, generator=g) * 100 - 50).to(dev)
    coors = torch.randint(-1, 40, (n, 3), generator=g, dtype=torch.int32).to(dev)
    fsum, cs, m, cnt = dynamic_point_to_voxel_forward(feats, coors, 'sum')
    fmean, cm, m2, cnt2 = dynamic_point_to_voxel_forward(feats, coors, 'mean')
    fmax, cx, m3, cnt3 = dynamic_point_to_voxel_forward(feats, coors, 'max')
    valid = (coors >= 0).all(1)
    assert int(cnt.sum()) == int(valid.sum()) and torch.equal(m, m2) and torch.equal(m, m3) and torch.equal(cnt, cnt2)
    assert torch.equal(cs, cm) and torch.equal(cs, cx)
    key = (cs[:, 0].long() * 64 + cs[:, 1].long()) * 64 + cs[:, 2].long()
    assert (key[1:] > key[:-1]).all(), 'out_coors are not strictly ascending (sorted unique)'
    assert torch.equal((m >= 0), valid) and int(m.max()) == cs.shape[0] - 1
    assert torch.allclose(fmean * cnt.unsqueeze(1), fsum, rtol=1e-4, atol=1e-2)
    assert (fmax >= fmean - 1e-3).all()
    assert torch.equal(cs[m[valid].long()], coors[valid])          # every valid point maps to ITS voxel


def prop_heatmap_topk(dev, batch, classes, grid, k, seed=0):
    from isfusion_b200 import _lib as L
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(batch, classes, grid, grid, generator=g).to(dev)
    wsb = ctypes.c_size_t(0)
    L.check(L.lib().isf_heatmap_topk_workspace_size(batch, classes, grid, grid, ctypes.byref(wsb)))
    ws = L.workspace(wsb.value, logits.device)
    top = torch.empty((batch, k), dtype=torch.int32, device=dev)
    score = torch.empty((batch, k), dtype=torch.float32, device=dev)
    L.check(L.lib().isf_heatmap_topk(L.ptr(logits), batch, classes, grid, grid, classes - 2, k, L.ptr(top), L.ptr(score), L.ptr(ws),
                                     ctypes.c_size_t(ws.numel()), L.stream()))
    assert (score[:, 1:] <= score[:, :-1]).all() and int(top.min()) >= 0 and int(top.max()) < grid * grid
    hm = logits.sigmoid().view(batch, classes, -1)
    best = hm.gather(2, top.long().unsqueeze(1).expand(-1, classes, -1)).max(1).values      # the cell's best class score
    assert (score <= best + 1e-6).all() and torch.allclose(score[:, 0], hm.view(batch, -1).max(1).values, atol=1e-6)


def prop_msda_constant_map(dev, seed=0):
    import isfusion_b200.fusion_encoder as FE
    g = torch.Generator().manual_seed(seed)
    B, H, D, Q, P = 2, 8, 16, 200, 16
    shapes = torch.tensor([[40, 44], [20, 22]])
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    const = torch.randn(1, 1, H, D, generator=g)
    value = const.expand(B, S, H, D).contiguous().to(dev)
    lo = 0.5 / 20 + 1e-3                                   # strictly inside the coarsest map: all four corners exist
    loc = (torch.rand(B, Q, H, 2, P, 2, generator=g) * (1 - 2 * lo) + lo).to(dev)
    aw = torch.softmax(torch.randn(B, Q, H, 2 * P, generator=g), -1).view(B, Q, H, 2, P).to(dev)
    lvl = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
    out = FE.ms_deform_attn(value, shapes.to(dev), lvl.to(dev), loc, aw)
    assert torch.allclose(out.view(B, Q, H, D).cpu(), const.expand(B, Q, H, D), atol=1e-5)


# ------------------------------------------------------------------------------------------------ CPU: emulator, small sizes
def test_properties_on_the_emulator(monkeypatch):
    import abi_emulator
    abi_emulator.install(monkeypatch)
    shape = [11, 80, 80]
    idx = _voxels(6000, shape, [1.35, 1.35, 0.8], [-54.0, -54.0, -5.0], 0, 'cpu')
    assert idx.shape[0] > 500
    v, n_out = prop_sparse_conv('cpu', idx, shape, 16, 16)
    assert n_out > 0 and v > 0
    prop_sparse_conv('cpu', idx, shape, 16, 32, seed=1)
    prop_dynamic_scatter('cpu', 5000, 4)
    prop_heatmap_topk('cpu', 2, 10, 24, 50)
    prop_msda_constant_map('cpu')


# ------------------------------------------------------------------------------------------------ GPU: full-frame sizes
@pytest.mark.gpu
@pytest.mark.parametrize('cin,cout', [(32, 32), (64, 64), (128, 128), (32, 64)])
def test_sparse_conv_properties_on_the_full_frame(cin, cout):
    shape = [41, 1440, 1440]
    idx = _voxels(300000, shape, [0.075, 0.075, 0.2], [-54.0, -54.0, -5.0], 0, 'cuda')
    v, n_out = prop_sparse_conv('cuda', idx, shape, cin, cout)
    assert v > 100000 and n_out > 50000


@pytest.mark.gpu
def test_scatter_topk_msda_properties_at_full_size():
    prop_dynamic_scatter('cuda', 1000000, 5)
    prop_dynamic_scatter('cuda', 300000, 64, seed=1)
    prop_heatmap_topk('cuda', 4, 10, 180, 200)
    prop_heatmap_topk('cuda', 1, 10, 256, 200, seed=1)
    prop_msda_constant_map('cuda')
