"""GPU parity: sparse conv rulebooks (bit-exact indices) and implicit GEMM vs the oracle."""
import pytest
import torch

from oracle import spconv_ops as so

pytestmark = pytest.mark.gpu


def _rand_tensor(shape, nvox, c, batch, seed, clustered=False):
    g = torch.Generator().manual_seed(seed)
    cells = batch * shape[0] * shape[1] * shape[2]
    if clustered:   # neighbours exist: random walk blobs
        base = torch.randint(0, cells, (nvox // 8,), generator=g)
        lin = (base[:, None] + torch.randint(-2, 3, (nvox // 8, 8), generator=g) * torch.tensor([1, shape[2], shape[1] * shape[2], 1, shape[2], 1, 1, shape[2]])).flatten() % cells
        lin = torch.unique(lin)
        lin = lin[torch.randperm(lin.numel(), generator=g)]
    else:
        lin = torch.randperm(cells, generator=g)[:nvox]
    nvox = lin.numel()
    idx = torch.stack([lin // (shape[0] * shape[1] * shape[2]), (lin // (shape[1] * shape[2])) % shape[0],
                       (lin // shape[2]) % shape[1], lin % shape[2]], 1).int()
    return so.SparseConvTensor(torch.randn(nvox, c, generator=g).half().float(), idx, shape, batch)


def _to_gpu(x):
    from isfusion_b200.spconv import SparseConvTensor
    return SparseConvTensor(x.features.half().cuda(), x.indices.cuda(), x.spatial_shape, x.batch_size)


CASES = [  # (subm, cin, cout, ksize, stride, padding)
    (True, 64, 32, 3, 1, 1), (True, 32, 32, 3, 1, 1), (False, 32, 64, 3, 2, 1), (True, 128, 128, 3, 1, 1),
    (False, 128, 256, 3, 2, (0, 1, 1)), (True, 256, 256, 3, 1, 1), (False, 256, 256, (3, 1, 1), (2, 1, 1), 0),
]


@pytest.mark.parametrize('subm,cin,cout,ksize,stride,padding', CASES)
def test_conv_parity(subm, cin, cout, ksize, stride, padding):
    import isfusion_b200.spconv as sp
    x = _rand_tensor([11, 40, 36], 6000, cin, 2, cin + cout, clustered=True)
    cls_o = so.SubMConv3d if subm else so.SparseConv3d
    cls_g = sp.SubMConv3d if subm else sp.SparseConv3d
    ref = cls_o(cin, cout, ksize, stride=stride, padding=padding, bias=False)
    ref.weight.data = (torch.randn_like(ref.weight) / (27 * cin) ** 0.5 * 3).half().float()
    mod = cls_g(cin, cout, ksize, stride=stride, padding=padding, bias=False)
    mod.load_state_dict(ref.state_dict())
    mod = mod.cuda().eval()
    ro = ref(x)
    go = mod(_to_gpu(x))
    assert go.spatial_shape == ro.spatial_shape
    assert torch.equal(go.indices.cpu(), ro.indices)                      # output sites + row order bit-exact
    rb = list(go.indice_dict.values())[-1]
    nbr_ref = so.get_indice_pairs(x.indices, x.batch_size, x.spatial_shape, ref.kernel_size, ref.stride, ref.padding,
                                  ref.dilation, subm)[2]
    assert torch.equal(rb.nbr[:, :rb.num_out].cpu().long(), nbr_ref)      # rulebook bit-exact
    err = (go.features.float().cpu() - ro.features).abs().max().item()
    assert err <= 1e-3 * max(1.0, ro.features.abs().max().item()), err    # 1e-3 rel (fp16 storage, fp32 accumulate)
    # spconv2 weight layout is accepted on load
    sd = {'weight': ref.weight.data.permute(4, 0, 1, 2, 3).contiguous()}
    mod2 = cls_g(cin, cout, ksize, stride=stride, padding=padding, bias=False)
    mod2.load_state_dict(sd)
    assert torch.equal(mod2.weight.data, ref.weight.data)


def test_conv_fused_epilogue_and_dense():
    import isfusion_b200.spconv as sp
    x = _rand_tensor([9, 30, 30], 3000, 64, 2, 5, clustered=True)
    ref = so.SubMConv3d(64, 64, 3, padding=1, bias=False)
    mod = sp.SubMConv3d(64, 64, 3, padding=1, bias=False)
    mod.load_state_dict(ref.state_dict())
    mod = mod.cuda().eval()
    scale, shift = torch.rand(64) + 0.5, torch.randn(64) * 0.1
    res = torch.randn(x.features.shape[0], 64).half()
    gx = _to_gpu(x)
    go = mod(gx, scale=scale.cuda(), shift=shift.cuda(), residual=res.cuda(), relu=True)
    want = torch.relu(ref(x).features * scale + shift + res.float())
    assert (go.features.float().cpu() - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
    dense = go.dense().float().cpu()
    want_d = so.SparseConvTensor(go.features.float().cpu(), x.indices, x.spatial_shape, x.batch_size).dense()
    assert torch.equal(dense, want_d)


def _randomize_bn(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.8, 1.2)
            mod.bias.data.normal_(0, 0.1)


def test_sparse_encoder_parity():
    """whole SparseEncoder (isfusion_0075voxel.py:72-82 channel plan) on a reduced grid, LiDAR-like clustered input."""
    from isfusion_b200.sparse_encoder import SparseEncoder
    torch.manual_seed(0)
    cfg = dict(in_channels=64, sparse_shape=[41, 128, 128], base_channels=32, output_channels=256,
               encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')
    ref = so.SparseEncoder(**cfg).eval()
    _randomize_bn(ref)
    for p in ref.parameters():
        if p.dim() == 5:
            p.data = p.data.half().float()
    x = _rand_tensor([41, 128, 128], 20000, 64, 2, 11, clustered=True)
    x.features = torch.relu(x.features)
    rs, rfeats, _ = ref(x.features, x.indices, 2)
    mod = SparseEncoder(**cfg)
    missing = mod.load_state_dict(ref.state_dict(), strict=True)
    mod = mod.cuda().eval()
    gs, gfeats, _ = mod(x.features.half().cuda(), x.indices.cuda(), 2)
    for a, b in zip(gfeats, rfeats):
        assert a.spatial_shape == b.spatial_shape and torch.equal(a.indices.cpu(), b.indices)
    assert gs.shape == rs.shape == (2, 512, 16, 16)
    num = (gs.float().cpu() - rs).norm().item()
    den = rs.norm().item()
    mx = (gs.float().cpu() - rs).abs().max().item()
    print('sparse encoder rel l2 err %.3e  max abs %.3e (max ref %.3f)' % (num / den, mx, rs.abs().max().item()))
    assert num / den <= 1e-3, (num / den)       # measured 2.1e-4


def test_sparse_encoder_parity_on_the_config_grid():
    """the whole SparseEncoder at the config's own geometry (isfusion_0075voxel.py: 41 x 1440 x 1440 grid, 0.075 m voxels) on the
    300 k-point LiDAR-shaped cloud of the bench workload: voxel coordinates from the oracle's dynamic voxelisation, random fp16
    voxel features; output sites of every stage bit-exact against the oracle, BEV features rel-L2 <= 1e-3 (measured 2.0e-4 over
    the 21 chained fp16 layers)."""
    import synth
    from oracle import voxel_ops as vo
    from isfusion_b200.sparse_encoder import SparseEncoder
    torch.manual_seed(1)
    shape = [41, 1440, 1440]
    cfg = dict(in_channels=64, sparse_shape=shape, base_channels=32, output_channels=256,
               encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')
    pts = synth.make_points(300000, 'S', 8)
    coors = vo.dynamic_voxelize(pts, [0.075, 0.075, 0.2], synth.PC_RANGE)              # [N, 3] (z, y, x), -1 = out of range
    coors = coors[(coors >= 0).all(1)]
    lin = (coors[:, 0].long() * shape[1] + coors[:, 1].long()) * shape[2] + coors[:, 2].long()
    lin = torch.unique(lin)                                                            # sorted: the order DynamicVFE emits
    idx = torch.stack([torch.zeros_like(lin), lin // (shape[1] * shape[2]), (lin // shape[2]) % shape[1], lin % shape[2]], 1).int()
    assert idx.shape[0] > 100000
    feats = torch.relu(torch.randn(idx.shape[0], 64)).half().float()
    ref = so.SparseEncoder(**cfg).eval()
    _randomize_bn(ref)
    for p in ref.parameters():
        if p.dim() == 5:
            p.data = p.data.half().float()
    rs, rfeats, _ = ref(feats, idx, 1)
    mod = SparseEncoder(**cfg)
    mod.load_state_dict(ref.state_dict(), strict=True)
    mod = mod.cuda().eval()
    gs, gfeats, _ = mod(feats.half().cuda(), idx.cuda(), 1)
    for a, b in zip(gfeats, rfeats):
        assert a.spatial_shape == b.spatial_shape and torch.equal(a.indices.cpu(), b.indices)
    assert gs.shape == rs.shape == (1, 512, 180, 180)
    num, den = (gs.float().cpu() - rs).norm().item(), rs.norm().item()
    mx = (gs.float().cpu() - rs).abs().max().item()
    print('sparse encoder on 41x1440x1440 (%d voxels): rel l2 err %.3e  max abs %.3e (max ref %.3f)' % (idx.shape[0], num / den, mx, rs.abs().max().item()))
    assert num / den <= 1e-3, (num / den)


@pytest.mark.parametrize('cs', [2])
def test_sparse_encoder_cluster_variant_bit_identical(cs):
    from isfusion_b200 import _lib as L
    from isfusion_b200.sparse_encoder import SparseEncoder
    torch.manual_seed(0)
    cfg = dict(in_channels=64, sparse_shape=[41, 128, 128], base_channels=32, output_channels=256,
               encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')
    mod = SparseEncoder(**cfg).cuda().eval()
    x = _rand_tensor([41, 128, 128], 20000, 64, 2, 11, clustered=True)
    f, c = torch.relu(x.features).half().cuda(), x.indices.cuda()
    try:
        L.set_option('gemm_cluster', 1)
        ref, _, _ = mod(f, c, 2)
        L.set_option('gemm_cluster', cs)
        out, _, _ = mod(f, c, 2)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
    finally:
        L.set_option('gemm_cluster', 1)


def test_bitmap_rulebook_matches_hash_rulebook():
    """sorted-row tensors take the bitmap-rank rulebook; it must give the hash rulebook's bits (indices, nbr, features)."""
    from isfusion_b200.sparse_encoder import SparseEncoder
    import isfusion_b200.spconv as sp
    torch.manual_seed(0)
    cfg = dict(in_channels=64, sparse_shape=[41, 128, 128], base_channels=32, output_channels=256,
               encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')
    mod = SparseEncoder(**cfg).cuda().eval()
    x = _rand_tensor([41, 128, 128], 20000, 64, 2, 3, clustered=True)
    key = so._lin(x.indices, x.spatial_shape)
    order = torch.argsort(key)
    f, c = torch.relu(x.features[order]).half().cuda(), x.indices[order].contiguous().cuda()
    ref, rfe, _ = mod(f, c, 2, sorted_coors=False)
    out, ofe, _ = mod(f, c, 2, sorted_coors=True)
    torch.cuda.synchronize()
    for a, b in zip(ofe, rfe):
        assert torch.equal(a.indices, b.indices)
    assert torch.equal(out, ref)
    # oracle cross-check of one strided rulebook built by the bitmap path
    xs = sp.SparseConvTensor(f, c, [41, 128, 128], 2)
    xs.sorted = True
    old = sp.TILE_ORDER
    try:
        sp.TILE_ORDER = None        # the plain table (the fused ordered builder never materialises it)
        rb = sp.build_rulebook(xs, [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False)
    finally:
        sp.TILE_ORDER = old
    oi, osz, nbr = so.get_indice_pairs(c.cpu(), 2, [41, 128, 128], 3, 2, 1, 1, False)
    assert torch.equal(rb.out_indices.cpu(), oi) and torch.equal(rb.nbr[:, :rb.num_out].cpu().long(), nbr)


def test_capacity_overflow_drops_sites_from_the_coordinate_map_too():
    """capacity (sync-free) mode with an output buffer SMALLER than the true site count: the surplus sites are dropped from
    the index list AND from the bitmap the next layer uses as its coordinate map, so no neighbour entry of the next layer
    points past the [cap, C] feature buffer (it used to: rank = prefix + popc >= cap -> out-of-bounds gather), the kept
    rows equal the first `cap` rows of the exact result, and check_overflow() reports the layer."""
    import isfusion_b200.spconv as sp
    x = _rand_tensor([41, 96, 96], 30000, 16, 2, 3, clustered=True)
    order = torch.argsort(so._lin(x.indices, x.spatial_shape))
    idx = x.indices[order].contiguous().cuda()
    n = idx.shape[0]
    exact = sp.SparseConvTensor(None, idx, x.spatial_shape, 2)
    exact.sorted = True
    rb_exact = sp.build_rulebook(exact, [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False)
    true_out = rb_exact.num_out
    cap = true_out // 2
    sp.check_overflow()
    num_dev = torch.tensor([n], dtype=torch.int32, device='cuda')
    xt = sp.SparseConvTensor(None, idx, x.spatial_shape, 2, num_dev=num_dev)
    xt.sorted = True
    rb = sp.build_rulebook(xt, [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False, out_capacity=cap)
    torch.cuda.synchronize()
    assert int(rb.num_dev.item()) == cap
    assert torch.equal(rb.out_indices[:cap], rb_exact.out_indices[:cap])
    # the next layer (a SubM conv over the strided conv's output) looks its neighbours up in the output bitmap
    yt = sp.SparseConvTensor(None, rb.out_indices, rb.out_shape, 2, num_dev=rb.num_dev)
    yt.sorted = True
    yt.indice_dict = xt.indice_dict
    rb2 = sp.build_rulebook(yt, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], True)
    torch.cuda.synchronize()
    tab = rb2.nbr_ordered if rb2.nbr_ordered is not None else rb2.nbr
    assert int(tab.max().item()) < cap, 'a neighbour entry points past the capacity-sized feature buffer'
    bad = sp.check_overflow()
    assert bad and bad[0][1] == cap


@pytest.mark.parametrize('order', [None, 31])
def test_sparse_encoder_v4_schedule_bit_identical(order):
    """isf_set_option("gemm_v4", 1): dedicated epilogue warps + double-buffered TMEM accumulator (csrc/tc_gemm4.cuh): every
    output row is accumulated over the same chunks in the same order and goes through the same epilogue arithmetic -> same bits."""
    from isfusion_b200 import _lib as L
    from isfusion_b200.sparse_encoder import SparseEncoder
    import isfusion_b200.spconv as sp
    torch.manual_seed(0)
    cfg = dict(in_channels=64, sparse_shape=[41, 128, 128], base_channels=32, output_channels=256,
               encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')
    mod = SparseEncoder(**cfg).cuda().eval()
    _randomize_bn(mod)
    x = _rand_tensor([41, 128, 128], 40000, 64, 2, 11, clustered=True)
    f, c = torch.relu(x.features).half().cuda(), x.indices.cuda()
    old, old_dyn = sp.TILE_ORDER, sp.DYNAMIC_TILES
    try:
        sp.TILE_ORDER = order
        sp.DYNAMIC_TILES = False
        L.set_option('gemm_pair', 0)
        L.set_option('gemm_v4', 0)
        ref, _, _ = mod(f, c, 2)
        L.set_option('gemm_v4', 2)
        out, _, _ = mod(f, c, 2)
        L.set_option('gemm_mt', 2)               # 256-row CTA tiles of the same schedule
        out2, _, _ = mod(f, c, 2)
        L.set_option('gemm_zskip', 1)            # + zero-skip of all-missing 4-row copies
        out3, _, _ = mod(f, c, 2)
        L.set_option('gemm_mt', 3)
        out4, _, _ = mod(f, c, 2)
        torch.cuda.synchronize()
        assert torch.equal(out, ref) and torch.equal(out2, ref) and torch.equal(out3, ref) and torch.equal(out4, ref)
        L.set_option('gemm_zskip', 0)
        # producer mapping (all eight warps per chunk / one warp per chunk), tile walk direction, 32-column layers on v4
        for wpc, rev, v32, mt in ((0, 0, 0, 3), (1, 0, 0, 3), (0, 1, 1, 3), (1, 1, 1, 3), (2, 1, 2, 3), (1, 1, 1, 2), (2, 0, 2, 1), (2, 1, 1, 2)):
            L.set_option('gemm_wpc', wpc)
            L.set_option('gemm_rev', rev)
            L.set_option('gemm_v4_32', v32)
            L.set_option('gemm_mt', mt)
            o, _, _ = mod(f, c, 2)
            torch.cuda.synchronize()
            assert torch.equal(o, ref), (wpc, rev, v32, mt)
        # dynamic tile scheduling (isf_spconv_tile_counter): twice per setting, the second run finds the counters re-armed
        sp.DYNAMIC_TILES = True
        for pair, v32 in ((0, 1), (0, 1), (128, 1), (128, 1), (1, 2), (1, 2)):
            L.set_option('gemm_pair', pair)
            L.set_option('gemm_v4_32', v32)
            L.set_option('gemm_mt', 3)
            o, _, _ = mod(f, c, 2)
            torch.cuda.synchronize()
            assert torch.equal(o, ref), ('dynamic tiles', pair, v32)
        sp.DYNAMIC_TILES = False
        # CTA pairs (tcgen05.mma.cta_group::2, csrc/tc_gemm5.cuh): every column tile >= 64 / only 256 / 128-and-up, both walks
        for pair, rev, mt, mma2 in ((1, 0, 3, 1), (1, 1, 3, 0), (256, 0, 3, 1), (128, 0, 3, 1), (128, 0, 3, 0), (128, 0, 1, 1), (1, 0, 1, 0)):
            L.set_option('gemm_pair', pair)
            L.set_option('gemm_rev', rev)
            L.set_option('gemm_mt', mt)
            L.set_option('gemm_mma2', mma2)
            o, _, _ = mod(f, c, 2)
            torch.cuda.synchronize()
            assert torch.equal(o, ref), ('pair', pair, rev, mt)
    finally:
        L.set_option('gemm_mma2', 1)
        L.set_option('gemm_pair', 128)
        L.set_option('gemm_v4', 2)
        L.set_option('gemm_mt', 3)
        L.set_option('gemm_zskip', 0)
        L.set_option('gemm_wpc', 0)
        L.set_option('gemm_rev', 0)
        L.set_option('gemm_v4_32', 1)
        sp.TILE_ORDER, sp.DYNAMIC_TILES = old, old_dyn
