"""Tile order of the sparse-conv implicit GEMM (csrc/spconv.cu, isf_spconv_tile_order; the default since round 2,
isfusion_b200.spconv.TILE_ORDER / ISF_TILE_ORDER).

CPU part: the host-side key layout and a numpy model of the order (tests/analysis/tile_occupancy.py) on an oracle rulebook.
GPU part: the device
order equals the model and every convolution over the regrouped table is BIT-IDENTICAL to the default launch."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_ORDER = 31        # isfusion_b200.spconv.TILE_ORDER as shipped (restored after every test that switches it)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'analysis'))


def test_key_bits_rank_z_moving_and_far_offsets_first():
    import tile_occupancy as to
    pos = to.library_tap_bits((3, 3, 3))
    assert sorted(pos) == list(range(27))
    offs = [(kz - 1, ky - 1, kx - 1) for kz in range(3) for ky in range(3) for kx in range(3)]
    assert pos[13] == 0                                              # the centre tap (always present) is the LSB
    z_moving = [pos[k] for k, o in enumerate(offs) if o[0] != 0]
    in_plane = [pos[k] for k, o in enumerate(offs) if o[0] == 0]
    assert min(z_moving) > max(in_plane)                             # 18 z-moving taps own the 18 top bits
    corner = [pos[k] for k, o in enumerate(offs) if all(o)]
    assert min(corner) >= 27 - 8                                     # the 8 corners sort first of all
    assert to.library_tap_bits((3, 1, 1)) == [2, 0, 1] or sorted(to.library_tap_bits((3, 1, 1))) == [0, 1, 2]
    from isfusion_b200 import _lib as L
    bad = (ctypes.c_ubyte * 32)()
    assert L.lib().isf_spconv_tile_order_bits(L.host_i32((3, 3, 3)), 9, bad) == 1      # taps != kD*kH*kW


def test_model_order_is_a_stable_permutation_and_cuts_zero_slots():
    import synth
    import tile_occupancy as to
    from oracle import spconv_ops as so, voxel_ops
    pts = synth.make_points(30000, 'S', 4)
    c = voxel_ops.dynamic_voxelize(pts, [0.075, 0.075, 0.2], synth.PC_RANGE)
    c = c[(c >= 0).all(1)]
    idx = torch.unique(torch.cat([torch.zeros(len(c), 1, dtype=c.dtype), c], 1), dim=0).int()
    idx, shape, _ = so.get_indice_pairs(idx, 1, [41, 1440, 1440], 3, 2, [1, 1, 1], 1, False)
    _, _, nbr = so.get_indice_pairs(idx, 1, shape, 3, 1, 1, 1, True)
    valid = (nbr >= 0).numpy()
    v = valid.shape[1]
    base, _ = to.issued_slots(valid, np.arange(v))
    for chunk in (12, 31):
        perm = to.tile_order_model(valid, chunk)
        assert np.array_equal(np.sort(perm), np.arange(v))
        if chunk < 31:                                               # rows never leave their chunk
            assert np.array_equal(perm >> chunk, np.arange(v) >> chunk)
        got, _ = to.issued_slots(valid, perm)
        assert got < 0.85 * base, (chunk, got, base)
    # stable: rows with identical masks keep their storage order
    perm = to.tile_order_model(valid, 31)
    key = to.mask_rank(to.occupancy_masks(valid))
    assert (np.diff(key[perm].astype(np.int64)) >= 0).all()
    same = key[perm][1:] == key[perm][:-1]
    assert (perm[1:][same] > perm[:-1][same]).all()


# ------------------------------------------------------------------------------------------------ GPU bring-up
def _device_order(nbr, v, chunk, num_dev=None, ksize=(3, 3, 3)):
    from isfusion_b200 import _lib as L
    lib = L.lib()
    taps, stride = nbr.shape
    wsb = ctypes.c_size_t(0)
    L.check(lib.isf_spconv_tile_order_workspace_size(v, ctypes.byref(wsb)))
    ws = torch.empty(wsb.value, dtype=torch.uint8, device='cuda')
    out = torch.full_like(nbr, -7)
    perm = torch.full((stride,), -7, dtype=torch.int32, device='cuda')
    L.check(lib.isf_spconv_tile_order(L.ptr(nbr), ctypes.c_longlong(stride), taps, L.host_i32(ksize), v, L.ptr(num_dev), chunk,
                                      L.ptr(out), L.ptr(perm), L.ptr(ws), ctypes.c_size_t(ws.numel()), L.stream()))
    torch.cuda.synchronize()
    return out, perm


@pytest.mark.gpu
@pytest.mark.parametrize('chunk', [7, 12, 31])
def test_device_order_equals_model(chunk):
    import tile_occupancy as to
    from isfusion_b200.spconv import SparseConvTensor, build_rulebook
    from test_spconv_gpu import _rand_tensor
    x = _rand_tensor([41, 96, 96], 30000, 16, 2, 3, clustered=True)
    xt = SparseConvTensor(x.features.half().cuda(), x.indices.cuda(), x.spatial_shape, x.batch_size)
    rb = build_rulebook(xt, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], True)
    v = rb.num_out
    out, perm = _device_order(rb.nbr, v, chunk)
    valid = (rb.nbr[:, :v] >= 0).cpu().numpy()
    model = to.tile_order_model(valid, chunk)
    assert np.array_equal(perm[:v].cpu().numpy(), model)
    assert (perm[v:] == -1).all() and (out[:, v:] == -1).all()
    from isfusion_b200 import _lib as L
    try:        # fewer sorted bits: same model with the low key bits dropped
        L.set_option('tile_order_drop_bits', 11)
        _, perm11 = _device_order(rb.nbr, v, chunk)
        assert np.array_equal(perm11[:v].cpu().numpy(), to.tile_order_model(valid, chunk, drop_bits=11))
    finally:
        L.set_option('tile_order_drop_bits', 0)
    assert torch.equal(out[:, :v], rb.nbr[:, perm[:v].long()])
    # capacity mode: only the first n rows are valid, the tail of the capacity sorts last and is dropped
    n = v - 1000
    num_dev = torch.tensor([n], dtype=torch.int32, device='cuda')
    out2, perm2 = _device_order(rb.nbr, v, chunk, num_dev)
    assert np.array_equal(perm2[:n].cpu().numpy(), to.tile_order_model(valid[:, :n], chunk))
    assert (perm2[n:] == -1).all() and (out2[:, n:] == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize('chunk', [12, 31])
def test_fused_ordered_rulebook_equals_two_call_path(chunk):
    """isf_spconv_neighbors_bitmap_ordered == isf_spconv_neighbors_bitmap + isf_spconv_tile_order, every geometry."""
    import isfusion_b200.spconv as sp
    from test_spconv_gpu import _rand_tensor
    x = _rand_tensor([41, 96, 96], 30000, 16, 2, 3, clustered=True)
    order = torch.argsort(((x.indices[:, 0].long() * 41 + x.indices[:, 1]) * 96 + x.indices[:, 2]) * 96 + x.indices[:, 3])
    idx = x.indices[order].contiguous().cuda()
    geos = [([3, 3, 3], [1, 1, 1], [1, 1, 1], True), ([3, 3, 3], [2, 2, 2], [1, 1, 1], False),
            ([3, 3, 3], [2, 2, 2], [0, 1, 1], False), ([3, 1, 1], [2, 1, 1], [0, 0, 0], False)]
    try:
        sp.TILE_ORDER = chunk
        for ks, st, pd, subm in geos:
            res = []
            for fused in (False, True):
                sp.TILE_ORDER_FUSED = fused
                xt = sp.SparseConvTensor(None, idx, x.spatial_shape, x.batch_size)
                xt.sorted = True
                rb = sp.build_rulebook(xt, ks, st, pd, [1, 1, 1], subm)
                torch.cuda.synchronize()
                res.append((rb.nbr_ordered.clone(), rb.row_perm.clone()))
            assert torch.equal(res[0][1], res[1][1]), (ks, st, pd, subm)
            assert torch.equal(res[0][0], res[1][0]), (ks, st, pd, subm)
    finally:
        sp.TILE_ORDER, sp.TILE_ORDER_FUSED = DEFAULT_ORDER, True


@pytest.mark.gpu
@pytest.mark.parametrize('chunk', [12, 31])
def test_sparse_encoder_bit_identical_with_tile_order(chunk):
    import isfusion_b200.spconv as sp
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
        sp.TILE_ORDER = None
        ref, _, _ = mod(f, c, 2)
        sp.TILE_ORDER = chunk
        out, _, _ = mod(f, c, 2)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
    finally:
        sp.TILE_ORDER = DEFAULT_ORDER


@pytest.mark.gpu
def test_graphed_neck_bit_identical_with_tile_order():
    """capacity mode + CUDA graph: device-side row counts, rulebooks (and their orders) built on the side stream."""
    import isfusion_b200.spconv as sp
    import synth
    from __graft_entry__ import make_pair
    from isfusion_b200.detector import GraphedNeck
    _, net = make_pair(0)
    net = net.cuda().eval()
    pts = [synth.make_points(20000, 'S', 5).cuda()]
    l2i, aug, laug = synth.make_calib(1)
    feats = [t.cuda() for t in synth.make_img_feats(1)]
    metas = [dict(input_shape=(256, 704))]
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    outs = {}
    try:
        for mode in (None, 31):
            sp.TILE_ORDER = mode
            caps = GraphedNeck.calibrate(net, [pts], feats, metas, **kw)
            g = GraphedNeck(net, 1, pts[0].shape[0] + 500, calib=kw, capacities=caps)
            g.load(pts, feats[1])
            g.capture()
            o, hm = g(pts, feats[1])
            torch.cuda.synchronize()
            assert not g.check()
            outs[mode] = [t.clone() for t in o] + [hm.clone()]
        for a, b in zip(outs[None], outs[31]):
            assert torch.equal(a, b)
    finally:
        sp.TILE_ORDER = DEFAULT_ORDER


def test_issued_flops_counts_exactly_the_chunks_the_kernel_walks():
    """bench.py's `achieved_issued` denominator: 64-wide K chunks skipped per 128-row tile when none of their taps has a
    neighbour there (CPU tensors; the helper is pure index arithmetic)."""
    from isfusion_b200.spconv import Rulebook, issued_flops
    taps, n = 27, 300                                   # 3 tiles (128, 128, 44 rows)
    nbr = torch.full((taps, 384), -1, dtype=torch.int32)
    nbr[13, :n] = torch.arange(n, dtype=torch.int32)    # centre tap everywhere
    nbr[0, 5] = 7                                       # tap 0 only in tile 0
    nbr[26, 290] = 1                                    # tap 26 only in tile 2
    rb = Rulebook(None, None, nbr, 384, taps, n)
    # cin = 64: one chunk per tap -> tiles have 2, 1, 2 active chunks
    assert issued_flops(rb, 64, 32) == 2.0 * 5 * 128 * 64 * 32
    # cin = 128: two chunks per tap
    assert issued_flops(rb, 128, 128) == 2.0 * 10 * 128 * 64 * 128
    # cin = 32: chunk q covers taps 2q, 2q+1 -> taps 0 | 12,13 | 26 sit in chunks 0, 6, 13
    assert issued_flops(rb, 32, 64) == 2.0 * 5 * 128 * 64 * 64
    # a tile without any neighbour still costs one chunk (the accumulator must be defined)
    nbr2 = torch.full((taps, 256), -1, dtype=torch.int32)
    nbr2[13, :128] = 0
    assert issued_flops(Rulebook(None, None, nbr2, 256, taps, 256), 64, 32) == 2.0 * 2 * 128 * 64 * 32


@pytest.mark.gpu
def test_row_per_thread_neighbour_kernel_bit_identical():
    """isf_set_option("nbr_rows", 1): same neighbour tables as the tap-per-thread kernel for every conv geometry of the
    encoder, exact and capacity mode."""
    from isfusion_b200 import _lib as L
    from isfusion_b200.spconv import SparseConvTensor, build_rulebook
    from test_spconv_gpu import _rand_tensor
    x = _rand_tensor([41, 96, 96], 30000, 16, 2, 3, clustered=True)
    order = torch.argsort(((x.indices[:, 0].long() * 41 + x.indices[:, 1]) * 96 + x.indices[:, 2]) * 96 + x.indices[:, 3])
    idx = x.indices[order].contiguous().cuda()
    geos = [([3, 3, 3], [1, 1, 1], [1, 1, 1], True), ([3, 3, 3], [2, 2, 2], [1, 1, 1], False),
            ([3, 3, 3], [2, 2, 2], [0, 1, 1], False), ([3, 1, 1], [2, 1, 1], [0, 0, 0], False)]
    try:
        for ks, st, pd, subm in geos:
            tabs = []
            for opt in (0, 1):
                L.set_option('nbr_rows', opt)
                xt = SparseConvTensor(None, idx, x.spatial_shape, x.batch_size)
                xt.sorted = True                                   # bitmap-rank rulebook
                rb = build_rulebook(xt, ks, st, pd, [1, 1, 1], subm)
                torch.cuda.synchronize()
                tabs.append((rb.nbr[:, :rb.num_out].clone(), rb.out_indices.clone()))
            assert torch.equal(tabs[0][1], tabs[1][1])
            assert torch.equal(tabs[0][0], tabs[1][0]), (ks, st, pd, subm)
    finally:
        L.set_option('nbr_rows', 0)


def _host_row_walk(in_idx, out_idx, in_shape, batch, ks, st, pd, subm):
    """run the row-per-thread kernels' walk (csrc/spconv.cu nbr_row_walk, compiled for the host) on numpy arrays."""
    from isfusion_b200 import _lib as L
    lib = L.lib()
    cells = batch * in_shape[0] * in_shape[1] * in_shape[2]
    words = (cells + 31) // 32
    key = ((in_idx[:, 0].astype(np.int64) * in_shape[0] + in_idx[:, 1]) * in_shape[1] + in_idx[:, 2]) * in_shape[2] + in_idx[:, 3]
    assert (np.diff(key) > 0).all(), 'bitmap-rank rulebooks need rows in ascending key order'
    bits = np.zeros(words, np.uint32)
    np.bitwise_or.at(bits, key >> 5, (np.uint32(1) << (key & 31).astype(np.uint32)))
    pop = np.array([bin(int(w)).count('1') for w in bits], np.int64)
    prefix = np.concatenate([[0], np.cumsum(pop)[:-1]]).astype(np.int32)
    v_out = out_idx.shape[0]
    kvol = ks[0] * ks[1] * ks[2]
    stride_rows = max((v_out + 127) // 128 * 128, 128)
    nbr = np.full((kvol, stride_rows), -7, np.int32)
    masks = np.zeros(v_out, np.uint32)
    out_c = np.ascontiguousarray(out_idx, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.check(lib.isf_debug_nbr_rows_host(p(out_c), v_out, p(bits), p(prefix), L.host_i32(in_shape), L.host_i32(ks), L.host_i32(st),
                                        L.host_i32(pd), L.host_i32([1, 1, 1]), 1 if subm else 0, p(nbr), ctypes.c_longlong(stride_rows),
                                        p(masks)))
    return nbr[:, :v_out], masks


@pytest.mark.parametrize('ks,st,pd,subm', [([3, 3, 3], [1, 1, 1], [1, 1, 1], True), ([3, 3, 3], [2, 2, 2], [1, 1, 1], False),
                                           ([3, 3, 3], [2, 2, 2], [0, 1, 1], False), ([3, 1, 1], [2, 1, 1], [0, 0, 0], False)])
def test_row_walk_device_logic_matches_oracle_rulebook_on_host(ks, st, pd, subm):
    """The source of nbr_bitmap_rows_kernel's inner walk, compiled for the HOST (isf_debug_nbr_rows_host), against the
    oracle's rulebook (itself pinned by the reference's spconv C++): neighbour rows and occupancy masks, incl. offsets
    that straddle a 32-cell bitmap word and cells on the grid border."""
    import tile_occupancy as to
    from oracle import spconv_ops as so
    rs = np.random.RandomState(3)
    shape, batch = [11, 37, 45], 2                      # x extent not a multiple of 32: words straddle rows
    cells = batch * shape[0] * shape[1] * shape[2]
    lin = np.sort(rs.permutation(cells)[:6000])
    idx = np.stack([lin // (shape[0] * shape[1] * shape[2]), (lin // (shape[1] * shape[2])) % shape[0],
                    (lin // shape[2]) % shape[1], lin % shape[2]], 1).astype(np.int32)
    oi, oshape, want = so.get_indice_pairs(torch.from_numpy(idx), batch, shape, ks, st, pd, [1, 1, 1], subm)
    got, masks = _host_row_walk(idx, oi.numpy(), shape, batch, ks, st, pd, subm)
    assert np.array_equal(got, want.numpy().astype(np.int32))
    pos = to.library_tap_bits(tuple(ks))
    want_mask = np.zeros(got.shape[1], np.uint32)
    for k in range(got.shape[0]):
        want_mask |= (want.numpy()[k] >= 0).astype(np.uint32) << np.uint32(pos[k])
    assert np.array_equal(masks, want_mask)
