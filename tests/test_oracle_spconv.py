"""CPU: the sparse-conv oracle against an independent dense F.conv3d on small grids."""
import pytest
import torch

from oracle import spconv_ops as so


def _rand_tensor(shape, nvox, c, batch, seed):
    g = torch.Generator().manual_seed(seed)
    cells = batch * shape[0] * shape[1] * shape[2]
    lin = torch.randperm(cells, generator=g)[:nvox]
    idx = torch.stack([lin // (shape[0] * shape[1] * shape[2]), (lin // (shape[1] * shape[2])) % shape[0],
                       (lin // shape[2]) % shape[1], lin % shape[2]], 1).int()
    return so.SparseConvTensor(torch.randn(nvox, c, generator=g), idx, shape, batch)


@pytest.mark.parametrize('stride,padding,ksize', [(1, 1, 3), (2, 1, 3), (2, (0, 1, 1), 3), ((2, 1, 1), 0, (3, 1, 1))])
def test_sparse_conv_matches_dense(stride, padding, ksize):
    x = _rand_tensor([9, 14, 12], 300, 8, 2, 0)
    conv = so.SparseConv3d(8, 16, ksize, stride=stride, padding=padding, bias=False)
    out = conv(x)
    dense = so.dense_conv_reference(x, conv)
    assert list(dense.shape[2:]) == out.spatial_shape
    got = out.dense()
    assert torch.allclose(got, dense, atol=1e-4)
    # active set = every output cell with at least one active input in its receptive field
    ones = so.SparseConvTensor(torch.ones(x.features.shape[0], 1), x.indices, x.spatial_shape, x.batch_size)
    cnt = so.SparseConv3d(1, 1, ksize, stride=stride, padding=padding, bias=False)
    cnt.weight.data.fill_(1.0)
    reach = so.dense_conv_reference(ones, cnt)[:, 0] > 0
    act = torch.zeros_like(reach)
    i = out.indices.long()
    act[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] = True
    assert torch.equal(act, reach)
    # canonical row order: ascending linear index
    key = so._lin(out.indices, out.spatial_shape)
    assert (key[1:] > key[:-1]).all()


def test_subm_conv_matches_masked_dense():
    x = _rand_tensor([7, 10, 11], 250, 6, 2, 1)
    conv = so.SubMConv3d(6, 10, 3, padding=1, bias=False)
    out = conv(x)
    assert torch.equal(out.indices, x.indices)
    dense = so.dense_conv_reference(x, conv)
    i = x.indices.long()
    assert torch.allclose(out.features, dense[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]], atol=1e-4)


def test_sparse_encoder_shapes():
    """ref: tests/test_models/test_common_modules/test_middle_encoders.py:8-27 (shape contract), with the
    isfusion_0075voxel.py:72-82 channel plan on a reduced grid."""
    torch.manual_seed(0)
    enc = so.SparseEncoder(in_channels=64, sparse_shape=[41, 64, 64], base_channels=32, output_channels=256,
                           encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
                           encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock').eval()
    x = _rand_tensor([41, 64, 64], 400, 64, 2, 2)
    sp, feats, _ = enc(x.features, x.indices, 2)
    assert sp.shape == (2, 512, 8, 8)
    assert [f.spatial_shape for f in feats] == [[41, 64, 64], [21, 32, 32], [11, 16, 16], [5, 8, 8], [5, 8, 8]]
    assert set(k for k in enc.state_dict()) >= {'conv_input.0.weight', 'encoder_layers.encoder_layer1.0.conv1.weight',
                                                'encoder_layers.encoder_layer1.2.0.weight', 'conv_out.1.running_mean'}


# ---------------------------------------------------------------- pinned by the reference's vendored spconv-1.x C++
def _oracle_canonical(idx, shape, batch, case):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden as mg
    oi, oshape, nbr = so.get_indice_pairs(torch.from_numpy(idx), batch, shape, case['ksize'], case['stride'], case['padding'],
                                          (1, 1, 1), case['subm'])
    k, o = torch.nonzero(nbr >= 0, as_tuple=True)
    trip = torch.stack([k, nbr[k, o], o], 1).numpy()
    return mg.canonical_rulebook(oi.numpy(), trip, oshape) + (list(oshape),)


def test_rulebook_matches_reference_spconv_cpp_golden():
    """Output sites and (kernel offset, input row, output site) pairs of oracle.get_indice_pairs == the reference's
    getIndicePairsConv / getIndicePairsSubM (geometry.h:144-297), for the four conv shapes of the encoder.  Row ORDER
    is implementation-defined in the reference (first-touch on CPU, sorted-unique on GPU), so both are canonicalised."""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden as mg
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'spconv_ref_cpp.npz'))
    for name, case in mg.SPCONV_CASES.items():
        co, ct, oshape = _oracle_canonical(g['indices'], mg.SPCONV_SHAPE, mg.SPCONV_BATCH, case)
        assert oshape == g[name + '_shape'].tolist(), name
        assert np.array_equal(co, g[name + '_out']), name
        assert np.array_equal(ct, g[name + '_pairs']), name


def test_rulebook_and_conv_match_reference_spconv_cpp_live():
    """Same check on fresh seeds against oracle/_ref/libref_spconv.so (built from the reference header where it lies),
    plus the arithmetic: gather -> mm -> scatter-add over the REFERENCE's pairs == oracle.indice_conv."""
    import numpy as np
    from oracle import ref_spconv
    if not ref_spconv.available():
        pytest.skip('oracle/_ref/libref_spconv.so not built (needs /root/reference)')
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden as mg
    shape, batch = [13, 16, 22], 3
    for seed in (1, 2):
        x = _rand_tensor(shape, 900, 6, batch, seed)
        idx = x.indices.numpy()
        for name, case in mg.SPCONV_CASES.items():
            oi, oshape, pairs, num = ref_spconv.indice_pairs(idx, batch, shape, case['ksize'], case['stride'], case['padding'],
                                                             (1, 1, 1), case['subm'])
            trip = [(k, pairs[k, 0, j], pairs[k, 1, j]) for k in range(pairs.shape[0]) for j in range(num[k])]
            rco, rct = mg.canonical_rulebook(oi, trip, oshape)
            co, ct, osh = _oracle_canonical(idx, shape, batch, case)
            assert osh == oshape and np.array_equal(co, rco) and np.array_equal(ct, rct), (seed, name)
            # arithmetic over the reference's own pairs (spconv_ops.h:260-361), rows then sorted like the oracle's
            conv = (so.SubMConv3d if case['subm'] else so.SparseConv3d)(6, 10, case['ksize'], stride=case['stride'],
                                                                          padding=case['padding'], bias=False)
            w = conv.weight.detach().reshape(-1, 6, 10)
            ref = torch.zeros(len(oi), 10)
            for k in range(pairs.shape[0]):
                n = int(num[k])
                if n:
                    ref.index_add_(0, torch.from_numpy(pairs[k, 1, :n].astype(np.int64)),
                                   x.features[torch.from_numpy(pairs[k, 0, :n].astype(np.int64))] @ w[k])
            out = conv(x)
            key = so._lin(torch.from_numpy(np.asarray(oi)), oshape)
            okey = so._lin(out.indices, out.spatial_shape)
            assert torch.equal(torch.sort(key).values, torch.sort(okey).values)
            assert torch.allclose(ref[torch.argsort(key)], out.features.detach()[torch.argsort(okey)], atol=1e-5), (seed, name)
