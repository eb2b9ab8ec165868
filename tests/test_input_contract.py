"""Input contract: multi-sweep point loading with the merge on the device.

CPU: the kernels' per-point source compiled for the host (isf_debug_merge_sweeps_host) and the product's
LoadPointsFromMultiSweeps (on the ABI emulator) against the reference's own loader run (tests/golden/ref_sweeps.npz).
GPU (bring-up, ISF_EXPERIMENTAL=1): the kernels against the same fixture and at 300 k points against the host-compiled code."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, 'golden'))
import golden_common as gc      # noqa: E402

experimental = pytest.mark.skipif(os.environ.get('ISF_EXPERIMENTAL') != '1', reason='bring-up test: set ISF_EXPERIMENTAL=1')
G = np.load(os.path.join(HERE, 'golden', 'ref_sweeps.npz'))
CASES = [('plain', dict(), 10), ('close', dict(remove_close=True), 10), ('test3', dict(test_mode=True, sweeps_num=3), 10),
         ('pad', dict(pad_empty_sweeps=True, remove_close=True, sweeps_num=2), 0)]


def _results(n_sw):
    key, ts, sweeps = gc.sweep_inputs()
    return dict(points=key, timestamp=ts, sweeps=sweeps[:n_sw])


def _host_merge(loader, results):
    """the plan of the product loader through the host-compiled kernel source."""
    from isfusion_b200 import _lib as L
    blocks, seg_start, flags, lag, rot, trans = loader.plan(results)
    raw = np.ascontiguousarray(np.concatenate(blocks, 0), dtype=np.float32)
    m, s = raw.shape[0], len(flags)
    out = np.zeros((m, 5), np.float32)
    count = ctypes.c_int(0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    dbl = lambda a: (ctypes.c_double * a.size)(*[float(v) for v in np.asarray(a, dtype=np.float64).reshape(-1)])
    L.check(L.lib().isf_debug_merge_sweeps_host(p(raw), m, 5, s, L.host_i32(seg_start), dbl(rot), dbl(trans), dbl(np.array(lag)),
                                                L.host_i32(flags), ctypes.c_float(1.0), p(out), ctypes.byref(count)))
    return out[:count.value]


@pytest.mark.parametrize('tag,kw,n_sw', CASES)
def test_merge_kernel_source_on_host_matches_the_reference_loader(tag, kw, n_sw):
    from isfusion_b200.sweeps import LoadPointsFromMultiSweeps
    loader = LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], device='cpu', **kw)
    got = _host_merge(loader, _results(n_sw))
    want = G[tag]
    assert got.shape == want.shape
    # float64 products rounded to float32 twice, as numpy does: bit-exact up to the summation order of the BLAS kernel
    # (a last-bit difference of a float64 sum can move a float32 rounding); allow <= 1 ulp on < 1e-4 of the values
    diff = got != want
    assert diff.mean() < 1e-4, diff.mean()
    assert np.allclose(got, want, rtol=2e-7, atol=1e-6)
    assert np.array_equal(got[:, 3:], want[:, 3:])              # intensity and time lag: exact


@pytest.mark.parametrize('tag,kw,n_sw', CASES)
def test_product_loader_on_the_emulator(tag, kw, n_sw, monkeypatch):
    import abi_emulator
    from isfusion_b200.registry import PIPELINES
    abi_emulator.install(monkeypatch)
    loader = PIPELINES.build(dict(type='LoadPointsFromMultiSweeps', use_dim=[0, 1, 2, 3, 4], device='cpu', **kw))
    res = loader(_results(n_sw))
    assert np.array_equal(res['points'].numpy(), G[tag])
    # use_dim selects columns like the reference (:903)
    loader = PIPELINES.build(dict(type='LoadPointsFromMultiSweeps', device='cpu', **kw))          # default use_dim [0, 1, 2, 4]
    assert np.array_equal(loader(_results(n_sw))['points'].numpy(), G[tag][:, [0, 1, 2, 4]])


def test_train_mode_draws_sweeps_like_the_reference(monkeypatch):
    """more sweeps than sweeps_num outside test mode: np.random.choice without replacement (:877-878), same stream."""
    import abi_emulator
    from isfusion_b200.sweeps import LoadPointsFromMultiSweeps
    from oracle import input_ops
    abi_emulator.install(monkeypatch)
    key, ts, sweeps = gc.sweep_inputs()
    np.random.seed(5)
    got = LoadPointsFromMultiSweeps(sweeps_num=4, use_dim=[0, 1, 2, 3, 4], device='cpu')(dict(points=key, timestamp=ts, sweeps=sweeps))
    np.random.seed(5)
    want = input_ops.load_points_from_multi_sweeps(key, ts, sweeps, sweeps_num=4)
    assert np.array_equal(got['points'].numpy(), want)


@pytest.mark.gpu
@experimental
@pytest.mark.parametrize('tag,kw,n_sw', CASES)
def test_merge_sweeps_kernels_against_the_reference_loader(tag, kw, n_sw):
    from isfusion_b200.sweeps import LoadPointsFromMultiSweeps
    loader = LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], **kw)
    got = loader(_results(n_sw))['points'].cpu().numpy()
    host = _host_merge(LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], device='cpu', **kw), _results(n_sw))
    assert np.array_equal(got, host)                              # device == the same source compiled for the host
    assert got.shape == G[tag].shape and np.allclose(got, G[tag], rtol=2e-7, atol=1e-6)


@pytest.mark.gpu
@experimental
def test_merge_sweeps_full_size():
    """a nuScenes-sized frame: key frame + 10 sweeps of ~27 k points (300 k points), with and without the close filter."""
    from isfusion_b200.sweeps import LoadPointsFromMultiSweeps
    key, ts, sweeps = gc.sweep_inputs(seed=3, n_key=30000, n_sweeps=10, n_sweep=27000)
    for rc in (False, True):
        res = dict(points=key, timestamp=ts, sweeps=sweeps)
        got = LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], remove_close=rc)(dict(res))['points'].cpu().numpy()
        host = _host_merge(LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], device='cpu', remove_close=rc), dict(res))
        assert np.array_equal(got, host)
