"""Input contract: multi-sweep point loading with the merge on the device.

CPU: the kernels' per-point source compiled for the host (isf_debug_merge_sweeps_host) and the product's
LoadPointsFromMultiSweeps (on the ABI emulator) against the reference's own loader run (tests/golden/ref_sweeps.npz).
GPU: the kernels against the same fixture and at 300 k points against the host-compiled code."""
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

G = np.load(os.path.join(HERE, 'golden', 'ref_sweeps.npz'))
CASES = [('plain', dict(), 10), ('close', dict(remove_close=True), 10), ('test3', dict(test_mode=True, sweeps_num=3), 10),
         ('pad', dict(pad_empty_sweeps=True, remove_close=True, sweeps_num=2), 0)]


def _results(n_sw):
    key, ts, sweeps = gc.sweep_inputs()
    return dict(points=key, timestamp=ts, sweeps=sweeps[:n_sw])


def _host_merge(loader, results, rng=None):
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
                                                L.host_i32(flags), ctypes.c_float(1.0), L.host_f32(rng) if rng is not None else None, p(out),
                                                ctypes.byref(count)))
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
    import isfusion_b200.sweeps  # noqa: F401  (registers the loader: the test must not depend on an earlier import)
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
@pytest.mark.parametrize('tag,kw,n_sw', CASES)
def test_merge_sweeps_kernels_against_the_reference_loader(tag, kw, n_sw):
    from isfusion_b200.sweeps import LoadPointsFromMultiSweeps
    loader = LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], **kw)
    got = loader(_results(n_sw))['points'].cpu().numpy()
    host = _host_merge(LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], device='cpu', **kw), _results(n_sw))
    assert np.array_equal(got, host)                              # device == the same source compiled for the host
    assert got.shape == G[tag].shape and np.allclose(got, G[tag], rtol=2e-7, atol=1e-6)


@pytest.mark.gpu
def test_merge_sweeps_full_size():
    """a nuScenes-sized frame: key frame + 10 sweeps of ~27 k points (300 k points), with and without the close filter."""
    from isfusion_b200.sweeps import LoadPointsFromMultiSweeps
    key, ts, sweeps = gc.sweep_inputs(seed=3, n_key=30000, n_sweeps=10, n_sweep=27000)
    for rc in (False, True):
        res = dict(points=key, timestamp=ts, sweeps=sweeps)
        got = LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], remove_close=rc)(dict(res))['points'].cpu().numpy()
        host = _host_merge(LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], device='cpu', remove_close=rc), dict(res))
        assert np.array_equal(got, host)


# ------------------------------------------------------------------------------------------------ camera side
GI = np.load(os.path.join(HERE, 'golden', 'ref_images.npz'))
IMG_CASES = [('test', None), ('train', 4), ('train2', 9)]


def _aug_cfg(seed):
    return dict(gc.IMG_AUG_TEST) if seed is None else dict(gc.IMG_AUG_TEST, resize_lim=[0.57, 0.825], rand_flip=True, is_train=True)


@pytest.mark.parametrize('tag,seed', IMG_CASES)
def test_image_kernel_source_on_host_matches_the_reference_pipeline(tag, seed):
    """plan + both passes (Pillow's integer resampler, crop, flip, ToTensor + Normalize) with the kernels' per-pixel source
    compiled for the host, against the reference's ImageAug3D + ImageNormalize run: bit-exact float32."""
    from isfusion_b200 import _lib as L
    from isfusion_b200.image_pipeline import ImageAug3D
    imgs = gc.camera_images()
    h, w = imgs[0].shape[:2]
    if seed is not None:
        np.random.seed(seed)
    data = ImageAug3D(**_aug_cfg(seed))(dict(img=list(imgs), ori_shape=(w, h)))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for i, (img, (dims, crop, flip)) in enumerate(zip(imgs, data['img_aug_plan'])):
        out = np.zeros((1, 3, crop[3] - crop[1], crop[2] - crop[0]), np.float32)
        L.check(L.lib().isf_debug_image_aug_host(p(np.ascontiguousarray(img)), 1, w, h, dims[0], dims[1], crop[0], crop[1], crop[2] - crop[0],
                                                 crop[3] - crop[1], int(flip), L.host_f32(gc.IMG_NORM['mean']), L.host_f32(gc.IMG_NORM['std']),
                                                 p(out)))
        assert np.array_equal(out[0], GI[tag + '_img'][i]), (tag, i)
        assert np.array_equal(data['img_aug_matrix'][i], GI[tag + '_aug'][i])


@pytest.mark.parametrize('tag,seed', IMG_CASES)
def test_product_image_pipeline_on_the_emulator(tag, seed, monkeypatch):
    """ImageAug3D -> ImageNormalize built by name from the config's dicts, PIL images in (as LoadMultiViewImageFromFilesV2
    delivers them), device tensors out: same draws, same matrices, same pixels as the reference run."""
    PIL = pytest.importorskip('PIL.Image')
    import abi_emulator
    from isfusion_b200.registry import PIPELINES
    import isfusion_b200.image_pipeline as IP
    abi_emulator.install(monkeypatch)
    monkeypatch.setattr(IP, '_PLANS', {})
    imgs = gc.camera_images()
    h, w = imgs[0].shape[:2]
    aug = PIPELINES.build(dict(type='ImageAug3D', **_aug_cfg(seed)))
    norm = PIPELINES.build(dict(type='ImageNormalize', device='cpu', **gc.IMG_NORM))
    if seed is not None:
        np.random.seed(seed)
    data = norm(aug(dict(img=[PIL.fromarray(i) for i in imgs], ori_shape=(w, h))))
    assert data['img_norm_cfg'] == dict(gc.IMG_NORM) and 'img_aug_plan' not in data
    for i, t in enumerate(data['img']):
        assert t.dtype == torch.float32 and np.array_equal(t.numpy(), GI[tag + '_img'][i])
        assert np.array_equal(data['img_aug_matrix'][i], GI[tag + '_aug'][i])
    # fp16 output for the backbone = the float32 result rounded once
    norm16 = PIPELINES.build(dict(type='ImageNormalize', device='cpu', out_dtype=torch.float16, **gc.IMG_NORM))
    if seed is not None:
        np.random.seed(seed)
    d16 = norm16(aug(dict(img=list(imgs), ori_shape=(w, h))))
    assert d16['img'][0].dtype == torch.float16 and torch.equal(d16['img'][0], torch.from_numpy(GI[tag + '_img'][0]).half())


def test_image_plan_full_size_against_pillow():
    """the config's real geometry (1600x900 -> resize 0.72 -> crop to 384x1056): the host-compiled kernel source against
    PIL + the float32 normalisation, bit-exact; and the plan only asks for the input rows below the crop line."""
    PIL = pytest.importorskip('PIL.Image')
    from isfusion_b200 import _lib as L
    from isfusion_b200.image_pipeline import ImageAug3D
    rs = np.random.RandomState(2)
    img = rs.randint(0, 256, (900, 1600, 3)).astype(np.uint8)
    aug = ImageAug3D(final_dim=(384, 1056), resize_lim=[0.72, 0.72], bot_pct_lim=[0.0, 0.0], rot_lim=[0.0, 0.0], rand_flip=False, is_train=False)
    resize, dims, crop, flip, rot = aug.sample_augmentation(dict(ori_shape=(1600, 900)))
    assert dims == (1152, 648) and crop == (48, 264, 1104, 648)
    ref = np.asarray(PIL.fromarray(img).resize(dims).crop(crop)).transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    ref = (ref - np.float32(gc.IMG_NORM['mean'])[:, None, None]) / np.float32(gc.IMG_NORM['std'])[:, None, None]
    out = np.zeros((1, 3, 384, 1056), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.check(L.lib().isf_debug_image_aug_host(p(img), 1, 1600, 900, 1152, 648, 48, 264, 1056, 384, 0, L.host_f32(gc.IMG_NORM['mean']),
                                             L.host_f32(gc.IMG_NORM['std']), p(out)))
    assert np.array_equal(out[0], ref)


@pytest.mark.gpu
@pytest.mark.parametrize('tag,seed', IMG_CASES)
def test_image_kernels_against_the_reference_pipeline(tag, seed):
    from isfusion_b200.image_pipeline import ImageAug3D, ImageNormalize
    imgs = gc.camera_images()
    h, w = imgs[0].shape[:2]
    if seed is not None:
        np.random.seed(seed)
    data = ImageNormalize(**gc.IMG_NORM)(ImageAug3D(**_aug_cfg(seed))(dict(img=list(imgs), ori_shape=(w, h))))
    for i, t in enumerate(data['img']):
        assert np.array_equal(t.cpu().numpy(), GI[tag + '_img'][i]), (tag, i)


@pytest.mark.gpu
def test_image_kernels_full_size_six_cameras():
    """6 x 1600x900 frames through the config's test-time geometry: device == the same source compiled for the host."""
    from isfusion_b200 import _lib as L
    from isfusion_b200.image_pipeline import image_aug_normalize
    rs = np.random.RandomState(3)
    imgs = rs.randint(0, 256, (6, 900, 1600, 3)).astype(np.uint8)
    got = image_aug_normalize(torch.from_numpy(imgs).cuda(), (1152, 648), (48, 264, 1104, 648), False, gc.IMG_NORM['mean'], gc.IMG_NORM['std'])
    out = np.zeros((6, 3, 384, 1056), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.check(L.lib().isf_debug_image_aug_host(p(imgs), 6, 1600, 900, 1152, 648, 48, 264, 1056, 384, 0, L.host_f32(gc.IMG_NORM['mean']),
                                             L.host_f32(gc.IMG_NORM['std']), p(out)))
    assert np.array_equal(got.cpu().numpy(), out)
    half = image_aug_normalize(torch.from_numpy(imgs).cuda(), (1152, 648), (48, 264, 1104, 648), False, gc.IMG_NORM['mean'], gc.IMG_NORM['std'],
                               out_dtype=torch.float16)
    assert torch.equal(half.cpu(), torch.from_numpy(out).half())


def test_inference_pipeline_assembled_from_config_dicts(monkeypatch):
    """the config's test-time transforms that touch the path's inputs (config:305-343), built by name and composed: sweeps merged,
    cameras resized / cropped / normalised, identity lidar_aug_matrix, points strictly inside the range — against the oracle
    statements of the same steps."""
    import abi_emulator
    from oracle import input_ops
    import isfusion_b200.image_pipeline  # noqa: F401  (registers the transforms)
    import isfusion_b200.image_pipeline as IP
    from isfusion_b200.sweeps import Compose
    abi_emulator.install(monkeypatch)
    monkeypatch.setattr(IP, '_PLANS', {})
    pcr = [-54, -54, -5, 54, 54, 3]
    pipe = Compose([
        dict(type='LoadPointsFromMultiSweeps', sweeps_num=10, use_dim=[0, 1, 2, 3, 4], device='cpu'),
        dict(type='ImageAug3D', **gc.IMG_AUG_TEST),
        dict(type='ImageNormalize', device='cpu', **gc.IMG_NORM),
        dict(type='GlobalRotScaleTransV2', resize_lim=[1.0, 1.0], rot_lim=[0.0, 0.0], trans_lim=0.0, is_train=False),
        dict(type='PointsRangeFilter', point_cloud_range=pcr),
    ])
    key, ts, sweeps = gc.sweep_inputs()
    key[:5, 0] = [54.0, -54.0, 60.0, 53.999, -53.999]            # on / outside / just inside the range faces
    imgs = gc.camera_images()
    h, w = imgs[0].shape[:2]
    data = pipe(dict(points=key.copy(), timestamp=ts, sweeps=sweeps, img=list(imgs), ori_shape=(w, h)))
    want = input_ops.load_points_from_multi_sweeps(key, ts, sweeps)
    r = np.array(pcr, np.float32)
    keep = (want[:, 0] > r[0]) & (want[:, 1] > r[1]) & (want[:, 2] > r[2]) & (want[:, 0] < r[3]) & (want[:, 1] < r[4]) & (want[:, 2] < r[5])
    assert 0 < keep.sum() < len(want) and not keep[0] and not keep[1] and not keep[2] and keep[3] and keep[4]
    assert np.array_equal(data['points'].numpy(), want[keep])
    assert np.array_equal(data['lidar_aug_matrix'], np.eye(4, dtype=np.float32))
    assert np.array_equal(torch.stack(data['img']).numpy(), GI['test_img']) and np.array_equal(np.stack(data['img_aug_matrix']), GI['test_aug'])


def test_points_range_filter_matches_the_reference_point_class():
    """PointsRangeFilter (device-side mask) against the reference's own BasePoints.in_range_3d evaluated on the merged cloud."""
    from isfusion_b200.sweeps import PointsRangeFilter
    pts = torch.from_numpy(G['plain'])
    got = PointsRangeFilter([-40, -40, -3, 40, 40, 1])(dict(points=pts))['points']
    mask = G['plain_in_range40']
    assert 0 < mask.sum() < len(mask) and np.array_equal(got.numpy(), G['plain'][mask])


def test_file_loaders_feed_the_pipeline(tmp_path, monkeypatch):
    """the config's test pipeline from FILE NAMES (config:297-332): .bin sweeps + PNG cameras on disk -> the same points /
    images as from arrays."""
    PIL = pytest.importorskip('PIL.Image')
    import abi_emulator
    import isfusion_b200.image_pipeline as IP
    from isfusion_b200.sweeps import Compose
    abi_emulator.install(monkeypatch)
    monkeypatch.setattr(IP, '_PLANS', {})
    key, ts, sweeps = gc.sweep_inputs()
    key.tofile(str(tmp_path / 'key.bin'))
    infos = []
    for i, s in enumerate(sweeps):
        s['points'].tofile(str(tmp_path / ('s%d.bin' % i)))
        infos.append(dict(data_path=str(tmp_path / ('s%d.bin' % i)), timestamp=s['timestamp'],
                          sensor2lidar_rotation=s['sensor2lidar_rotation'], sensor2lidar_translation=s['sensor2lidar_translation']))
    names = []
    for i, img in enumerate(gc.camera_images()):
        PIL.fromarray(img).save(str(tmp_path / ('cam%d.png' % i)))          # lossless, so the pixels are the fixture's
        names.append(str(tmp_path / ('cam%d.png' % i)))
    pipe = Compose([
        dict(type='LoadPointsFromFile', coord_type='LIDAR', load_dim=5, use_dim=5, painting=False),
        dict(type='LoadPointsFromMultiSweeps', sweeps_num=10, use_dim=[0, 1, 2, 3, 4], painting=False, device='cpu'),
        dict(type='LoadMultiViewImageFromFilesV2', to_float32=True),
        dict(type='ImageAug3D', **gc.IMG_AUG_TEST),
        dict(type='ImageNormalize', device='cpu', **gc.IMG_NORM),
    ])
    data = pipe(dict(pts_filename=str(tmp_path / 'key.bin'), timestamp=ts, sweeps=infos, img_filename=names))
    assert np.array_equal(data['points'].numpy(), G['plain'])
    assert data['ori_shape'] == (400, 225) and np.array_equal(torch.stack(data['img']).numpy(), GI['test_img'])


@pytest.mark.parametrize('fw', [263, 262, 264])
def test_image_kernel_source_scalar_and_word_paths_against_the_oracle(fw):
    """row pitches that are / are not a multiple of 4 bytes take the word-vectorised / the scalar vertical pass: both equal
    the oracle (itself pinned by the reference run and by Pillow)."""
    from oracle import input_ops
    from isfusion_b200 import _lib as L
    img = gc.camera_images()[1]
    h, w = img.shape[:2]
    dims, crop = input_ops.image_aug_geometry((w, h), (90, fw), 0.72)
    for flip in (False, True):
        want = input_ops.image_aug_normalize(img, dims, crop, flip, **gc.IMG_NORM)
        out = np.zeros((1, 3, 90, fw), np.float32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        L.check(L.lib().isf_debug_image_aug_host(p(np.ascontiguousarray(img)), 1, w, h, dims[0], dims[1], crop[0], crop[1], fw, 90, int(flip),
                                                 L.host_f32(gc.IMG_NORM['mean']), L.host_f32(gc.IMG_NORM['std']), p(out)))
        assert np.array_equal(out[0], want), (fw, flip)


def test_files_to_boxes_through_the_reference_entry_point(tmp_path, monkeypatch):
    """the whole inference story on the CPU emulator: sweep / camera FILES -> the config's test-time transforms (built by name)
    -> Collect3DV2 -> collate -> `detector(return_loss=False, **batch)` (the reference's call, detectors/base.py:46-62) ->
    decoded boxes.  The detector is the config's whole model (synth.full_model_config)."""
    PIL = pytest.importorskip('PIL.Image')
    import abi_emulator
    import synth
    import isfusion_b200.head as H
    import isfusion_b200.img_neck as N
    import isfusion_b200.image_pipeline as IP
    from isfusion_b200.sweeps import Compose, collate
    from __graft_entry__ import make_full_net
    abi_emulator.install(monkeypatch)
    monkeypatch.setattr(IP, '_PLANS', {})
    monkeypatch.setattr(H, '_conv', abi_emulator._conv_cpu)
    monkeypatch.setattr(N, '_conv', abi_emulator._conv_cpu)
    key, ts, sweeps = gc.sweep_inputs(n_key=1500, n_sweeps=3, n_sweep=800)
    key.tofile(str(tmp_path / 'key.bin'))
    infos = []
    for i, s in enumerate(sweeps):
        s['points'].tofile(str(tmp_path / ('s%d.bin' % i)))
        infos.append(dict(data_path=str(tmp_path / ('s%d.bin' % i)), timestamp=s['timestamp'],
                          sensor2lidar_rotation=s['sensor2lidar_rotation'], sensor2lidar_translation=s['sensor2lidar_translation']))
    names = []
    for i, img in enumerate(gc.camera_images(n=6, hw=(90, 160))):
        PIL.fromarray(img).save(str(tmp_path / ('cam%d.png' % i)))
        names.append(str(tmp_path / ('cam%d.png' % i)))
    pipe = Compose([
        dict(type='LoadPointsFromFile', coord_type='LIDAR', load_dim=5, use_dim=5, painting=False),
        dict(type='LoadPointsFromMultiSweeps', sweeps_num=10, use_dim=[0, 1, 2, 3, 4], painting=False, device='cpu'),
        dict(type='LoadMultiViewImageFromFilesV2', to_float32=True),
        dict(type='ImageAug3D', final_dim=(64, 96), resize_lim=[0.72, 0.72], bot_pct_lim=[0.0, 0.0], rot_lim=[0.0, 0.0], rand_flip=False,
             is_train=False),
        dict(type='ImageNormalize', device='cpu', out_dtype=torch.float16, **gc.IMG_NORM),
        dict(type='GlobalRotScaleTransV2', resize_lim=[1.0, 1.0], rot_lim=[0.0, 0.0], trans_lim=0.0, is_train=False),
        dict(type='PointsRangeFilter', point_cloud_range=[-54, -54, -5, 54, 54, 3]),
        dict(type='Collect3DV2', keys=['points', 'img'],
             meta_keys=['camera_intrinsics', 'camera2ego', 'lidar2ego', 'lidar2camera', 'camera2lidar', 'lidar2img', 'img_aug_matrix',
                        'lidar_aug_matrix']),
    ])
    l2i = synth.make_calib(1, img_hw=(64, 96))[0][0]                       # [6, 4, 4] of the sample
    sample = pipe(dict(pts_filename=str(tmp_path / 'key.bin'), timestamp=ts, sweeps=infos, img_filename=names, lidar2img=list(l2i.numpy())))
    assert tuple(sample['img'].shape) == (6, 3, 64, 96) and sample['img'].dtype == torch.float16
    assert tuple(sample['img_aug_matrix'].shape) == (6, 4, 4) and tuple(sample['lidar_aug_matrix'].shape) == (4, 4)
    assert sample['img_metas']['ori_shape'] == (160, 90) and 'img_norm_cfg' in sample['img_metas']
    batch = collate([sample])
    assert tuple(batch['img'][0].shape) == (1, 6, 3, 64, 96) and tuple(batch['lidar2img'].shape) == (1, 6, 4, 4)
    net = make_full_net(0).eval()
    res = net(return_loss=False, **batch)
    box = res[0]['pts_bbox']
    assert box['boxes_3d'].shape == (200, 9) and torch.isfinite(box['boxes_3d'].float()).all()
    assert batch['img_metas'][0][0]['input_shape'] == (64, 96)            # recorded by extract_img_feat for the projection


@pytest.mark.parametrize('rc', [False, True])
def test_range_filter_fused_into_the_merge(rc, monkeypatch):
    """point_cloud_range on the loader = the pipeline's PointsRangeFilter folded into the merge kernel: the kernel source on the
    host and the product loader on the emulator give the reference loader's cloud filtered by the reference's in_range_3d."""
    import abi_emulator
    from isfusion_b200.sweeps import LoadPointsFromMultiSweeps
    rng = [-40, -40, -3, 40, 40, 1]
    tag = 'close' if rc else 'plain'
    r = np.array(rng, np.float32)
    c = G[tag]
    want = c[(c[:, 0] > r[0]) & (c[:, 1] > r[1]) & (c[:, 2] > r[2]) & (c[:, 0] < r[3]) & (c[:, 1] < r[4]) & (c[:, 2] < r[5])]
    if not rc:
        assert np.array_equal(want, G['plain'][G['plain_in_range40']])          # = the reference's own mask
    got = _host_merge(LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], device='cpu', remove_close=rc), _results(10), rng)
    assert np.array_equal(got, want)
    abi_emulator.install(monkeypatch)
    res = LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], device='cpu', remove_close=rc, point_cloud_range=rng)(_results(10))
    assert np.array_equal(res['points'].numpy(), want)
