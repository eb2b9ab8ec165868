"""Oracle: the input contract of the path — multi-sweep point loading.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED by outputs of the reference's own
mmdet3d/datasets/pipelines/loading.py (LoadPointsFromFile + LoadPointsFromMultiSweeps with core/points/*.py) run in the
build container through tests/golden/ref_shim.py on synthetic sweep files (tests/golden/ref_sweeps.npz,
tests/test_oracle_vs_reference.py).  numpy restatement, same dtypes as the reference at every step (float32 sweeps,
float64 calibration, results cast back to float32).
"""
import numpy as np


def remove_close(points, radius=1.0):
    """ref: loading.py:824-844: drop points with |x| < radius AND |y| < radius."""
    close = np.logical_and(np.abs(points[:, 0]) < radius, np.abs(points[:, 1]) < radius)
    return points[np.logical_not(close)]


def load_points_from_multi_sweeps(key_points, ts, sweeps, sweeps_num=10, load_dim=5, use_dim=(0, 1, 2, 3, 4),
                                  pad_empty_sweeps=False, remove_close_pts=False, test_mode=False, rng=None):
    """ref: LoadPointsFromMultiSweeps.__call__ (loading.py:846-906).

    key_points [N0, load_dim] float32 (the key frame as LoadPointsFromFile delivers it); ts = results['timestamp'] (s);
    sweeps = list of dict(points [Ni, load_dim] float32 raw, timestamp (us), sensor2lidar_rotation [3,3] float64,
    sensor2lidar_translation [3] float64).  Returns [N, len(use_dim)] float32: key frame (time lag 0, :862) followed by the
    chosen sweeps in order, each rotated + translated into the key frame's LiDAR frame (:886-888; float64 arithmetic,
    stored as float32 after the rotation and again after the translation) with column 4 = ts - sweep_ts/1e6 (:889)."""
    points = np.array(key_points, dtype=np.float32, copy=True)
    points[:, 4] = 0
    out = [points]
    if pad_empty_sweeps and len(sweeps) == 0:                          # :865-870
        for _ in range(sweeps_num):
            out.append(remove_close(points) if remove_close_pts else points)
    else:
        if len(sweeps) <= sweeps_num:                                  # :872-878
            choices = np.arange(len(sweeps))
        elif test_mode:
            choices = np.arange(sweeps_num)
        else:
            choices = (rng or np.random).choice(len(sweeps), sweeps_num, replace=False)
        for idx in choices:
            sweep = sweeps[idx]
            ps = np.copy(sweep['points']).reshape(-1, load_dim)
            if remove_close_pts:
                ps = remove_close(ps)
            sweep_ts = sweep['timestamp'] / 1e6
            ps[:, :3] = ps[:, :3] @ sweep['sensor2lidar_rotation'].T
            ps[:, :3] += sweep['sensor2lidar_translation']
            ps[:, 4] = ts - sweep_ts
            out.append(ps)
    return np.concatenate(out, 0)[:, list(use_dim)]


# ------------------------------------------------------------------------------------------------
# camera images: ImageAug3D (resize -> crop -> flip) + ImageNormalize  (transforms_3d.py:25-145)
# ------------------------------------------------------------------------------------------------
# The reference calls PIL: `img.resize(resize_dims)` (Pillow's default filter for resize() is BICUBIC, with the filter
# support widened by the down-scale factor), `img.crop(crop)`, optional FLIP_LEFT_RIGHT, `img.rotate(rotate)` (0 at test
# time), then torchvision ToTensor (uint8 HWC -> float32 CHW / 255) and Normalize ((x - mean) / std).  Pillow is a
# third-party dependency of the reference (not vendored; the build container has Pillow 12.2): its 8-bit resampler
# (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc)
# is restated here in numpy integer arithmetic and pinned against Pillow itself through the reference's own classes.
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1, np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


def resample_coeffs(in_size, out_size):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter over the whole axis ->
    (xmin [out], count [out], coef int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    coef = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = _bicubic((np.arange(n) + lo - center + 0.5) * ss)
        ww = w.sum()
        if ww != 0.0:
            w = w / ww
        k = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
        coef[xx, :n] = k.astype(np.int64).astype(np.int32)             # C cast: truncation toward zero
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, coef


def _resample_axis(img, xmin, cnt, coef, axis):
    """one 8-bit pass of Pillow's resampler along `axis` of a [H, W, C] uint8 image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)                    # [in, other, C]
    out = np.empty((len(xmin),) + src.shape[1:], np.uint8)
    for xx in range(len(xmin)):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for j in range(cnt[xx]):
            acc += src[xmin[xx] + j] * int(coef[xx, j])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bicubic(img, out_w, out_h):
    """PIL.Image.resize((out_w, out_h)) of a [H, W, 3] uint8 image: horizontal pass over the rows the vertical pass needs,
    rounded to uint8, then the vertical pass (ImagingResample)."""
    h, w = img.shape[:2]
    ymin, ycnt, ycoef = resample_coeffs(h, out_h)
    xmin, xcnt, xcoef = resample_coeffs(w, out_w)
    tmp = _resample_axis(img, xmin, xcnt, xcoef, 1) if out_w != w else img
    return _resample_axis(tmp, ymin, ycnt, ycoef, 0) if out_h != h else tmp


def image_aug_geometry(ori_wh, final_dim, resize, bot_pct=0.0, crop_w=None):
    """ImageAug3D.sample_augmentation, test branch (transforms_3d.py:71-79) / train branch with the random draws given."""
    W, H = ori_wh
    fH, fW = final_dim
    new_w, new_h = int(W * resize), int(H * resize)
    crop_h = int((1 - bot_pct) * new_h) - fH
    if crop_w is None:
        crop_w = int(max(0, new_w - fW) / 2)
    return (new_w, new_h), (crop_w, crop_h, crop_w + fW, crop_h + fH)


def image_aug_normalize(img, resize_dims, crop, flip, mean, std):
    """[H, W, 3] uint8 -> float32 [3, fH, fW]: img_transform with rotate = 0 (transforms_3d.py:82-90) + ImageNormalize (:25-43).
    A crop window that leaves the resized image reads zeros there, as PIL's crop does."""
    r = pil_resize_bicubic(img, resize_dims[0], resize_dims[1])
    x0, y0, x1, y1 = crop
    c = np.zeros((y1 - y0, x1 - x0, 3), np.uint8)
    sy0, sx0 = max(y0, 0), max(x0, 0)
    sy1, sx1 = min(y1, r.shape[0]), min(x1, r.shape[1])
    c[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = r[sy0:sy1, sx0:sx1]
    if flip:
        c = c[:, ::-1]
    t = c.transpose(2, 0, 1).astype(np.float32) / np.float32(255)                       # ToTensor
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return (t - m) / s                                                                   # Normalize (float32)


def image_aug_matrix(resize, crop, flip):
    """the 4x4 img_aug_matrix of ImageAug3D for rotate = 0 (transforms_3d.py:92-110,136-140), float32 like torch.Tensor."""
    rot = np.eye(2, dtype=np.float32) * np.float32(resize)
    tran = np.zeros(2, np.float32) - np.asarray(crop[:2], np.float32)
    if flip:
        A = np.array([[-1, 0], [0, 1]], np.float32)
        rot = A @ rot
        tran = A @ tran + np.array([crop[2] - crop[0], 0], np.float32)
    m = np.eye(4, dtype=np.float32)
    m[:2, :2] = rot
    m[:2, 3] = tran
    return m


def sample_augmentation(ori_wh, final_dim, resize_lim, bot_pct_lim, rot_lim, rand_flip, is_train, rng=np.random):
    """ImageAug3D.sample_augmentation (transforms_3d.py:57-80), both branches, the train branch drawing from `rng` in the
    reference's order -> (resize, resize_dims, crop, flip, rotate)."""
    W, H = ori_wh
    fH, fW = final_dim
    if is_train:
        resize = rng.uniform(*resize_lim)
        new_w, new_h = int(W * resize), int(H * resize)
        crop_h = int((1 - rng.uniform(*bot_pct_lim)) * new_h) - fH
        crop_w = int(rng.uniform(0, max(0, new_w - fW)))
        flip = bool(rand_flip and rng.choice([0, 1]))
        rotate = rng.uniform(*rot_lim)
    else:
        resize = np.mean(resize_lim)
        new_w, new_h = int(W * resize), int(H * resize)
        crop_h = int((1 - np.mean(bot_pct_lim)) * new_h) - fH
        crop_w = int(max(0, new_w - fW) / 2)
        flip, rotate = False, 0
    return resize, (new_w, new_h), (crop_w, crop_h, crop_w + fW, crop_h + fH), flip, rotate
