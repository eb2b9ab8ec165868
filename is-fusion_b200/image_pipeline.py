"""Input contract, camera side: `ImageAug3D` + `ImageNormalize` with the pixel work on the device.

Mirrors the reference's two pipeline classes (mmdet3d/datasets/pipelines/transforms_3d.py:25-145: same constructor
arguments, same `data` keys, same random draws, same `img_aug_matrix`).  The reference resizes / crops / flips every camera
with PIL on the host and normalises with torchvision; here `ImageAug3D` only SAMPLES the geometry (and records it), and
`ImageNormalize` runs resize -> crop -> flip -> ToTensor -> Normalize as two kernels on the raw uint8 frames
(`isf_image_aug_normalize`, csrc/image.cu — Pillow's 8-bit bicubic resampler in integer arithmetic, bit-identical values).
`data['img']` becomes a list of device tensors [3, fH, fW] (float32 like the reference, or fp16 for the backbone).
"""
import ctypes

import numpy as np
import torch

from . import _lib as L
from .registry import PIPELINES

_PLANS = {}


def _plan(geom, device):
    """device coefficient tables of one (input size, resize dims, crop, flip) — computed once per geometry."""
    key = (geom, str(device))
    hit = _PLANS.get(key)
    if hit is None:
        in_w, in_h, rs_w, rs_h, cx, cy, ow, oh, flip = geom
        nbytes = ctypes.c_size_t(0)
        L.check(L.lib().isf_image_plan_size(in_w, in_h, rs_w, rs_h, ow, oh, ctypes.byref(nbytes)))
        tab = torch.empty(nbytes.value // 4, dtype=torch.int32, device=device)
        rows = (ctypes.c_int * 2)()
        L.check(L.lib().isf_image_plan(in_w, in_h, rs_w, rs_h, cx, cy, ow, oh, flip, L.ptr(tab), ctypes.c_size_t(nbytes.value), rows,
                                       L.stream()))
        hit = (tab, int(rows[0]), int(rows[1]))
        if len(_PLANS) > 256:           # training draws a new geometry per image: keep the cache bounded
            _PLANS.clear()
        _PLANS[key] = hit
    return hit


def image_aug_normalize(imgs, resize_dims, crop, flip, mean, std, out_dtype=torch.float32):
    """imgs [n, H, W, 3] uint8 DEVICE tensor (n cameras sharing one geometry) -> [n, 3, fH, fW] `out_dtype`:
    PIL resize(resize_dims) (bicubic) -> crop(crop) -> optional left-right flip -> /255 -> (x - mean) / std."""
    L.require_cuda(imgs)
    if imgs.dtype != torch.uint8 or imgs.dim() != 4 or imgs.size(3) != 3:
        raise L.IsfError('image_aug_normalize: uint8 [n, H, W, 3] expected')
    if out_dtype not in (torch.float32, torch.float16):
        raise L.IsfError('image_aug_normalize: float32 or float16 output')
    n, in_h, in_w = imgs.size(0), imgs.size(1), imgs.size(2)
    ow, oh = crop[2] - crop[0], crop[3] - crop[1]
    geom = (in_w, in_h, int(resize_dims[0]), int(resize_dims[1]), int(crop[0]), int(crop[1]), ow, oh, 1 if flip else 0)
    tab, row_lo, row_hi = _plan(geom, imgs.device)
    wsb = ctypes.c_size_t(0)
    L.check(L.lib().isf_image_aug_workspace_size(n, ow, row_lo, row_hi, ctypes.byref(wsb)))
    ws = L.workspace(wsb.value, imgs.device)
    out = torch.empty((n, 3, oh, ow), dtype=out_dtype, device=imgs.device)
    L.check(L.lib().isf_image_aug_normalize(
        L.ptr(imgs), n, in_w, in_h, geom[2], geom[3], ow, oh, row_lo, row_hi, L.ptr(tab), L.host_f32(mean), L.host_f32(std),
        L.ptr(out), 1 if out_dtype == torch.float16 else 0, L.ptr(ws), ctypes.c_size_t(ws.numel()), L.stream()))
    return out


def _as_u8(img):
    """PIL image / numpy array / tensor -> contiguous uint8 [H, W, 3] numpy array or tensor (no copy when already so)."""
    if isinstance(img, torch.Tensor):
        return img
    a = np.asarray(img)
    if a.dtype != np.uint8:
        raise L.IsfError('camera frames must be uint8 (LoadMultiViewImageFromFilesV2 without to_float32)')
    return np.ascontiguousarray(a)


@PIPELINES.register_module()
class ImageAug3D:
    """ref: transforms_3d.py:46-145.  Samples (resize, crop, flip, rotate) per camera exactly like the reference (same
    numpy draws in the same order) and writes `img_aug_matrix`; the pixels are left untouched for `ImageNormalize`."""

    def __init__(self, final_dim, resize_lim, bot_pct_lim, rot_lim, rand_flip, is_train):
        self.final_dim, self.resize_lim, self.bot_pct_lim = final_dim, resize_lim, bot_pct_lim
        self.rand_flip, self.rot_lim, self.is_train = rand_flip, rot_lim, is_train

    def sample_augmentation(self, results):
        """(resize, resize_dims, crop, flip, rotate) of one camera.  Training draws, in the reference's order (transforms_3d.py:
        60-70): scale, bottom fraction, horizontal offset, flip coin, angle — so a seeded run reproduces the reference's
        augmentation; inference uses the interval means and a centred crop (:71-79)."""
        src_w, src_h = results['ori_shape']
        out_h, out_w = self.final_dim
        draw = (lambda lim: np.random.uniform(*lim)) if self.is_train else (lambda lim: np.mean(lim))
        scale = draw(self.resize_lim)
        dims = (int(src_w * scale), int(src_h * scale))
        top = int((1 - draw(self.bot_pct_lim)) * dims[1]) - out_h
        slack = max(0, dims[0] - out_w)
        left = int(np.random.uniform(0, slack)) if self.is_train else int(slack / 2)
        flip = bool(self.is_train and self.rand_flip and np.random.choice([0, 1]))
        angle = np.random.uniform(*self.rot_lim) if self.is_train else 0
        return scale, dims, (left, top, left + out_w, top + out_h), flip, angle

    @staticmethod
    def aug_matrix(resize, crop, flip):
        """the post-homography of img_transform for rotate = 0 (:92-110) as the 4x4 float32 matrix of :136-140."""
        rotation = torch.eye(2) * resize
        translation = torch.zeros(2) - torch.Tensor(crop[:2])
        if flip:
            A = torch.Tensor([[-1, 0], [0, 1]])
            rotation = A.matmul(rotation)
            translation = A.matmul(translation) + torch.Tensor([crop[2] - crop[0], 0])
        transform = torch.eye(4)
        transform[:2, :2] = rotation
        transform[:2, 3] = translation
        return transform.numpy()

    def __call__(self, data):
        if 'img' not in data:
            return data
        plans, transforms = [], []
        for _ in data['img']:
            resize, resize_dims, crop, flip, rotate = self.sample_augmentation(data)
            if rotate != 0:
                raise NotImplementedError('ImageAug3D on the device: in-plane rotation (rot_lim != 0) is not implemented')
            plans.append((resize_dims, crop, flip))
            transforms.append(self.aug_matrix(resize, crop, flip))
        data['img_aug_plan'] = plans
        data['img_aug_matrix'] = transforms
        return data


@PIPELINES.register_module()
class ImageNormalize:
    """ref: transforms_3d.py:25-43, fused with the pixel half of ImageAug3D: consumes the raw uint8 frames + the recorded
    geometry; cameras that share a geometry (always, at test time) go through one launch pair."""

    def __init__(self, mean, std, device='cuda', out_dtype=torch.float32):
        self.mean, self.std, self.device, self.out_dtype = mean, std, device, out_dtype

    def __call__(self, data):
        if 'img' not in data:
            return data
        frames = [_as_u8(i) for i in data['img']]
        plans = data.pop('img_aug_plan', None)
        if plans is None:               # no ImageAug3D before: identity geometry
            plans = [((f.shape[1], f.shape[0]), (0, 0, f.shape[1], f.shape[0]), False) for f in frames]
        out = [None] * len(frames)
        groups = {}
        for i, (f, p) in enumerate(zip(frames, plans)):
            groups.setdefault((tuple(f.shape), p), []).append(i)
        for (shape, (resize_dims, crop, flip)), idx in groups.items():
            if isinstance(frames[idx[0]], torch.Tensor):
                dev = torch.stack([frames[i] for i in idx]).to(self.device, non_blocking=True)
            else:
                host = torch.empty((len(idx),) + shape, dtype=torch.uint8)
                if self.device != 'cpu' and torch.cuda.is_available():
                    host = host.pin_memory()
                np.stack([frames[i] for i in idx], 0, out=host.numpy())
                dev = host.to(self.device, non_blocking=True)
            res = image_aug_normalize(dev.contiguous(), resize_dims, crop, flip, self.mean, self.std, self.out_dtype)
            for j, i in enumerate(idx):
                out[i] = res[j]
        data['img'] = out
        data['img_norm_cfg'] = dict(mean=self.mean, std=self.std)
        return data
