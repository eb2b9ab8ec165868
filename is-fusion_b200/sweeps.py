"""Input contract of the path: multi-sweep point loading with the merge on the device.

Mirrors the reference's `LoadPointsFromMultiSweeps` (mmdet3d/datasets/pipelines/loading.py:735-911: same constructor
arguments, same `results` keys, same sweep selection), but where the reference rotates / translates / filters every sweep
with numpy on the host and concatenates, this module gathers the raw key frame + sweeps into ONE pinned host block, copies
it once and lets `isf_merge_sweeps` (csrc/voxel.cu) do the per-sweep work on the GPU: `results['points']` is the device
tensor [N, len(use_dim)] the detector consumes (the reference wraps the same values in a LiDARPoints object).
"""
import ctypes

import numpy as np
import torch

from . import _lib as L
from .registry import PIPELINES


def merge_sweeps(raw, seg_start, seg_flags, lag, rot=None, trans=None, radius=1.0, load_dim=5, point_cloud_range=None):
    """raw [M, load_dim] fp32 DEVICE tensor (key frame + sweeps, in order); seg_start [S+1] row offsets; seg_flags [S]
    (bit 0: transform, bit 1: remove close); lag [S] seconds; rot [S,3,3] / trans [S,3] float64 (host); point_cloud_range =
    PointsRangeFilter fused into the same pass (strictly-inside test on the merged coordinates).  Returns
    (points [M, 5] fp32 with capacity M rows, count [1] int32 device = rows valid)."""
    L.require_cuda(raw)
    if raw.dtype != torch.float32 or raw.dim() != 2 or raw.size(1) != load_dim:
        raise L.IsfError('merge_sweeps: raw must be float32 [M, load_dim]')
    m, s = raw.size(0), len(seg_flags)
    out = torch.empty((m, 5), dtype=torch.float32, device=raw.device)
    count = torch.zeros(1, dtype=torch.int32, device=raw.device)
    wsb = ctypes.c_size_t(0)
    L.check(L.lib().isf_merge_sweeps_workspace_size(m, ctypes.byref(wsb)))
    ws = L.workspace(wsb.value, raw.device)
    rot = np.zeros((s, 3, 3)) if rot is None else np.ascontiguousarray(rot, dtype=np.float64).reshape(s, 3, 3)
    trans = np.zeros((s, 3)) if trans is None else np.ascontiguousarray(trans, dtype=np.float64).reshape(s, 3)
    dbl = lambda a: (ctypes.c_double * a.size)(*[float(v) for v in a.reshape(-1)])
    L.check(L.lib().isf_merge_sweeps(
        L.ptr(raw), m, load_dim, s, L.host_i32(seg_start), dbl(rot), dbl(trans), dbl(np.asarray(lag, dtype=np.float64)),
        L.host_i32(seg_flags), ctypes.c_float(radius), L.host_f32(point_cloud_range) if point_cloud_range is not None else None,
        L.ptr(out), L.ptr(count), L.ptr(ws), ctypes.c_size_t(ws.numel()), L.stream()))
    return out, count


@PIPELINES.register_module()
class LoadPointsFromMultiSweeps(object):
    """ref: loading.py:735-911.  `results['points']` = key-frame points (numpy [N0, load_dim] / tensor / an object with
    `.tensor`), `results['timestamp']`, `results['sweeps']` = list of dict(data_path | points, timestamp,
    sensor2lidar_rotation, sensor2lidar_translation).  A sweep dict may carry its raw `points` array instead of a file."""

    def __init__(self, sweeps_num=10, load_dim=5, use_dim=[0, 1, 2, 4], file_client_args=dict(backend='disk'),
                 pad_empty_sweeps=False, remove_close=False, test_mode=False, painting=False, device='cuda',
                 point_cloud_range=None):
        if painting:
            raise NotImplementedError('LoadPointsFromMultiSweeps: painting is outside the path')
        self.load_dim, self.sweeps_num, self.use_dim = load_dim, sweeps_num, list(use_dim)
        self.pad_empty_sweeps, self.remove_close, self.test_mode = pad_empty_sweeps, remove_close, test_mode
        self.device = device
        # extension: the pipeline's later PointsRangeFilter folded into the merge kernel (same points, one pass, one count)
        self.point_cloud_range = None if point_cloud_range is None else [float(v) for v in np.asarray(point_cloud_range, dtype=np.float32)]

    def _load_points(self, sweep):
        if 'points' in sweep:
            return np.asarray(sweep['points'], dtype=np.float32)
        path = sweep['data_path']
        return np.load(path) if path.endswith('.npy') else np.fromfile(path, dtype=np.float32)      # loading.py:778-798

    def plan(self, results):
        """host side: which sweeps, as segments of one raw block -> (blocks, seg_start, seg_flags, lag, rot, trans)."""
        key = results['points']
        key = key.tensor if hasattr(key, 'tensor') else key
        key = key.cpu().numpy() if isinstance(key, torch.Tensor) else np.asarray(key)
        key = np.ascontiguousarray(key, dtype=np.float32).reshape(-1, self.load_dim)
        ts = results['timestamp']
        blocks, flags, lag, rot, trans = [key], [0], [0.0], [np.eye(3)], [np.zeros(3)]
        sweeps = results['sweeps']
        if self.pad_empty_sweeps and len(sweeps) == 0:                              # loading.py:865-870
            for _ in range(self.sweeps_num):
                blocks.append(key); flags.append(2 if self.remove_close else 0); lag.append(0.0)
                rot.append(np.eye(3)); trans.append(np.zeros(3))
        else:
            if len(sweeps) <= self.sweeps_num:                                      # loading.py:872-878
                choices = np.arange(len(sweeps))
            elif self.test_mode:
                choices = np.arange(self.sweeps_num)
            else:
                choices = np.random.choice(len(sweeps), self.sweeps_num, replace=False)
            for idx in choices:
                sweep = sweeps[idx]
                blocks.append(self._load_points(sweep).reshape(-1, self.load_dim))
                flags.append(1 | (2 if self.remove_close else 0))
                lag.append(ts - sweep['timestamp'] / 1e6)
                rot.append(np.asarray(sweep['sensor2lidar_rotation'], dtype=np.float64))
                trans.append(np.asarray(sweep['sensor2lidar_translation'], dtype=np.float64))
        seg_start = np.concatenate([[0], np.cumsum([b.shape[0] for b in blocks])]).astype(np.int64)
        return blocks, seg_start, flags, lag, np.stack(rot), np.stack(trans)

    def __call__(self, results):
        blocks, seg_start, flags, lag, rot, trans = self.plan(results)
        m = int(seg_start[-1])
        host = torch.empty((m, self.load_dim), dtype=torch.float32)
        if self.device != 'cpu' and torch.cuda.is_available():
            host = host.pin_memory()
        np.concatenate(blocks, 0, out=host.numpy())
        raw = host.to(self.device, non_blocking=True)
        pts, count = merge_sweeps(raw, seg_start, flags, lag, rot, trans, radius=1.0, load_dim=self.load_dim,
                                  point_cloud_range=self.point_cloud_range)
        filtered = self.remove_close or self.point_cloud_range is not None
        n = int(count.item()) if filtered else m                     # only the filters make the size data dependent
        results['points'] = pts[:n][:, self.use_dim]
        return results

    def __repr__(self):
        return f'{self.__class__.__name__}(sweeps_num={self.sweeps_num})'


@PIPELINES.register_module()
class PointsRangeFilter(object):
    """ref: transforms_3d.py:2002-2043 with BasePoints.in_range_3d (core/points/base_points.py:208-230): keeps the points
    STRICTLY inside the range, order preserved.  Device tensor in, device tensor out (one masked select)."""

    def __init__(self, point_cloud_range):
        self.pcd_range = np.array(point_cloud_range, dtype=np.float32)

    def __call__(self, input_dict):
        pts = input_dict['points']
        r = torch.as_tensor(self.pcd_range, device=pts.device)
        mask = ((pts[:, 0] > r[0]) & (pts[:, 1] > r[1]) & (pts[:, 2] > r[2]) & (pts[:, 0] < r[3]) & (pts[:, 1] < r[4]) & (pts[:, 2] < r[5]))
        input_dict['points'] = pts[mask]
        for key in ('pts_instance_mask', 'pts_semantic_mask'):
            if input_dict.get(key) is not None:
                input_dict[key] = input_dict[key][mask.cpu().numpy()]
        return input_dict

    def __repr__(self):
        return f'{self.__class__.__name__}(point_cloud_range={self.pcd_range.tolist()})'


@PIPELINES.register_module()
class GlobalRotScaleTransV2:
    """ref: transforms_3d.py:1871-1915, inference branch: `lidar_aug_matrix` = identity (the train branch augments the
    ground-truth boxes as well, which is outside the path)."""

    def __init__(self, resize_lim, rot_lim, trans_lim, is_train):
        self.resize_lim, self.rot_lim, self.trans_lim, self.is_train = resize_lim, rot_lim, trans_lim, is_train

    def __call__(self, data):
        if self.is_train or isinstance(self.rot_lim, float):
            raise NotImplementedError('GlobalRotScaleTransV2: only the inference branch (identity) is on the path')
        data['lidar_aug_matrix'] = np.eye(4).astype(np.float32)
        return data


class Compose:
    """minimal mmdet Compose over the PIPELINES registry: a list of config dicts / callables applied in order."""

    def __init__(self, transforms):
        self.transforms = [PIPELINES.build(t) if isinstance(t, dict) else t for t in transforms]

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
            if data is None:
                return None
        return data


@PIPELINES.register_module()
class LoadPointsFromFile(object):
    """ref: loading.py `LoadPointsFromFile.__call__` (:1465-1520, the path's arguments: coord_type='LIDAR', load_dim,
    use_dim): raw float32 file -> `results['points']` [N, len(use_dim)] numpy (the reference wraps it in LiDARPoints)."""

    def __init__(self, coord_type='LIDAR', load_dim=6, use_dim=[0, 1, 2], shift_height=False, use_color=False,
                 file_client_args=dict(backend='disk'), painting=False, **kwargs):
        if shift_height or use_color or painting:
            raise NotImplementedError('LoadPointsFromFile: shift_height / use_color / painting are outside the path')
        assert coord_type in ('CAMERA', 'LIDAR', 'DEPTH')
        self.load_dim = load_dim
        self.use_dim = list(range(use_dim)) if isinstance(use_dim, int) else list(use_dim)
        assert max(self.use_dim) < load_dim, f'Expect all used dimensions < {load_dim}, got {use_dim}'

    def __call__(self, results):
        path = results['pts_filename']
        pts = np.load(path) if path.endswith('.npy') else np.fromfile(path, dtype=np.float32)
        results['points'] = pts.reshape(-1, self.load_dim)[:, self.use_dim]
        return results


@PIPELINES.register_module()
class LoadMultiViewImageFromFilesV2:
    """ref: loading.py:14-72: opens the camera files with PIL; the frames stay undecoded-lazy PIL images until ImageNormalize
    takes their uint8 pixels to the device."""

    def __init__(self, to_float32=False, color_type='unchanged'):
        self.to_float32, self.color_type = to_float32, color_type

    def __call__(self, results):
        if 'img_filename' not in results:
            return results
        from PIL import Image
        images = [Image.open(name) for name in results['img_filename']]
        results['filename'] = results['img_filename']
        results['img'] = images
        results['img_shape'] = results['ori_shape'] = results['pad_shape'] = images[0].size
        results['scale_factor'] = 1.0
        return results


@PIPELINES.register_module()
class Collect3DV2:
    """ref: formating.py:183-254.  Same key sets; the values the reference wraps in mmcv DataContainers are plain tensors
    here (`meta_keys`: float tensors, lists stacked) and `img_metas` is the plain dict of the `meta_lis_keys` present."""

    def __init__(self, keys, meta_keys=('camera_intrinsics', 'camera2ego', 'img_aug_matrix', 'lidar_aug_matrix'),
                 meta_lis_keys=('filename', 'timestamp', 'ori_shape', 'img_shape', 'lidar2image', 'depth2img', 'cam2img', 'pad_shape',
                                'scale_factor', 'flip', 'pcd_horizontal_flip', 'pcd_vertical_flip', 'box_mode_3d', 'box_type_3d',
                                'img_norm_cfg', 'pcd_trans', 'token', 'pcd_scale_factor', 'pcd_rotation', 'lidar_path',
                                'transformation_3d_flow', 'img_mask_idx')):
        self.keys, self.meta_keys, self.meta_lis_keys = keys, meta_keys, meta_lis_keys

    def __call__(self, results):
        data = {}
        for key in self.keys:
            if key not in self.meta_keys:
                val = results[key]
                data[key] = torch.stack(list(val)) if key == 'img' and isinstance(val, (list, tuple)) else val   # DefaultFormatBundle3D's stack
        for key in self.meta_keys:
            if key in results:
                data[key] = torch.as_tensor(np.array(results[key]))
        data['img_metas'] = {k: results[k] for k in self.meta_lis_keys if k in results}
        return data


def collate(samples):
    """the batch `forward(return_loss=False, **batch)` takes (mmcv collate + scatter + MultiScaleFlipAug3D's single-entry outer
    lists): points -> [[per-sample tensors]], img -> [[B, N, 3, H, W]], img_metas -> [[dicts]], matrices -> [B, ...]."""
    batch = {'points': [[s['points'] for s in samples]], 'img_metas': [[s['img_metas'] for s in samples]]}
    if 'img' in samples[0]:
        batch['img'] = [torch.stack([s['img'] for s in samples])]
    for key in samples[0]:
        if key not in ('points', 'img', 'img_metas'):
            batch[key] = torch.stack([torch.as_tensor(s[key]) for s in samples])
    return batch
