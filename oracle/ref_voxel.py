"""ctypes access to oracle/_ref/libref_voxel.so (the reference's own C++ CPU
voxelization, see oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np
import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'libref_voxel.so')
_lib = None


def available():
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (libtorch must be loaded first)
        _lib = ctypes.CDLL(_SO)
        _lib.ref_hard_voxelize_cpu.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dynamic_voxelize(points, voxel_size, coors_range):
    lib = _load()
    p = np.ascontiguousarray(points.detach().cpu().numpy(), dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    out = np.zeros((p.shape[0], 3), dtype=np.int32)
    lib.ref_dynamic_voxelize_cpu(_fp(p), p.shape[0], p.shape[1], _fp(vs), _fp(cr), _fp(out))
    return torch.from_numpy(out)


def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels):
    lib = _load()
    p = np.ascontiguousarray(points.detach().cpu().numpy(), dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    voxels = np.zeros((max_voxels, max_points, p.shape[1]), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    npv = np.zeros((max_voxels,), dtype=np.int32)
    n = lib.ref_hard_voxelize_cpu(_fp(p), p.shape[0], p.shape[1], _fp(vs), _fp(cr), int(max_points),
                                  int(max_voxels), _fp(voxels), _fp(coors), _fp(npv))
    return torch.from_numpy(voxels[:n].copy()), torch.from_numpy(coors[:n].copy()), torch.from_numpy(npv[:n].copy())
