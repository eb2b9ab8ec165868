"""ctypes access to oracle/_ref/libref_spconv.so — the reference's vendored spconv-1.x CPU rulebook generator
(bevfusion-ops/spconv/include/spconv/geometry.h, see oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'libref_spconv.so')
_lib = None


def available():
    return os.path.exists(_SO)


def _i32(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def indice_pairs(indices, batch, spatial_shape, ksize, stride, padding, dilation, subm):
    """Returns (indices_out [Vo,4], out_shape, pairs [K,2,n] (-1 padded), num [K]) exactly as getIndicePair
    (spconv_ops.h:33-137) hands them to indiceConv on the CPU path."""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
        _lib.ref_spconv_indice_pairs.restype = ctypes.c_int
    idx = _i32(indices)
    n = idx.shape[0]
    ks, st, pd, dl = _i32(ksize), _i32(stride), _i32(padding), _i32(dilation)
    if subm:
        out_shape = _i32(spatial_shape)
    else:   # ops.py:20-31 get_conv_output_size
        out_shape = _i32([(spatial_shape[i] + 2 * pd[i] - dl[i] * (ks[i] - 1) - 1) // st[i] + 1 for i in range(3)])
    kvol = int(ks.prod())
    pairs = np.full((kvol, 2, n), -1, dtype=np.int32)
    num = np.zeros((kvol,), dtype=np.int32)
    out = np.zeros((max(n * kvol, 1), 4), dtype=np.int32)
    vo = _lib.ref_spconv_indice_pairs(_p(idx), n, int(batch), _p(ks), _p(st), _p(pd), _p(dl), _p(out_shape), int(bool(subm)),
                                      _p(out), _p(pairs), _p(num))
    return (idx if subm else out[:vo].copy()), out_shape.tolist(), pairs, num
