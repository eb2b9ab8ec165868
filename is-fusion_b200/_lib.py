"""ctypes binding of libisfusion_b200.so (the C-ABI in include/isfusion_b200.h).

The product path has NO fallback: if the library is missing or a call fails the
caller gets an exception.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libisfusion_b200.so')
_lib = None


class IsfError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IsfError('%s not found: build it with `python is-fusion_b200/build.py` '
                           '(there is no CPU fallback)' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.isf_last_error.restype = ctypes.c_char_p
        _lib.isf_launch_count.restype = ctypes.c_ulonglong
        for item in os.environ.get('ISF_OPTS', '').split(','):        # e.g. ISF_OPTS=gemm_v4=1,gemm_ln_fast=0 (isf_set_option)
            if '=' in item:
                name, value = item.split('=')
                if _lib.isf_set_option(name.strip().encode(), int(value)) != 0:
                    raise IsfError('ISF_OPTS: %s' % _lib.isf_last_error().decode())
    return _lib


def check(rc):
    if rc != 0:
        raise IsfError('isfusion_b200 error %d: %s' % (rc, lib().isf_last_error().decode()))


def ptr(t):
    """device pointer of a tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_f32(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def host_i32(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def is_device(t):
    """True for a CUDA tensor (the only kind the library takes)."""
    return t.is_cuda


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not is_device(t):
            raise IsfError('expected a CUDA tensor (the product path has no CPU fallback)')
        if not t.is_contiguous():
            raise IsfError('expected a contiguous tensor')


_ws_cache = {}


def workspace(nbytes, device):
    """grow-only scratch buffer per (device, stream): kernels on concurrent streams never share scratch memory."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def launch_count(reset=False):
    return int(lib().isf_launch_count(1 if reset else 0))


# 16-bit element type of the Hierarchical-Scene-Fusion path (camera -> BEV, conv_fusion, the SST levels): fp16 (default; what meets
# north_star's 1e-3) or bf16 (BASELINE configs[2]; the `_bf16` entry points: same kernels, bf16 encodings).  Set it BEFORE the first
# forward: the modules cache their packed weights in this type.
DTYPE = torch.float16


def set_dtype(dtype):
    global DTYPE
    if dtype not in (torch.float16, torch.bfloat16):
        raise IsfError('set_dtype: torch.float16 or torch.bfloat16')
    DTYPE = dtype


def fn16(name, dtype):
    """entry point `name` (an `..._f16` / untyped symbol) for 16-bit tensors of `dtype`."""
    if dtype == torch.float16:
        return getattr(lib(), name)
    if dtype != torch.bfloat16:
        raise IsfError('%s: fp16 or bf16 tensors expected, got %s' % (name, dtype))
    return getattr(lib(), name[:-4] + '_bf16' if name.endswith('_f16') else name + '_bf16')


def set_option(name, value):
    """runtime options of the library (isf_set_option), e.g. set_option('gemm_cluster', 4)."""
    check(lib().isf_set_option(name.encode(), int(value)))
