"""Host wrapper of the tcgen05 GEMM (csrc/tc_gemm.cuh, isf_gemm_f16)."""
import ctypes
import os

import torch

from . import _lib as L

ACT = {None: 0, 'none': 0, 'relu': 1, 'gelu': 2}



def pack_weight(w, dtype=None):
    """[N, K] (nn.Linear layout) -> K-major fp16 (bf16 under _lib.set_dtype) [N, Kpad], Kpad = K rounded up to 64, zero padded."""
    dtype = L.DTYPE if dtype is None else dtype
    n, k = w.shape
    kpad = (k + 63) // 64 * 64
    out = torch.zeros((n, kpad), dtype=dtype, device=w.device)
    out[:, :k] = w.to(dtype)
    return out.contiguous()


def gemm(a, w_packed, k=None, scale=None, shift=None, residual=None, act=None, ln=None, out=None, out_dtype=None,
         out_rows=None, ldo=None):
    """out = epi(a[M,K] @ w[N,K]^T): y = acc*scale + shift (+ residual) -> LayerNorm(gamma,beta,eps) -> act."""
    L.require_cuda(a, w_packed, scale, shift, residual)
    if out is not None and (not L.is_device(out) or out.stride(-1) != 1):
        raise L.IsfError('gemm: out must be a CUDA tensor with unit column stride')
    if a.dtype not in (torch.float16, torch.bfloat16) or w_packed.dtype != a.dtype or (residual is not None and residual.dtype != a.dtype):
        raise L.IsfError('gemm: fp16 (or bf16) operands of one type expected')
    out_dtype = a.dtype if out_dtype is None else out_dtype
    m, ka = a.shape
    k = ka if k is None else k
    n, kpad = w_packed.shape
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    ldo = out.stride(0) if ldo is None else ldo
    f16 = out.dtype != torch.float32
    if f16 and out.dtype != a.dtype:
        raise L.IsfError('gemm: a 16-bit output has the type of the operands')
    g, b, eps = (ln if ln is not None else (None, None, 0.0))
    L.check(L.fn16('isf_gemm_f16', a.dtype)(
        L.ptr(a), L.ptr(w_packed), m, k, kpad, n, L.ptr(scale), L.ptr(shift), L.ptr(residual),
        residual.stride(0) if residual is not None else 0, ACT[act], L.ptr(g), L.ptr(b), ctypes.c_float(eps),
        L.ptr(out if f16 else None), L.ptr(None if f16 else out), ldo, L.ptr(out_rows), L.ptr(None), L.stream()))
    return out
