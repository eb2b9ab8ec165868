"""Sparse 3-D convolution modules on the C-ABI CUDA library.

Mirrors the spconv Python API the reference builds on (SparseConvTensor / SubMConv3d / SparseConv3d /
SparseSequential / SparseModule, see mmdet3d/ops/bevfusion-ops/spconv/{structure.py:20-63, conv.py:55-223,
modules.py} and the call sites mmdet3d/models/middle_encoders/sparse_encoder.py:11-15,122-136,
mmdet3d/ops/sparse_block.py:8-20) with weights in the spconv1/mmcv layout (kD,kH,kW,Cin,Cout); spconv2
checkpoints (Cout,kD,kH,kW,Cin) are accepted on load (write_spconv2.py:62-94).
"""
import ctypes
import itertools
import os

import torch
from torch import nn

from . import _lib as L
from .registry import CONV_LAYERS


def _triple(v):
    return [int(x) for x in v] if isinstance(v, (list, tuple)) else [int(v)] * 3


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, num_dev=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = grid
        # capacity mode (sync-free / CUDA-graph path): rows [0, *num_dev) of features / indices are valid, the
        # rest of the buffers is capacity.  None = exact mode (every row valid), the spconv API contract.
        self.num_dev = num_dev
        # rows are in ascending linear (b,z,y,x) key order (true for everything this library produces): enables the
        # bitmap-rank rulebook; False = unknown order -> hash-table rulebook
        self.sorted = False

    @property
    def spatial_size(self):
        n = 1
        for s in self.spatial_shape:
            n *= s
        return n

    def replace_feature(self, feature):
        t = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid, self.num_dev)
        t.indice_dict = self.indice_dict
        t.sorted = self.sorted
        return t

    def find_indice_pair(self, key):
        return self.indice_dict.get(key) if key is not None else None

    def exact(self):
        """capacity mode -> exact mode (one host read of the row count)."""
        if self.num_dev is None:
            return self
        n = int(self.num_dev.item())
        t = SparseConvTensor(self.features[:n], self.indices[:n], self.spatial_shape, self.batch_size, self.grid)
        return t

    def dense(self, channels_first=True):
        if self.num_dev is not None:
            return self.exact().dense(channels_first)
        f = self.features
        if f.requires_grad and torch.is_grad_enabled():      # training: device-side index glue that autograd understands
            d, h, w = self.spatial_shape
            idx = self.indices.long()
            out = torch.zeros((self.batch_size, d, h, w, f.shape[1]), dtype=f.dtype, device=f.device)
            out = out.index_put((idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]), f)
            return out.permute(0, 4, 1, 2, 3) if channels_first else out
        if f.dtype != torch.float16:
            f = f.half()
        L.require_cuda(f, self.indices)
        d, h, w = self.spatial_shape
        c = f.shape[1]
        out = torch.empty((self.batch_size, c, d, h, w), dtype=torch.float16, device=f.device)
        L.check(L.lib().isf_sparse_to_dense(L.ptr(f.contiguous()), L.ptr(self.indices), f.shape[0], c, self.batch_size,
                                            d, h, w, L.ptr(out), L.stream()))
        out = out.to(self.features.dtype)
        return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()


class Rulebook:
    """output indices + neighbour table of one (coordinate set, kernel geometry)."""
    __slots__ = ('out_indices', 'out_shape', 'nbr', 'nbr_stride', 'taps', 'num_out', 'num_dev', 'pairs', 'nbr_ordered',
                 'row_perm', 'tile_ctr', 'inverse', 'pair_lists')

    def __init__(self, out_indices, out_shape, nbr, nbr_stride, taps, num_out, num_dev=None):
        self.out_indices, self.out_shape, self.nbr = out_indices, out_shape, nbr
        self.nbr_stride, self.taps, self.num_out = nbr_stride, taps, num_out      # num_out = rows (capacity)
        self.num_dev = num_dev                                                    # device-side valid-row count
        self.pairs = None
        self.nbr_ordered = self.row_perm = None     # tile order (TILE_ORDER): regrouped table + tile slot -> storage row
        self.tile_ctr = None                        # int32[2], zero between launches: dynamic tile scheduling (DYNAMIC_TILES)
        self.inverse = None                         # training: rulebook of the data gradient of a strided conv (inverse_rulebook)
        self.pair_lists = None                      # training: (input rows, output rows, per-offset counts) of the weight gradient


# Opt-in tile order of the implicit GEMM (csrc/spconv.cu "Tile order", tests/analysis/tile_occupancy.py): None = rows in
# None = storage order; an int = rows regrouped by neighbour pattern inside chunks of 2**TILE_ORDER consecutive rows
# (31 = one global sort, the default: measured -12.5 % on the 21 implicit-GEMM launches of the bench frame, round 2).
# Results are bit-identical either way.  Environment: ISF_TILE_ORDER=<int> | none.
_to = os.environ.get('ISF_TILE_ORDER', '31')
TILE_ORDER = None if _to in ('none', 'off', '') else int(_to)
if os.environ.get('ISF_NBR_ROWS'):      # opt-in row-per-thread neighbour-table kernel (csrc/spconv.cu nbr_bitmap_rows_kernel)
    L.set_option('nbr_rows', int(os.environ['ISF_NBR_ROWS']))


# Dynamic tile scheduling of the implicit GEMM (isf_spconv_tile_counter): the kernel's CTAs draw their row tiles from a counter
# owned by the rulebook instead of walking them round-robin (32- / 64-column layers, tc::gemm4_kernel); bit-identical results.
# Opt-in (ISF_DYNAMIC_TILES=1): measured -4 % on the stage-2 launches, -0.8 % on the 21 launches (call 36) against a CPU model that
# promised -22 %: a tile's cost is dominated by its fixed part, not by its 1 ... 27 kernel offsets.
DYNAMIC_TILES = os.environ.get('ISF_DYNAMIC_TILES', '0') == '1'
TILE_ORDER_FUSED = os.environ.get('ISF_TILE_ORDER_FUSED', '1') != '0'   # sorted tensors: build the table in tile order directly


def _tile_order(rb, ksize):
    """attach the regrouped neighbour table to a freshly built rulebook (same stream as the rulebook build)."""
    if TILE_ORDER is None or rb.taps < 2:
        return rb
    lib = L.lib()
    dev = rb.nbr.device
    wsb = ctypes.c_size_t(0)
    L.check(lib.isf_spconv_tile_order_workspace_size(rb.num_out, ctypes.byref(wsb)))
    ws = L.workspace(wsb.value, dev)
    rb.nbr_ordered = torch.empty_like(rb.nbr)
    rb.row_perm = torch.empty(rb.nbr_stride, dtype=torch.int32, device=dev)
    L.check(lib.isf_spconv_tile_order(L.ptr(rb.nbr), ctypes.c_longlong(rb.nbr_stride), rb.taps, L.host_i32(ksize), rb.num_out,
                                      L.ptr(rb.num_dev), int(TILE_ORDER), L.ptr(rb.nbr_ordered), L.ptr(rb.row_perm), L.ptr(ws),
                                      ctypes.c_size_t(ws.numel()), L.stream()))
    return rb


def _hash_table(x):
    """coordinate hash of a tensor's indices (cached on the tensor: all convs reading it share it)."""
    key = ('hash', x.indices.data_ptr(), x.indices.shape[0])
    tab = x.indice_dict.get(key)
    if tab is None:
        v = x.indices.shape[0]
        cap = L.lib().isf_hash_capacity(v)
        table = torch.empty(2 * cap, dtype=torch.int32, device=x.indices.device)
        L.check(L.lib().isf_hash_build(L.ptr(x.indices), v, L.ptr(x.num_dev), L.host_i32(x.spatial_shape), x.batch_size,
                                       L.ptr(table), cap, L.stream()))
        tab = (table, cap)
        x.indice_dict[key] = tab
    return tab


OVERFLOW_WATCH = []     # (count tensor, capacity) of every capacity-mode strided conv since the last check


def _watch(cnt, cap):
    """bounded: eager nosync loops that never call check_overflow() must not pin device tensors without limit."""
    OVERFLOW_WATCH.append((cnt, cap))
    if len(OVERFLOW_WATCH) > 512:
        del OVERFLOW_WATCH[:-128]


def check_overflow():
    """capacity mode: did any strided conv produce more output sites than its buffer holds? (one host sync)"""
    bad = [(int(c.item()), cap) for c, cap in OVERFLOW_WATCH if int(c.item()) >= cap]
    OVERFLOW_WATCH.clear()
    return bad


def _bitmap(x):
    """bitmap + prefix coordinate->row map of a tensor in sorted row order (cached on the tensor)."""
    key = ('bitmap', x.indices.data_ptr(), x.indices.shape[0])
    bm = x.indice_dict.get(key)
    if bm is None:
        lib = L.lib()
        dev = x.indices.device
        words = int(lib.isf_bitmap_words(L.host_i32(x.spatial_shape), x.batch_size))
        bits = torch.empty(words, dtype=torch.int32, device=dev)
        prefix = torch.empty(words, dtype=torch.int32, device=dev)
        wsb = ctypes.c_size_t(0)
        L.check(lib.isf_bitmap_workspace_size(ctypes.c_longlong(words), ctypes.byref(wsb)))
        ws = L.workspace(wsb.value, dev)
        L.check(lib.isf_bitmap_build(L.ptr(x.indices), x.indices.shape[0], L.ptr(x.num_dev), L.host_i32(x.spatial_shape),
                                     x.batch_size, L.ptr(bits), L.ptr(prefix), L.ptr(ws), ctypes.c_size_t(ws.numel()),
                                     L.stream()))
        bm = (bits, prefix)
        x.indice_dict[key] = bm
    return bm


def build_rulebook_bitmap(x, ksize, stride, padding, dilation, subm, out_capacity=None):
    """rulebook of a tensor in sorted row order: no hash table, no sort (see csrc/spconv.cu, bitmap-rank rulebook)."""
    lib = L.lib()
    dev = x.indices.device
    v_in = x.indices.shape[0]
    bits, prefix = _bitmap(x)
    taps = ksize[0] * ksize[1] * ksize[2]
    num_dev = None
    if subm:
        out_indices, out_shape, v_out, num_dev = x.indices, list(x.spatial_shape), v_in, x.num_dev
    else:
        oshape = (ctypes.c_int * 3)()
        L.check(lib.isf_spconv_out_shape(L.host_i32(x.spatial_shape), L.host_i32(ksize), L.host_i32(stride),
                                         L.host_i32(padding), L.host_i32(dilation), oshape))
        out_shape = list(oshape)
        cells = x.batch_size * out_shape[0] * out_shape[1] * out_shape[2]
        reach = 1
        for d in range(3):
            reach *= -(-ksize[d] // stride[d])
        cap_out = min(v_in * reach, cells)
        capacity_mode = x.num_dev is not None
        if capacity_mode and out_capacity is not None:
            cap_out = min(cap_out, int(out_capacity))
        words = int(lib.isf_bitmap_words(L.host_i32(out_shape), x.batch_size))
        obits = torch.empty(words, dtype=torch.int32, device=dev)
        oprefix = torch.empty(words, dtype=torch.int32, device=dev)
        wsb = ctypes.c_size_t(0)
        L.check(lib.isf_bitmap_workspace_size(ctypes.c_longlong(words), ctypes.byref(wsb)))
        ws = L.workspace(wsb.value, dev)
        out_buf = torch.empty((max(cap_out, 1), 4), dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        L.check(lib.isf_bitmap_out_sites(L.ptr(x.indices), v_in, L.ptr(x.num_dev), L.host_i32(x.spatial_shape), x.batch_size,
                                         L.host_i32(ksize), L.host_i32(stride), L.host_i32(padding), L.host_i32(dilation),
                                         L.ptr(obits), L.ptr(oprefix), L.ptr(out_buf), cap_out, L.ptr(cnt), L.ptr(ws),
                                         ctypes.c_size_t(ws.numel()), L.stream()))
        if capacity_mode:
            v_out, out_indices, num_dev = cap_out, out_buf, cnt
            _watch(cnt, cap_out)
        else:
            v_out = int(cnt.item())
            out_indices = out_buf[:v_out]
        # the output-site bitmap IS the coordinate map of the output tensor: the next layer reuses it
        x.indice_dict[('bitmap', out_indices.data_ptr(), out_indices.shape[0])] = (obits, oprefix)
    stride_rows = max((v_out + 127) // 128 * 128, 128)
    nbr = torch.empty((taps, stride_rows), dtype=torch.int32, device=dev)
    if TILE_ORDER is not None and TILE_ORDER_FUSED and taps >= 2 and out_indices.data_ptr() % 16 == 0:
        wsb = ctypes.c_size_t(0)
        L.check(lib.isf_spconv_tile_order_workspace_size(v_out, ctypes.byref(wsb)))
        ws = L.workspace(wsb.value, dev)
        perm = torch.empty(stride_rows, dtype=torch.int32, device=dev)
        L.check(lib.isf_spconv_neighbors_bitmap_ordered(
            L.ptr(out_indices), v_out, L.ptr(num_dev), L.ptr(bits), L.ptr(prefix), L.host_i32(x.spatial_shape), L.host_i32(ksize),
            L.host_i32(stride), L.host_i32(padding), L.host_i32(dilation), 1 if subm else 0, int(TILE_ORDER), L.ptr(nbr),
            ctypes.c_longlong(stride_rows), L.ptr(perm), L.ptr(ws), ctypes.c_size_t(ws.numel()), L.stream()))
        rb = Rulebook(out_indices, out_shape, nbr, stride_rows, taps, v_out, num_dev)
        rb.nbr_ordered, rb.row_perm = nbr, perm          # rb.nbr (same tensor) only serves the pair count
        return rb
    L.check(lib.isf_spconv_neighbors_bitmap(L.ptr(out_indices), v_out, L.ptr(num_dev), L.ptr(bits), L.ptr(prefix),
                                            L.host_i32(x.spatial_shape), L.host_i32(ksize), L.host_i32(stride),
                                            L.host_i32(padding), L.host_i32(dilation), 1 if subm else 0, L.ptr(nbr),
                                            ctypes.c_longlong(stride_rows), L.stream()))
    return _tile_order(Rulebook(out_indices, out_shape, nbr, stride_rows, taps, v_out, num_dev), ksize)


def build_rulebook(x, ksize, stride, padding, dilation, subm, out_capacity=None):
    if x.sorted:
        return build_rulebook_bitmap(x, ksize, stride, padding, dilation, subm, out_capacity)
    lib = L.lib()
    dev = x.indices.device
    v_in = x.indices.shape[0]
    table, cap = _hash_table(x)
    taps = ksize[0] * ksize[1] * ksize[2]
    num_dev = None
    if subm:
        out_indices, out_shape, v_out, num_dev = x.indices, list(x.spatial_shape), v_in, x.num_dev
    else:
        oshape = (ctypes.c_int * 3)()
        L.check(lib.isf_spconv_out_shape(L.host_i32(x.spatial_shape), L.host_i32(ksize), L.host_i32(stride),
                                         L.host_i32(padding), L.host_i32(dilation), oshape))
        out_shape = list(oshape)
        cells = x.batch_size * out_shape[0] * out_shape[1] * out_shape[2]
        # an input site reaches at most ceil(k/s) outputs per axis
        reach = 1
        for d in range(3):
            reach *= -(-ksize[d] // stride[d])
        cap_out = min(v_in * reach, cells)
        capacity_mode = x.num_dev is not None
        if capacity_mode and out_capacity is not None:
            cap_out = min(cap_out, int(out_capacity))
        ws_bytes = ctypes.c_size_t(0)
        L.check(lib.isf_spconv_out_indices_workspace_size(v_in, taps, cap_out, ctypes.byref(ws_bytes)))
        ws = L.workspace(ws_bytes.value, dev)
        out_buf = torch.empty((max(cap_out, 1), 4), dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        L.check(lib.isf_spconv_out_indices(L.ptr(x.indices), v_in, L.ptr(x.num_dev), L.host_i32(x.spatial_shape),
                                           x.batch_size, L.host_i32(ksize), L.host_i32(stride), L.host_i32(padding),
                                           L.host_i32(dilation), L.ptr(out_buf), cap_out, L.ptr(cnt), L.ptr(ws),
                                           ctypes.c_size_t(ws.numel()), L.stream()))
        if capacity_mode:
            v_out, out_indices, num_dev = cap_out, out_buf, cnt
            _watch(cnt, cap_out)
        else:
            v_out = int(cnt.item())      # exact mode: one host read per strided layer
            out_indices = out_buf[:v_out]
    stride_rows = max((v_out + 127) // 128 * 128, 128)
    nbr = torch.empty((taps, stride_rows), dtype=torch.int32, device=dev)
    L.check(lib.isf_spconv_neighbors(L.ptr(out_indices), v_out, L.ptr(num_dev), L.ptr(table), cap,
                                     L.host_i32(x.spatial_shape), L.host_i32(ksize), L.host_i32(stride),
                                     L.host_i32(padding), L.host_i32(dilation), 1 if subm else 0, L.ptr(nbr),
                                     ctypes.c_longlong(stride_rows), L.stream()))
    return _tile_order(Rulebook(out_indices, out_shape, nbr, stride_rows, taps, v_out, num_dev), ksize)


PROFILE = None      # bench.py: list collecting the arguments of every implicit-GEMM launch of a frame


def useful_flops(rb_cin_cout):
    """2 * pairs * Cin * Cout of one launch (pairs counted once per rulebook, outside any timed region)."""
    rb, cin, cout = rb_cin_cout
    if rb.pairs is None:
        n = rb.num_out if rb.num_dev is None else int(rb.num_dev.item())
        rb.pairs = int((rb.nbr[:, :n] >= 0).sum().item())
    return 2.0 * rb.pairs * cin * cout


def implicit_gemm(feats, rb, w_packed, cin, cout, scale=None, shift=None, residual=None, relu=False, out=None):
    if PROFILE is not None:
        PROFILE.append((feats, rb, w_packed, cin, cout, scale, shift, residual, relu))
    return _implicit_gemm(feats, rb, w_packed, cin, cout, scale, shift, residual, relu, out)


def issued_flops(rb, cin, cout, kpad=None):
    """FLOPs the kernel actually issues for one launch: it walks K in 64-wide chunks and skips a chunk for a whole
    128-row tile when none of the taps the chunk covers has a neighbour in that tile; every other chunk is a
    128 x 64 x Cout MMA block, zero-filled slots included.  Uses the tile-ordered table when the rulebook has one."""
    n = rb.num_out if rb.num_dev is None else int(rb.num_dev.item())
    nbr = rb.nbr_ordered if rb.nbr_ordered is not None else rb.nbr
    taps = rb.taps
    kpad = kpad if kpad is not None else (taps * cin + 63) // 64 * 64
    valid = nbr[:, :n] >= 0
    pad = (-n) % 128
    if pad:
        valid = torch.cat([valid, torch.zeros((taps, pad), dtype=torch.bool, device=valid.device)], 1)
    tap_active = valid.view(taps, -1, 128).any(2)                       # [taps, tiles]
    chunks = kpad // 64
    q = torch.arange(chunks, device=valid.device)
    t0 = (q * 64) // cin
    t1 = torch.clamp((q * 64 + 63) // cin, max=taps - 1)
    csum = torch.cat([torch.zeros((1, tap_active.shape[1]), dtype=torch.int32, device=valid.device),
                      torch.cumsum(tap_active.int(), 0)], 0)            # prefix over taps
    inside = t0 < taps
    chunk_active = ((csum[t1 + 1] - csum[torch.clamp(t0, max=taps - 1)]) > 0) & inside[:, None]
    tiles_without_any = (~chunk_active.any(0)).sum()                    # the kernel still issues one chunk there
    blocks = int(chunk_active.sum().item()) + int(tiles_without_any.item())
    return 2.0 * blocks * 128 * 64 * cout


def time_launches(records, reps=5):
    """re-issue every recorded implicit-GEMM launch `reps` times back to back between CUDA events on the current
    stream (queue kept full, so the event time is kernel time) -> [(ms per launch, issued flops, useful flops)]."""
    out = []
    for (feats, rb, w, cin, cout, scale, shift, residual, relu) in records:
        buf = torch.empty((rb.num_out, cout), dtype=torch.float16, device=feats.device)
        _implicit_gemm(feats, rb, w, cin, cout, scale, shift, residual, relu, buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _implicit_gemm(feats, rb, w, cin, cout, scale, shift, residual, relu, buf)
        e1.record()
        torch.cuda.synchronize()
        n = rb.num_out if rb.num_dev is None else int(rb.num_dev.item())
        out.append((e0.elapsed_time(e1) / reps, issued_flops(rb, cin, cout, w.shape[1]), useful_flops((rb, cin, cout)), cin, cout, n))
    return out


def _implicit_gemm(feats, rb, w_packed, cin, cout, scale=None, shift=None, residual=None, relu=False, out=None):
    if feats.dtype != torch.float16:
        raise L.IsfError('sparse conv: fp16 features expected')
    L.require_cuda(feats, w_packed, scale, shift, residual)
    if feats.shape[0] * cin * 2 >= 2 ** 32:
        raise L.IsfError('sparse conv: feature matrix larger than 4 GiB (the kernel stages 32-bit row byte offsets)')
    if out is None:
        out = torch.empty((rb.num_out, cout), dtype=torch.float16, device=feats.device)
    ldr = residual.stride(0) if residual is not None else 0
    if DYNAMIC_TILES:
        if rb.tile_ctr is None:
            rb.tile_ctr = torch.zeros(2, dtype=torch.int32, device=feats.device)
        L.check(L.lib().isf_spconv_tile_counter(L.ptr(rb.tile_ctr)))       # consumed by the GEMM call below
    if rb.row_perm is not None:
        L.check(L.lib().isf_spconv_implicit_gemm_ordered(
            L.ptr(feats), cin, L.ptr(w_packed), w_packed.shape[1], cout, L.ptr(rb.nbr_ordered),
            ctypes.c_longlong(rb.nbr_stride), rb.taps, rb.num_out, L.ptr(rb.num_dev), L.ptr(rb.row_perm), L.ptr(scale),
            L.ptr(shift), L.ptr(residual), ldr, 1 if relu else 0, L.ptr(out), out.stride(0), L.stream()))
        return out
    L.check(L.lib().isf_spconv_implicit_gemm(
        L.ptr(feats), cin, L.ptr(w_packed), w_packed.shape[1], cout, L.ptr(rb.nbr), ctypes.c_longlong(rb.nbr_stride),
        rb.taps, rb.num_out, L.ptr(rb.num_dev), L.ptr(scale), L.ptr(shift), L.ptr(residual), ldr, 1 if relu else 0,
        L.ptr(out), out.stride(0), L.stream()))
    return out


# ------------------------------------------------------------------------------------------------------------------
# Training: backward of the implicit GEMM.
#   data gradient   din[i] = sum_k dout[o(k, i)] . W[k]^T  is the SAME kernel over the transposed rulebook: for a submanifold conv
#                   the transposed table is the table itself with the kernel offsets mirrored (o at offset k of i  <=>  i at offset
#                   K-1-k of o), i.e. the forward launch with weights W'[k] = W[K-1-k]^T; for a strided conv the inverse table
#                   inv[k][i] = o is scattered from the forward table once per rulebook (device-side index glue);
#   weight gradient dW[k] = X[nbr[k]]^T . dout  (a gathered A^T B per kernel offset): isf_spconv_wgrad (csrc/wgrad.cu) over the
#                   rulebook's pair lists; ISF_WGRAD=torch = index_select copies + one library GEMM per offset (the cross-check).
# Mixed precision like the inference path: fp16 operands, fp32 accumulation, fp32 master weights and gradients.
# ------------------------------------------------------------------------------------------------------------------
def _slot_rows(rb):
    """storage row of every tile slot of the rulebook's table (identity without tile order)."""
    if rb.row_perm is not None:
        return rb.row_perm[:rb.nbr_stride].long()
    return None


def inverse_rulebook(rb, v_in):
    """rulebook of the data gradient of a strided conv: inv[k][i] = output row that reads input row i at kernel offset k."""
    if rb.inverse is None:
        tab = rb.nbr_ordered if rb.nbr_ordered is not None else rb.nbr
        taps = rb.taps
        n = rb.num_out
        tab = tab[:, :n] if rb.row_perm is None else tab
        rows = _slot_rows(rb)
        stride = (v_in + 127) // 128 * 128
        inv = torch.full((taps, stride), -1, dtype=torch.int32, device=tab.device)
        k_idx = torch.arange(taps, device=tab.device).view(taps, 1).expand_as(tab)
        o_idx = (rows.view(1, -1).expand_as(tab) if rows is not None else
                 torch.arange(tab.shape[1], device=tab.device).view(1, -1).expand_as(tab))
        ok = (tab >= 0) & (o_idx >= 0)
        inv[k_idx[ok], tab[ok].long()] = o_idx[ok].int()
        rb.inverse = Rulebook(None, None, inv.contiguous(), stride, taps, v_in)
    return rb.inverse


def pair_lists(rb):
    """(input rows, output rows) of every (kernel offset, output site) pair of the rulebook, grouped by kernel offset, and the
    per-offset pair counts as a host list: built once per rulebook (one host read), shared by the convolutions that share it."""
    if rb.pair_lists is None:
        tab = rb.nbr_ordered if rb.nbr_ordered is not None else rb.nbr[:, :rb.num_out]
        rows = _slot_rows(rb)
        ok = tab >= 0
        if rows is not None:
            ok = ok & (rows >= 0).view(1, -1)
        counts = ok.sum(1).tolist()                                 # the one host read
        kk, ss = ok.nonzero(as_tuple=True)                          # row-major: grouped by kernel offset, slots ascending
        in_rows = tab[kk, ss].long()
        out_rows = rows[ss] if rows is not None else ss
        offs = torch.tensor([0] + list(itertools.accumulate(counts)), dtype=torch.int32).to(tab.device)
        rb.pair_lists = (in_rows.contiguous(), out_rows.contiguous(), counts, in_rows.int().contiguous(),
                         out_rows.int().contiguous(), offs)
    return rb.pair_lists


# weight gradient: 'own' = the fused gather + mma.sync kernel of csrc/wgrad.cu, 'torch' = gathered copies + one library GEMM per
# kernel offset (the first training path; kept as the in-product cross-check of the own kernel, tests/test_sparse_train_gpu.py)
WGRAD = os.environ.get('ISF_WGRAD', 'own')


def _weight_grad(xh, gh, rb, taps, cin, cout):
    """sum over the pairs of each kernel offset of x[in]^T g[out]: [taps, Cin, Cout] fp32 (reference: indice_conv_backward's
    filter gradient, bevfusion-ops/spconv/include/spconv/spconv_ops.h:364-470)."""
    in_rows, out_rows, counts, in32, out32, offs = pair_lists(rb)
    gw = torch.zeros((taps, cin, cout), dtype=torch.float32, device=xh.device)
    if WGRAD == 'own' and cin % 8 == 0 and cout % 8 == 0:
        L.require_cuda(xh, gh)
        L.check(L.lib().isf_spconv_wgrad(L.ptr(xh), cin, L.ptr(gh), cout, L.ptr(in32), L.ptr(out32), L.ptr(offs), taps,
                                         max(counts) if counts else 0, L.ptr(gw), L.stream()))
        return gw
    xg = xh.index_select(0, in_rows)                       # [pairs, Cin]  fp16, grouped by kernel offset
    gg = gh.index_select(0, out_rows)                      # [pairs, Cout] fp16 (scaled)
    off = 0
    for k in range(taps):                                  # counts live on the host: no synchronisation in this loop
        n = counts[k]
        if n:
            gw[k] = _mm_f32(xg[off:off + n].t(), gg[off:off + n])      # fp16 operands on the tensor cores, fp32 result
            off += n
    return gw


def pack_conv_weight_transposed(weight, mirror):
    """weights of the data-gradient launch: [Cin, Kpad] with column k*Cout + c = W[k' ][cin][c], k' = K-1-k when the table is the
    forward table of a submanifold conv (mirror), k' = k over an inverse table."""
    kd, kh, kw, cin, cout = weight.shape
    taps = kd * kh * kw
    w = weight.detach().reshape(taps, cin, cout)
    if mirror:
        w = w.flip(0)
    w = w.permute(1, 0, 2).reshape(cin, taps * cout)            # [Cin, taps * Cout]
    kpad = (taps * cout + 63) // 64 * 64
    out = torch.zeros((cin, kpad), dtype=torch.float16, device=weight.device)
    out[:, :taps * cout] = w.half()
    return out.contiguous()


def _mm_f32(a, b):
    """a @ b for fp16 operands with an fp32 result (library GEMM of the weight gradient)."""
    if a.is_cuda:
        try:
            return torch.mm(a, b, out_dtype=torch.float32)
        except (TypeError, RuntimeError):
            pass
    return a.float() @ b.float()


class ImplicitGemmFunction(torch.autograd.Function):
    """y = conv(x; W) over a rulebook, differentiable in x and W (train mode of SparseConvolution)."""

    @staticmethod
    def forward(ctx, feats, weight, rb, subm):
        cin, cout = weight.shape[3], weight.shape[4]
        xh = feats.detach().half().contiguous()
        y = _implicit_gemm(xh, rb, pack_conv_weight(weight), cin, cout)
        ctx.save_for_backward(xh, weight)
        ctx.rb, ctx.subm, ctx.in_dtype = rb, subm, feats.dtype
        return y.to(feats.dtype if feats.dtype != torch.float16 else torch.float16)

    @staticmethod
    def backward(ctx, grad_out):
        xh, weight = ctx.saved_tensors
        rb, subm = ctx.rb, ctx.subm
        cin, cout = weight.shape[3], weight.shape[4]
        taps = rb.taps
        g32 = grad_out.detach().float()
        grad_x = grad_w = None
        # per-tensor scaling into fp16's range (gradients deep in the encoder are ~1e-6: they would flush to zero in fp16);
        # the scale is a device-side power of two, so the un-scaling is exact and nothing synchronises with the host
        amax = g32.abs().amax().clamp(min=1e-30)
        scale = torch.exp2(torch.floor(torch.log2(1024.0 / amax)))
        gh = (g32 * scale).half().contiguous()
        if ctx.needs_input_grad[0]:
            if subm:
                gx = _implicit_gemm(gh, rb, pack_conv_weight_transposed(weight, True), cout, cin)
            else:
                gx = _implicit_gemm(gh, inverse_rulebook(rb, xh.shape[0]), pack_conv_weight_transposed(weight, False), cout, cin)
            grad_x = (gx.float() / scale).to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            gw = _weight_grad(xh, gh, rb, taps, cin, cout)
            grad_w = (gw / scale).view_as(weight).to(weight.dtype)
        return grad_x, grad_w, None, None


def pack_conv_weight(weight):
    """(kD,kH,kW,Cin,Cout) -> K-major fp16 [Cout, Kpad] with column k*Cin + c (zero padded to 64)."""
    kd, kh, kw, cin, cout = weight.shape
    taps = kd * kh * kw
    w = weight.detach().reshape(taps * cin, cout).t()
    kpad = (taps * cin + 63) // 64 * 64
    out = torch.zeros((cout, kpad), dtype=torch.float16, device=weight.device)
    out[:, :taps * cin] = w.half()
    return out.contiguous()


class SparseModule(nn.Module):
    pass


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 fused_bn=False, **kwargs):
        super().__init__()
        if ndim != 3 or groups != 1 or transposed or inverse:
            raise L.IsfError('sparse conv: only 3-D, groups=1, non-transposed convolutions are built')
        self.ndim, self.in_channels, self.out_channels = ndim, in_channels, out_channels
        self.kernel_size, self.stride = _triple(kernel_size), _triple(stride)
        self.padding, self.dilation = _triple(padding), _triple(dilation)
        self.subm, self.indice_key = subm, indice_key
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self._packed = None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        self._packed = None
        key = prefix + 'weight'
        if key in state_dict:
            w = state_dict[key]
            exp = tuple(self.weight.shape)
            if tuple(w.shape) != exp and tuple(w.shape) == (exp[4], exp[0], exp[1], exp[2], exp[3]):
                state_dict[key] = w.permute(1, 2, 3, 4, 0).contiguous()      # spconv2 layout
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def train(self, mode=True):
        self._packed = None
        return super().train(mode)

    def packed_weight(self):
        if self._packed is None or self._packed.device != self.weight.device:
            self._packed = pack_conv_weight(self.weight)
        return self._packed

    def rulebook(self, x):
        key = self.indice_key
        if key is None:   # the reference rebuilds here (sparse_block.py passes no indice_key); same coordinate
            key = ('auto', self.subm, tuple(self.kernel_size), tuple(self.stride), tuple(self.padding),  # set ->
                   tuple(self.dilation), x.indices.data_ptr(), x.indices.shape[0])                # same rulebook
        rb = x.indice_dict.get(key)
        if rb is None:
            # the same (coordinate set, geometry) under any indice_key is the same rulebook
            geo = ('geo', self.subm, tuple(self.kernel_size), tuple(self.stride), tuple(self.padding),
                   tuple(self.dilation), x.indices.data_ptr(), x.indices.shape[0])
            rb = x.indice_dict.get(geo)
            if rb is None:
                rb = build_rulebook(x, self.kernel_size, self.stride, self.padding, self.dilation, self.subm,
                                    out_capacity=getattr(self, 'out_capacity', None))
                x.indice_dict[geo] = rb
            x.indice_dict[key] = rb
        return rb

    def forward(self, x, scale=None, shift=None, residual=None, relu=False):
        """conv (+ optional fused scale/shift/residual/ReLU epilogue used by the encoder's executor)."""
        rb = self.rulebook(x)
        ev = x.indice_dict.get(('event', id(rb)))
        if ev is not None:      # rulebook built on a side stream (SparseEncoder.prebuild_rulebooks)
            torch.cuda.current_stream().wait_event(ev)
        if self.training and torch.is_grad_enabled() and (x.features.requires_grad or self.weight.requires_grad):
            if scale is not None or shift is not None or residual is not None or relu:
                raise L.IsfError('sparse conv: the fused epilogue is an eval-mode feature')
            if rb.num_dev is not None:
                raise L.IsfError('sparse conv: training runs in exact mode (no capacity buffers)')
            f = ImplicitGemmFunction.apply(x.features, self.weight, rb, self.subm)
            if self.bias is not None:
                f = f + self.bias.to(f.dtype)
            out = SparseConvTensor(f, rb.out_indices, rb.out_shape, x.batch_size)
            out.indice_dict = x.indice_dict
            out.sorted = x.sorted if self.subm else True
            return out
        feats = x.features if x.features.dtype == torch.float16 else x.features.half()
        if self.bias is not None:
            shift = self.bias.float() if shift is None else shift + self.bias.float() * (scale if scale is not None else 1)
        f = implicit_gemm(feats.contiguous(), rb, self.packed_weight(), self.in_channels, self.out_channels, scale,
                          shift, residual, relu)
        out = SparseConvTensor(f, rb.out_indices, rb.out_shape, x.batch_size, num_dev=rb.num_dev)
        out.indice_dict = x.indice_dict
        out.sorted = x.sorted if self.subm else True      # strided outputs are sorted-unique by construction
        return out


@CONV_LAYERS.register_module('SubMConv3d')
class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key)


@CONV_LAYERS.register_module('SparseConv3d')
class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, False,
                         indice_key=indice_key)


def fold_bn(bn):
    """eval-mode BN -> fp32 (scale, shift); cached on the module, invalidated when its buffers change version."""
    key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, bn.weight.device)
    cached = bn.__dict__.get('_isf_fold')
    if cached is not None and cached[0] == key:
        return cached[1]
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float()
    shift = (bn.bias - bn.running_mean * scale).detach().float()
    out = (scale.contiguous(), shift.contiguous())
    bn.__dict__['_isf_fold'] = (key, out)
    return out


class SparseSequential(SparseModule, nn.Sequential):
    """ref: spconv modules.SparseSequential.  In eval mode the pattern conv -> BatchNorm1d -> ReLU is executed as
    one fused kernel (BN folded into the conv epilogue)."""

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, SparseConvolution):
                if (not self.training and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d)):
                    scale, shift = fold_bn(mods[i + 1])
                    relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                    x = m(x, scale=scale, shift=shift, relu=relu)
                    i += 3 if relu else 2
                    continue
                x = m(x)
            elif isinstance(m, SparseModule):
                x = m(x)
            else:
                x = x.replace_feature(m(x.features))
            i += 1
        return x
