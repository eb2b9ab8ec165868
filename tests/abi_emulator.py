"""A software stand-in for libisfusion_b200.so, for CPU tests of the product's HOST logic.

TEST INFRASTRUCTURE ONLY.  `install(monkeypatch)` replaces `isfusion_b200._lib.lib()` by an object whose `isf_*` methods
are torch statements of what each entry point computes according to include/isfusion_b200.h, makes `_lib.ptr` hand the
tensor itself over (instead of its device address) and neutralises the CUDA-only plumbing (streams, workspaces, the
`is_cuda` checks, cuDNN's fused conv).  The product's Python modules then run UNMODIFIED on the CPU: packing, layouts,
token orders, the transposed-kernel trick, epilogue arguments, row scatters ... everything except the kernels themselves,
which the GPU suite covers.  Emulated: the entry points of the fusion encoder / SECONDV2 / SECONDFPN path and of the sparse
encoder with bitmap-rank rulebooks (incl. the opt-in tile order), voxelization and the two-call DynamicVFE.
"""
import ctypes
import os
import sys

import torch
import torch.nn.functional as F


sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'analysis'))      # tile_occupancy (order model)


_PARAM_NAMES = {}
# untyped (void*) data pointers are fp16 matrices unless listed here: scratch / opaque bytes, or either precision (flagged)
_VOID_ANY = {'workspace', 'table', 'in_table', 'stream', 'value', 'plan_dev'}
_VOID_ANY_FN = {('isf_image_aug_normalize', 'out')}          # fp32 or fp16 by flag


def _header_params():
    """{entry point: [C type of every parameter]} parsed from include/isfusion_b200.h."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r'/\*.*?\*/', '', open(os.path.join(root, 'include', 'isfusion_b200.h')).read(), flags=re.S)
    out = {}
    for m in re.finditer(r'\b(?:int|long long|unsigned long long|const char\*)\s+(isf_\w+)\s*\(([^;]*?)\)\s*;', hdr, flags=re.S):
        params = [p.strip() for p in m.group(2).split(',')] if m.group(2).strip() not in ('', 'void') else []
        out[m.group(1)] = [re.sub(r'\s*\w+$', '', p) for p in params]          # drop the parameter name
        _PARAM_NAMES[m.group(1)] = [re.search(r'(\w+)$', p).group(1) for p in params]
    return out


_ELEM = {'float': (torch.float32,), 'int': (torch.int32,), 'unsigned int': (torch.int32, torch.uint32),
         'long long': (torch.int64,), 'unsigned long long': (torch.int64, torch.uint64), 'unsigned char': (torch.uint8,)}
_CELEM = {'float': (ctypes.c_float,), 'int': (ctypes.c_int,), 'long long': (ctypes.c_longlong,), 'double': (ctypes.c_double,)}


def _check_kinds(name, ctypes_of, args):
    """every argument must be of the KIND the header declares: a pointer gets a tensor / ctypes array / byref / None,
    an integer an int, a float a (ctypes) float — catches swapped or shifted arguments that ctypes itself would accept."""
    for i, (ct, a) in enumerate(zip(ctypes_of, args)):
        is_ptr = '*' in ct
        ok_ptr = a is None or isinstance(a, (torch.Tensor, ctypes.Array, bytes)) or type(a).__name__ == 'CArgObject' \
            or isinstance(a, ctypes.c_void_p)
        if is_ptr:
            assert ok_ptr, '%s: argument %d (%s) got %r' % (name, i, ct, type(a))
            base = ct.replace('const', '').replace('*', '').strip()
            if isinstance(a, torch.Tensor) and base in _ELEM:            # typed pointer: the tensor's element type must be it
                assert a.dtype in _ELEM[base], '%s: argument %d is %s but the tensor holds %s' % (name, i, ct, a.dtype)
            elif isinstance(a, torch.Tensor) and base == 'void' and _PARAM_NAMES.get(name) and _PARAM_NAMES[name][i] not in _VOID_ANY \
                    and (name, _PARAM_NAMES[name][i]) not in _VOID_ANY_FN:
                want = torch.bfloat16 if name.endswith('_bf16') else torch.float16       # the `_bf16` twins take bf16 matrices
                assert a.dtype == want, '%s: %s is a %s matrix in the library but the tensor holds %s' % (
                    name, _PARAM_NAMES[name][i], want, a.dtype)
            elif isinstance(a, ctypes.Array) and base in _CELEM:
                assert a._type_ in _CELEM[base], '%s: argument %d is %s but the host array holds %s' % (name, i, ct, a._type_)
        elif ct in ('float', 'double'):
            assert isinstance(a, (float, ctypes.c_float, ctypes.c_double)), '%s: argument %d (%s) got %r' % (name, i, ct, type(a))
        else:
            assert isinstance(a, (int, bool, ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t)) and not isinstance(a, torch.Tensor), \
                '%s: argument %d (%s) got %r' % (name, i, ct, type(a))


def _val(x):
    return x.value if hasattr(x, 'value') else x


def _set_size(byref_obj, n):
    byref_obj._obj.value = n


class _FusionLib:
    # ------------------------------------------------------------------ plumbing
    def isf_last_error(self):
        return b'(emulator)'

    def isf_launch_count(self, reset):
        return 0

    def isf_heatmap_topk_workspace_size(self, batch, classes, h, w, out):
        _set_size(out, 1024)
        return 0

    def isf_channel_attention_workspace_size(self, batch, grid, channels, out):
        _set_size(out, 1024)
        return 0

    # ------------------------------------------------------------------ isf_ins_context_decoder (a11, one launch)
    def isf_ins_context_decoder_sizes(self, layers, nbytes, nfloats):
        _set_size(nbytes, (128 + layers * 1280 + 256) * 128 * 2)
        _set_size(nfloats, 640 + layers * 2048 + 256)
        return 0

    def isf_ins_context_decoder(self, toks, top_idx, value, ldv, wp, fpar, batch, grid, k, c, heads, points, layers, out, kv_out,
                                stream):
        """torch statement of csrc/decoder.cu from the PACKED buffers (fragment order undone), fp32 throughout."""
        assert c == 128 and heads == 8 and points == 16 and k <= 224 and wp.dtype == torch.float16 and fpar.dtype == torch.float32
        n_rows = wp.numel() // 128
        w = wp.view(n_rows // 8, 8, 4, 8, 2, 2).permute(0, 1, 3, 4, 2, 5).reshape(n_rows, 128).float()
        f = fpar
        g, s = grid, grid * grid
        toks = toks.reshape(batch, s, c).float()
        value = value.reshape(batch, s, ldv).float()
        ln = lambda x, gm, bt: F.layer_norm(x, (c,), gm, bt, 1e-5)
        for b in range(batch):
            n = top_idx[b].long()
            ref = torch.stack([((n // g).float() + 0.5) / g, ((n % g).float() + 0.5) / g], -1)
            x = toks[b][n]
            h1 = torch.relu((ref @ f[0:256].view(c, 2).t()) * f[256:384] + f[384:512])
            qpe = h1 @ w[0:128].t() + f[512:640]
            for li in range(layers):
                wl, fl = w[128 + li * 1280:], f[640 + li * 2048:]
                qk = (x + qpe) @ wl[0:256].t() + fl[0:256]
                q, kk = qk[:, :c].view(k, 8, 16), qk[:, c:].view(k, 8, 16)
                v = (x @ wl[256:384].t() + fl[256:384]).view(k, 8, 16)
                att = torch.softmax(torch.einsum('qhd,khd->hqk', q, kk) * 0.25, -1)
                o = torch.einsum('hqk,khd->qhd', att, v).reshape(k, c)
                x = ln(x + o @ wl[384:512].t() + fl[384:512], fl[512:640], fl[640:768])
                soaw = (x + qpe) @ wl[512:896].t() + fl[768:1152]
                so = soaw[:, :256].view(k, 8, 16, 2)
                aw = torch.softmax(soaw[:, 256:].view(k, 8, 16), -1)
                loc = ref[:, None, None, :] + so / g                                        # component 0 = width coordinate
                vmap = value[b, :, li * c:(li + 1) * c].reshape(g, g, 8, 16).permute(2, 3, 0, 1)     # [8,16,Y,X]
                grids = (2 * loc - 1).permute(1, 0, 2, 3)                                   # [8,K,16,2]
                smp = F.grid_sample(vmap, grids, mode='bilinear', padding_mode='zeros', align_corners=False)   # [8,16,K,16]
                ms = (smp * aw.permute(1, 0, 2)[:, None]).sum(-1).permute(2, 0, 1).reshape(k, c)
                x = ln(x + ms @ wl[896:1024].t() + fl[1152:1280], fl[1280:1408], fl[1408:1536])
                hdn = torch.relu(x @ wl[1024:1152].t() + fl[1536:1664])
                x = ln(x + hdn @ wl[1152:1280].t() + fl[1664:1792], fl[1792:1920], fl[1920:2048])
            out[b] = x
            if kv_out is not None:
                wkv, bkv = w[128 + layers * 1280:], f[640 + layers * 2048:]
                kv_out[b] = (x @ wkv[:256].t() + bkv[:256]).half()
        return 0

    # ------------------------------------------------------------------ isf_conv3x3_nhwc_f16
    def isf_conv3x3_nhwc_f16(self, x, batch, h, w, cin, ldx, wpk, cout_pad, shift, stride, relu, out, ldo, cout_store, stream):
        assert cin % 64 == 0 and ldx >= cin and cout_pad in (32, 64, 128, 256) and wpk.shape == (cout_pad, 9 * cin)
        assert stride in (1, 2) and cout_store % 8 == 0 and cout_store <= min(cout_pad, ldo)
        xx = x.reshape(batch, h, w, ldx)[..., :cin].float().permute(0, 3, 1, 2)
        wt = wpk.float().view(cout_pad, 3, 3, cin).permute(0, 3, 1, 2)
        y = F.conv2d(xx, wt, shift, stride=stride, padding=1)
        y = (torch.relu(y) if _val(relu) else y).permute(0, 2, 3, 1)
        o = out.view(batch, y.shape[1], y.shape[2], ldo)
        o[..., :cout_store] = y[..., :cout_store].to(o.dtype)
        return 0

    # ------------------------------------------------------------------ isf_gemm_f16
    def isf_gemm_f16(self, a, w, m, k, kpad, n, scale, shift, residual, ldr, act, ln_g, ln_b, ln_eps, out_f16, out_f32, ldo,
                     out_rows, m_dev, stream):
        assert a.shape[0] >= m and w.shape == (n, kpad) and k % 8 == 0 and kpad % 64 == 0 and m_dev is None
        y = a[:m, :k].float() @ w[:, :k].float().t()
        if scale is not None:
            y = y * scale
        if shift is not None:
            y = y + shift
        if residual is not None:
            assert residual.stride(0) == ldr
            y = y + residual[:m, :n].float()
        if ln_g is not None:
            y = F.layer_norm(y, (n,), ln_g, ln_b, _val(ln_eps))
        act = _val(act)
        y = torch.relu(y) if act == 1 else (F.gelu(y) if act == 2 else y)
        out = out_f16 if out_f16 is not None else out_f32
        assert out.stride(0) == ldo
        if out_rows is not None:
            rows = out_rows[:m].long()
            ok = rows >= 0
            out[rows[ok], :n] = y[ok].to(out.dtype)
        else:
            out[:m, :n] = y.to(out.dtype)
        return 0

    # ------------------------------------------------------------------ isf_cam2bev
    def isf_cam2bev(self, pillars, num_pillars, num_dev, ppp, nfeat, coors, cdim, b_override, feat, proj, bs, ncam, fh, fw, c,
                    img_w, img_h, gh, gw, out, ldo, ch_off, stream):
        assert out.shape[-1] == ldo and feat.shape == (bs, ncam, fh, fw, c)
        if num_dev is not None:                                          # capacity mode: device-side pillar count
            num_pillars = min(int(num_dev.item()), num_pillars)
        img_w, img_h = _val(img_w), _val(img_h)
        pts = pillars[:num_pillars, :, :3].float()
        co = coors[:num_pillars].long()
        bidx = co[:, 0] if cdim == 4 else torch.full((num_pillars,), b_override, dtype=torch.long)
        cy, cx = co[:, cdim - 2], co[:, cdim - 1]
        for b in range(bs):
            sel = bidx == b
            if not sel.any():
                continue
            p = pts[sel].reshape(-1, 3)
            acc = torch.zeros(c, p.shape[0])
            for cam in range(ncam):
                pr = proj[b, cam].float()
                m, t, a, at = pr[:9].view(3, 3), pr[9:12], pr[12:18].view(2, 3), pr[18:20]
                q = p @ m.t() + t
                qz = q[:, 2].clamp(1e-5, 1e5)
                uvz = torch.stack([q[:, 0] / qz, q[:, 1] / qz, qz], 1)
                ua = uvz @ a.t() + at
                grid = torch.stack([(ua[:, 0] / img_w - 0.5) * 2, (ua[:, 1] / img_h - 0.5) * 2], -1)
                fm = feat[b, cam].permute(2, 0, 1).float()[None]
                acc += F.grid_sample(fm, grid.view(1, -1, 1, 2), align_corners=False)[0, :, :, 0]
            cell = acc.view(c, -1, ppp).sum(2).t()                     # [pillars, C]
            out[b, cy[sel], cx[sel], ch_off:ch_off + c] = cell.to(out.dtype)
        return 0

    # ------------------------------------------------------------------ isf_window_attention (SST, 6x6 windows)
    def isf_window_attention(self, qkv, pos_q, pos_k, batch, grid, channels, heads, shifted, out, stream):
        c, d = channels, channels // heads
        s = grid
        off = 3 if shifted else 0
        nwin = (s + off + 5) // 6
        x = qkv.float().view(batch, s, s, 3 * c)
        res = torch.zeros(batch, s, s, c)
        iy, ix = torch.meshgrid(torch.arange(6), torch.arange(6), indexing='ij')
        for wy in range(nwin):
            for wx in range(nwin):
                ys, xs = (wy * 6 - off + iy).flatten(), (wx * 6 - off + ix).flatten()
                ok = (ys >= 0) & (ys < s) & (xs >= 0) & (xs < s)
                if not ok.any():
                    continue
                slots = torch.nonzero(ok).squeeze(1)
                tok = x[:, ys[ok], xs[ok]]                              # [B, n, 3C]
                q = (tok[..., :c] + pos_q[slots]).view(batch, -1, heads, d).transpose(1, 2)
                k = (tok[..., c:2 * c] + pos_k[slots]).view(batch, -1, heads, d).transpose(1, 2)
                v = tok[..., 2 * c:].view(batch, -1, heads, d).transpose(1, 2)
                o = F.scaled_dot_product_attention(q, k, v)             # softmax over the window's valid tokens only
                res[:, ys[ok], xs[ok]] = o.transpose(1, 2).reshape(batch, -1, c)
        out.copy_(res.view(batch * s * s, c).to(out.dtype))
        return 0

    # ------------------------------------------------------------------ isf_heatmap_topk
    def isf_heatmap_topk(self, logits, batch, classes, h, w, plain_from, k, top_idx, top_score, ws, ws_bytes, stream):
        lg = logits.float().view(batch, classes, h, w)
        keep = torch.zeros_like(lg, dtype=torch.bool)
        keep[:, :, 1:-1, 1:-1] = lg[:, :, 1:-1, 1:-1] == F.max_pool2d(lg, 3, 1, 0)
        keep[:, plain_from:] = True
        score = torch.where(keep, lg.sigmoid(), torch.zeros_like(lg)).view(batch, -1)
        order = torch.argsort(score, dim=1, descending=True, stable=True)[:, :k]
        top_idx.copy_((order % (h * w)).int())
        if top_score is not None:
            top_score.copy_(torch.gather(score, 1, order))
        return 0

    # ------------------------------------------------------------------ isf_ms_deform_attn_forward
    def isf_ms_deform_attn_forward(self, value, is_half, shapes, lvl_start, loc, attn, b, s, h, d, lv, q, p, out, stream):
        from oracle.fusion_ops import ms_deform_attn_core_pytorch
        out.copy_(ms_deform_attn_core_pytorch(value.float(), shapes, loc.float(), attn.float()))
        return 0

    def isf_ms_deform_attn_backward(self, value, is_half, shapes, lvl_start, loc, attn, grad_out, b, s, h, d, lv, q, p, grad_value,
                                    grad_loc, grad_attn, stream):
        from oracle.fusion_ops import ms_deform_attn_core_pytorch
        with torch.enable_grad():
            v64, l64, a64 = (t.detach().double().requires_grad_(True) for t in (value, loc, attn))
            y = ms_deform_attn_core_pytorch(v64, shapes, l64, a64)
            gv, gl, ga = torch.autograd.grad(y, (v64, l64, a64), grad_out.double())
        grad_value.add_(gv.float())                 # accumulated into, as the kernel does
        grad_loc.copy_(gl.float())
        grad_attn.copy_(ga.float())
        return 0

    # ------------------------------------------------------------------ isf_i2s_attention / isf_channel_attention
    def isf_i2s_attention(self, q, kv, batch, tokens, nk, channels, heads, out, stream):
        d = channels // heads
        qq = q.float().view(batch, tokens, heads, d).transpose(1, 2)
        kk = kv[..., :channels].float().view(batch, nk, heads, d).transpose(1, 2)
        vv = kv[..., channels:].float().view(batch, nk, heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(qq, kk, vv)
        out.copy_(o.transpose(1, 2).reshape(batch * tokens, channels).to(out.dtype))
        return 0

    def isf_channel_attention(self, scene, lds, ins, ldi, batch, grid, channels, out, ldo, ws, ws_bytes, stream):
        sp = scene.float().reshape(batch, grid, grid, channels).permute(0, 3, 1, 2)
        ip = ins.float().reshape(batch, grid, grid, channels).permute(0, 3, 1, 2)
        o = sp + torch.softmax(sp @ ip.transpose(2, 3), dim=-1) @ ip
        out.copy_(o.permute(0, 2, 3, 1).reshape(out.shape).to(out.dtype))
        return 0


# ---------------------------------------------------------------------- sparse convolution entry points
def _ints(a, n=3):
    return [int(a[i]) for i in range(n)]


def _lin(idx, shape):
    idx = idx.long()
    return ((idx[:, 0] * shape[0] + idx[:, 1]) * shape[1] + idx[:, 2]) * shape[2] + idx[:, 3]


class SparseMixin:
    """bitmap-rank rulebooks (sorted tensors), tile order, implicit GEMM, dense().  The bitmap / prefix arrays are opaque to
    the host code, so the emulator keeps the coordinate set they stand for in a side table keyed by the `bits` tensor."""

    def _maps(self):
        if not hasattr(self, '_map'):
            self._map = {}
        return self._map

    def isf_bitmap_words(self, shape, batch):
        s = _ints(shape)
        return (int(batch) * s[0] * s[1] * s[2] + 31) // 32

    def isf_bitmap_workspace_size(self, words, out):
        _set_size(out, 1024)
        return 0

    def isf_bitmap_build(self, indices, v, v_dev, shape, batch, bits, prefix, ws, ws_bytes, stream):
        n = int(v_dev.item()) if v_dev is not None else v
        self._maps()[bits.data_ptr()] = (indices[:n].clone(), _ints(shape))
        return 0

    def isf_spconv_out_shape(self, in_shape, ks, st, pd, dl, out_shape):
        i, k, s, p, d = _ints(in_shape), _ints(ks), _ints(st), _ints(pd), _ints(dl)
        for a in range(3):
            out_shape[a] = (i[a] + 2 * p[a] - d[a] * (k[a] - 1) - 1) // s[a] + 1
        return 0

    def isf_bitmap_out_sites(self, in_indices, v_in, v_in_dev, in_shape, batch, ks, st, pd, dl, out_bits, out_prefix, out_indices,
                             out_cap, num_out_dev, ws, ws_bytes, stream):
        from oracle import spconv_ops as so
        n = int(v_in_dev.item()) if v_in_dev is not None else v_in
        oi, oshape, _ = so.get_indice_pairs(in_indices[:n], batch, _ints(in_shape), _ints(ks), _ints(st), _ints(pd), _ints(dl), False)
        cnt = min(oi.shape[0], out_cap)
        out_indices[:cnt] = oi[:cnt]
        num_out_dev[0] = cnt
        self._maps()[out_bits.data_ptr()] = (oi[:cnt].clone(), list(oshape))
        return 0

    # hash-table rulebooks (tensors of unknown row order): the table is opaque too
    def isf_hash_capacity(self, n):
        c = 1024
        while c < 2 * n:
            c <<= 1
        return c

    def isf_hash_build(self, indices, v, v_dev, shape, batch, table, capacity, stream):
        n = int(v_dev.item()) if v_dev is not None else v
        self._maps()[table.data_ptr()] = (indices[:n].clone(), _ints(shape))
        return 0

    def isf_spconv_out_indices_workspace_size(self, v_in, kvol, out_cap, out):
        _set_size(out, 1024)
        return 0

    def isf_spconv_out_indices(self, in_indices, v_in, v_in_dev, in_shape, batch, ks, st, pd, dl, out_indices, out_cap, num_out_dev,
                               ws, ws_bytes, stream):
        from oracle import spconv_ops as so
        n = int(v_in_dev.item()) if v_in_dev is not None else v_in
        oi, _, _ = so.get_indice_pairs(in_indices[:n], batch, _ints(in_shape), _ints(ks), _ints(st), _ints(pd), _ints(dl), False)
        cnt = min(oi.shape[0], out_cap)
        out_indices[:cnt] = oi[:cnt]
        num_out_dev[0] = cnt
        return 0

    def isf_spconv_neighbors(self, out_indices, v_out, v_dev, table, capacity, in_shape, ks, st, pd, dl, subm, nbr, nbr_stride, stream):
        n = int(v_dev.item()) if v_dev is not None else v_out
        nbr[:, :n] = self._neighbours(out_indices, n, table, in_shape, ks, st, pd, dl, bool(subm), need_sorted=False)
        return 0

    def _neighbours(self, out_indices, n, in_bits, in_shape, ks, st, pd, dl, subm, need_sorted=True):
        in_idx, shape = self._maps()[in_bits.data_ptr()]
        assert shape == _ints(in_shape)
        ks, st, pd, dl = _ints(ks), ([1, 1, 1] if subm else _ints(st)), None if subm else _ints(pd), _ints(dl)
        key = _lin(in_idx, shape)
        order = torch.arange(len(key))
        if need_sorted:
            assert (key[1:] > key[:-1]).all(), 'bitmap-rank rulebooks need rows in ascending key order'
        else:
            key, order = torch.sort(key)
        oi = out_indices[:n].long()
        taps = ks[0] * ks[1] * ks[2]
        nbr = torch.full((taps, n), -1, dtype=torch.int32)
        k = 0
        for kz in range(ks[0]):
            for ky in range(ks[1]):
                for kx in range(ks[2]):
                    kk = torch.tensor([kz, ky, kx])
                    pad = torch.tensor([(ks[d] // 2) * dl[d] for d in range(3)]) if subm else torch.tensor(pd)
                    ip = oi[:, 1:] * torch.tensor(st) - pad + kk * torch.tensor(dl)
                    ok = ((ip >= 0) & (ip < torch.tensor(shape))).all(1)
                    q = _lin(torch.cat([oi[:, :1], ip.clamp(min=0)], 1), shape)
                    pos = torch.searchsorted(key, q).clamp(max=max(len(key) - 1, 0))
                    hit = ok & (key[pos] == q) if len(key) else torch.zeros_like(ok)
                    nbr[k, hit] = order[pos[hit]].int()
                    k += 1
        return nbr

    def isf_spconv_neighbors_bitmap(self, out_indices, v_out, v_dev, in_bits, in_prefix, in_shape, ks, st, pd, dl, subm, nbr,
                                    nbr_stride, stream):
        n = int(v_dev.item()) if v_dev is not None else v_out
        nbr[:, :n] = self._neighbours(out_indices, n, in_bits, in_shape, ks, st, pd, dl, bool(subm))
        return 0

    def isf_spconv_tile_order_workspace_size(self, v, out):
        _set_size(out, 1024)
        return 0

    def _order(self, nbr_valid_rows, chunk_log2, ksize):
        import tile_occupancy as to
        return torch.from_numpy(to.tile_order_model((nbr_valid_rows >= 0).numpy(), int(chunk_log2), tuple(_ints(ksize)))).long()

    def isf_spconv_tile_order(self, nbr, nbr_stride, taps, ksize, v, v_dev, chunk_log2, nbr_out, row_perm, ws, ws_bytes, stream):
        n = min(int(v_dev.item()), v) if v_dev is not None else v
        perm = self._order(nbr[:, :n], chunk_log2, ksize)
        nbr_out.fill_(-1)
        row_perm.fill_(-1)
        nbr_out[:, :n] = nbr[:, perm]
        row_perm[:n] = perm.int()
        return 0

    def isf_spconv_neighbors_bitmap_ordered(self, out_indices, v_out, v_dev, in_bits, in_prefix, in_shape, ks, st, pd, dl, subm,
                                            chunk_log2, nbr_out, nbr_stride, row_perm, ws, ws_bytes, stream):
        n = min(int(v_dev.item()), v_out) if v_dev is not None else v_out
        nbr = self._neighbours(out_indices, n, in_bits, in_shape, ks, st, pd, dl, bool(subm))
        perm = self._order(nbr, chunk_log2, ks)
        nbr_out.fill_(-1)
        row_perm.fill_(-1)
        nbr_out[:, :n] = nbr[:, perm]
        row_perm[:n] = perm.int()
        return 0

    def _implicit(self, feats, cin, w, kpad, cout, nbr, taps, v_out, v_dev, row_perm, scale, shift, residual, relu, out):
        n = min(int(v_dev.item()), v_out) if v_dev is not None else v_out
        assert w.shape == (cout, kpad) and kpad >= taps * cin
        acc = torch.zeros((n, cout))
        f = feats.reshape(-1, cin).float()               # the kernel sees a flat matrix of cin-wide rows
        for k in range(taps):
            r = nbr[k, :n].long()
            hit = r >= 0
            if hit.any():
                acc[hit] += f[r[hit]] @ w[:, k * cin:(k + 1) * cin].float().t()
        if scale is not None:
            acc = acc * scale
        if shift is not None:
            acc = acc + shift
        rows = row_perm[:n].long() if row_perm is not None else torch.arange(n)
        assert (rows >= 0).all()
        if residual is not None:
            acc = acc + residual[rows].float()
        if relu:
            acc = torch.relu(acc)
        out[rows] = acc.to(out.dtype)
        return 0

    def isf_spconv_tile_counter(self, counter2):
        assert counter2 is None or (counter2.dtype == torch.int32 and counter2.numel() == 2 and int(counter2.abs().sum()) == 0)
        return 0

    def isf_spconv_wgrad(self, x, cin, g, cout, in_rows, out_rows, tap_offsets, taps, max_pairs, dw, stream):
        assert x.dtype == torch.float16 and g.dtype == torch.float16 and dw.dtype == torch.float32
        assert in_rows.dtype == torch.int32 and out_rows.dtype == torch.int32 and tap_offsets.numel() == taps + 1
        assert tuple(dw.shape) == (taps, cin, cout) and cin % 8 == 0 and cout % 8 == 0
        offs = tap_offsets.tolist()
        assert max(b - a for a, b in zip(offs[:-1], offs[1:])) <= max_pairs
        for k in range(taps):
            a, b = offs[k], offs[k + 1]
            if b > a:
                dw[k] += x[in_rows[a:b].long()].float().t() @ g[out_rows[a:b].long()].float()
        return 0

    def isf_spconv_implicit_gemm(self, feats, cin, w, kpad, cout, nbr, nbr_stride, taps, v_out, v_dev, scale, shift, residual, ldr,
                                 relu, out, ldo, stream):
        return self._implicit(feats, cin, w, kpad, cout, nbr, taps, v_out, v_dev, None, scale, shift, residual, relu, out)

    def isf_spconv_implicit_gemm_ordered(self, feats, cin, w, kpad, cout, nbr, nbr_stride, taps, v_out, v_dev, row_perm, scale, shift,
                                         residual, ldr, relu, out, ldo, stream):
        return self._implicit(feats, cin, w, kpad, cout, nbr, taps, v_out, v_dev, row_perm, scale, shift, residual, relu, out)

    def isf_sparse_to_dense(self, feats, indices, v, c, batch, d, h, w, out, stream):
        out.zero_()
        i = indices[:v].long()
        out[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]] = feats[:v].to(out.dtype)
        return 0


class VoxelMixin:
    """voxelization, the two-call DynamicVFE (group / compute), the BEV view of the sparse output."""

    def isf_dynamic_voxelize(self, points, n, nfeat, voxel_size, coors_range, coors, cpu_semantics, stream):
        from oracle import voxel_ops
        vs, cr = [float(voxel_size[i]) for i in range(3)], [float(coors_range[i]) for i in range(6)]
        coors[:n] = voxel_ops.dynamic_voxelize(points[:n], vs, cr, mode='cpu' if cpu_semantics else 'cuda')
        return 0

    def isf_hard_voxelize_workspace_size(self, n, out):
        _set_size(out, 1024)
        return 0

    def isf_hard_voxelize(self, points, n, nfeat, voxel_size, coors_range, max_points, max_voxels, voxels, coors, npv, cnt, ws,
                          ws_bytes, stream):
        from oracle import voxel_ops
        vs, cr = [float(voxel_size[i]) for i in range(3)], [float(coors_range[i]) for i in range(6)]
        v, c, k = voxel_ops.hard_voxelize(points[:n], vs, cr, max_points, max_voxels)
        m = v.shape[0]
        voxels[:m], coors[:m], npv[:m] = v, c, k
        cnt[0] = m
        return 0

    def isf_dynamic_scatter_workspace_size(self, n, out):
        _set_size(out, 1024)
        return 0

    def isf_dynamic_point_to_voxel_forward(self, feats, coors, n, c, ndim, dims, reduce_type, reduced, out_coors, coors_map, rcount,
                                           nv, ws, ws_bytes, stream):
        from oracle import voxel_ops
        red, oc, cmap, cnt = voxel_ops.dynamic_scatter(feats[:n], coors[:n], {0: 'sum', 1: 'mean', 2: 'max'}[reduce_type])
        v = red.shape[0]
        reduced[:v], out_coors[:v], rcount[:v] = red, oc, cnt
        coors_map[:n] = cmap
        nv[0] = v
        return 0

    def isf_dynamic_scatter_backward_workspace_size(self, v, c, bytes_ref):
        _set_size(bytes_ref, 256)
        return 0

    def isf_dynamic_point_to_voxel_backward(self, grad_feats, grad_reduced, feats, reduced, cmap, count, n, v, c, reduce_type, ws,
                                            ws_bytes, stream):
        from oracle import voxel_ops as vo
        grad_feats.copy_(vo.dynamic_scatter_backward(grad_reduced, feats, reduced, cmap, count, {0: 'sum', 1: 'mean', 2: 'max'}[reduce_type]))
        return 0

    def isf_merge_sweeps_workspace_size(self, m, bytes_ref):
        _set_size(bytes_ref, 256)
        return 0

    def isf_merge_sweeps(self, raw, m, load_dim, nseg, seg_start, rot, trans, lag, seg_flags, radius, rng, out, count, ws, ws_bytes,
                         stream):
        from oracle import input_ops
        import numpy as np
        r = raw.numpy()
        R = np.array(list(rot), dtype=np.float64).reshape(nseg, 3, 3)
        t = np.array(list(trans), dtype=np.float64).reshape(nseg, 3)
        parts = []
        for s in range(nseg):
            ps = np.copy(r[seg_start[s]:seg_start[s + 1]])
            if seg_flags[s] & 2:
                ps = input_ops.remove_close(ps, _val(radius))
            if seg_flags[s] & 1:                               # the three numpy statements of loading.py:886-888
                ps[:, :3] = ps[:, :3] @ R[s].T
                ps[:, :3] += t[s]
            ps[:, 4] = lag[s]
            parts.append(ps[:, :5])
        res = np.concatenate(parts, 0)
        if rng is not None:                                    # BasePoints.in_range_3d on the merged cloud
            r = np.array(list(rng), dtype=np.float32)
            res = res[(res[:, 0] > r[0]) & (res[:, 1] > r[1]) & (res[:, 2] > r[2]) & (res[:, 0] < r[3]) & (res[:, 1] < r[4]) & (res[:, 2] < r[5])]
        out[:res.shape[0]].copy_(torch.from_numpy(res))
        count.fill_(res.shape[0])
        return 0

    def isf_image_plan_size(self, in_w, in_h, rs_w, rs_h, out_w, out_h, bytes_ref):
        _set_size(bytes_ref, 64)
        return 0

    def isf_image_plan(self, in_w, in_h, rs_w, rs_h, crop_x, crop_y, out_w, out_h, flip, plan, plan_bytes, rows, stream):
        from oracle import input_ops
        plan[:3] = torch.tensor([crop_x, crop_y, flip], dtype=torch.int32)        # what the emulated kernel needs of the tables
        ymin, ycnt, _ = input_ops.resample_coeffs(in_h, rs_h) if rs_h != in_h else (list(range(in_h)), [1] * in_h, None)
        valid = [oy + crop_y for oy in range(out_h) if 0 <= oy + crop_y < rs_h]
        rows[0] = min([int(ymin[r]) for r in valid], default=0)
        rows[1] = max([int(ymin[r] + ycnt[r]) for r in valid], default=1)
        return 0

    def isf_image_aug_workspace_size(self, n, out_w, row_lo, row_hi, bytes_ref):
        assert 0 <= row_lo < row_hi
        _set_size(bytes_ref, 256)
        return 0

    def isf_image_aug_normalize(self, img, n, in_w, in_h, rs_w, rs_h, out_w, out_h, row_lo, row_hi, plan, mean, std, out, is_half,
                                ws, ws_bytes, stream):
        from oracle import input_ops
        cx, cy, flip = (int(v) for v in plan[:3])
        assert tuple(img.shape) == (n, in_h, in_w, 3) and 0 <= row_lo < row_hi <= in_h
        for i in range(n):
            r = input_ops.image_aug_normalize(img[i].numpy(), (rs_w, rs_h), (cx, cy, cx + out_w, cy + out_h), bool(flip), list(mean), list(std))
            out[i].copy_(torch.from_numpy(r).to(out.dtype))
        return 0

    def isf_ingroup_indices(self, group_inds, n, key_bits, out_inds, ws, ws_bytes, stream):
        g = group_inds[:n]
        order = torch.argsort(g, stable=True)
        gs = g[order]
        start = torch.ones_like(gs, dtype=torch.bool)
        start[1:] = gs[1:] != gs[:-1]
        first = torch.nonzero(start).squeeze(1)
        out_inds[order] = torch.arange(n) - first[torch.cumsum(start.long(), 0) - 1]
        return 0

    def isf_dynamic_vfe_workspace_size(self, n, out):
        _set_size(out, 1024)
        return 0

    isf_dynamic_vfe_tc_workspace_size = isf_dynamic_vfe_workspace_size

    def isf_dynamic_vfe_group(self, coors, n, dims, vcoors, nv, ws, ws_bytes, stream):
        c = coors[:n].long()
        d = [int(dims[i]) for i in range(4)]
        assert ((c >= 0) & (c < torch.tensor(d))).all(), 'DynamicVFE takes in-range points (SURVEY Appendix C #5)'
        key = ((c[:, 0] * d[1] + c[:, 1]) * d[2] + c[:, 2]) * d[3] + c[:, 3]
        uniq, inv = torch.unique(key, return_inverse=True)                    # ascending = sorted-unique (b,z,y,x) order
        v = uniq.numel()
        vc = torch.stack([uniq // (d[1] * d[2] * d[3]), (uniq // (d[2] * d[3])) % d[1], (uniq // d[3]) % d[2], uniq % d[3]], 1)
        vcoors[:v] = vc.int()
        nv[0] = v
        if not hasattr(self, '_groups'):
            self._groups = {}
        self._groups[ws.data_ptr()] = (inv, vc, v)
        return 0

    def _vfe(self, points, n, voxel_size, offsets, w1, s1, b1, w2, s2, b2, ws):
        inv, vc, v = self._groups[ws.data_ptr()]
        p = points[:n].float()
        c = p.shape[1]
        seg_mean = torch.zeros(v, 3).index_add_(0, inv, p[:, :3]) / torch.bincount(inv, minlength=v)[:, None]
        vs, off = torch.tensor([float(voxel_size[i]) for i in range(3)]), torch.tensor([float(offsets[i]) for i in range(3)])
        centre = torch.stack([vc[inv, 3] * vs[0] + off[0], vc[inv, 2] * vs[1] + off[1], vc[inv, 1] * vs[2] + off[2]], 1)
        f = torch.cat([p[:, :c], p[:, :3] - seg_mean[inv], p[:, :3] - centre], 1)                  # [N, 11]
        h1 = torch.relu((f @ w1.float().t()) * s1 + b1)
        idx = inv[:, None].expand(-1, h1.shape[1])
        vmax = torch.full((v, h1.shape[1]), float('-inf')).scatter_reduce(0, idx, h1, 'amax')
        h2 = torch.relu((torch.cat([h1, vmax[inv]], 1) @ w2.float().t()) * s2 + b2)
        return torch.full((v, h2.shape[1]), float('-inf')).scatter_reduce(0, inv[:, None].expand(-1, h2.shape[1]), h2, 'amax'), v

    def isf_dynamic_vfe_compute_tc(self, points, n, nfeat, dims, voxel_size, offsets, w1, s1, b1, w2a, w2b, s2, b2, feats, ws,
                                   ws_bytes, stream):
        out, v = self._vfe(points, n, voxel_size, offsets, w1, s1, b1, torch.cat([w2a.float()[:, :64], w2b.float()[:, :64]], 1),
                           s2, b2, ws)
        feats[:v] = out.to(feats.dtype)
        return 0

    def isf_dynamic_vfe_compute(self, points, n, nfeat, dims, voxel_size, offsets, w1, s1, b1, w2, s2, b2, in_ch, feat_ch, f32, f16,
                                ws, ws_bytes, stream):
        out, v = self._vfe(points, n, voxel_size, offsets, w1, s1, b1, w2, s2, b2, ws)
        dst = f16 if f16 is not None else f32
        dst[:v] = out.to(dst.dtype)
        return 0

    def isf_sparse_to_bev_nhwc(self, feats, indices, v, v_dev, c, d, h, w, out, ldo, ch_offset, stream):
        n = min(int(v_dev.item()), v) if v_dev is not None else v
        i = indices[:n].long()
        ch = ch_offset + torch.arange(c)[None, :] * d + i[:, 1:2]                                 # channel c*D + z
        out[i[:, 0:1].expand(-1, c), i[:, 2:3].expand(-1, c), i[:, 3:4].expand(-1, c), ch] = feats[:n].to(out.dtype)
        return 0


class RowsAroundMixin:
    """entry points of the rows around the path: head (top-K with class, query -> scene attention) and Swin window attention."""

    def isf_heatmap_topk_class(self, logits, batch, classes, h, w, plain_from, k, top_idx, top_score, ws, ws_bytes, stream):
        lg = logits.float().view(batch, classes, h, w)
        keep = torch.zeros_like(lg, dtype=torch.bool)
        keep[:, :, 1:-1, 1:-1] = lg[:, :, 1:-1, 1:-1] == F.max_pool2d(lg, 3, 1, 0)
        keep[:, plain_from:] = True
        score = torch.where(keep, lg.sigmoid(), torch.zeros_like(lg)).view(batch, -1)
        top_idx.copy_(torch.argsort(score, dim=1, descending=True, stable=True)[:, :k].int())
        return 0

    def isf_query_scene_attention_workspace_size(self, batch, queries, nk, channels, heads, out):
        _set_size(out, 1024)
        return 0

    def isf_query_scene_attention(self, q, kv, batch, queries, nk, channels, heads, out, ws, ws_bytes, stream):
        assert q.shape == (batch, queries, channels) and kv.shape == (batch, nk, 2 * channels)
        d = channels // heads
        qq = q.float().view(batch, queries, heads, d).transpose(1, 2)
        kk = kv[..., :channels].float().view(batch, nk, heads, d).transpose(1, 2)
        vv = kv[..., channels:].float().view(batch, nk, heads, d).transpose(1, 2)
        out.copy_(F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2).reshape(batch, queries, channels).to(out.dtype))
        return 0

    def isf_swin_window_attention(self, qkv, qkv_bias, rel_bias, batch, h, w, c, heads, shift, out, stream):
        """the reference's sequence (swin.py:206-283) on the token-major qkv matrix; padding positions carry the bias."""
        ws, d = 7, c // heads
        assert qkv.shape == (batch * h * w, 3 * c) and rel_bias.shape == (heads, 49, 49) and qkv_bias.shape == (3 * c,)
        x = qkv.float().view(batch, h, w, 3 * c)
        hp, wp = (h + ws - 1) // ws * ws, (w + ws - 1) // ws * ws
        full = qkv_bias.float().view(1, 1, 1, -1).expand(batch, hp, wp, -1).clone()
        full[:, :h, :w] = x
        mask = None
        if shift > 0:
            full = torch.roll(full, shifts=(-shift, -shift), dims=(1, 2))
            region = torch.zeros(hp, wp)
            cnt = 0
            for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
                for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
                    region[hs, wsl] = cnt
                    cnt += 1
            mw = region.view(hp // ws, ws, wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, 49)
            diff = mw.unsqueeze(1) - mw.unsqueeze(2)
            mask = torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))[None, :, None]          # [1, nW, 1, N, N]
        win = full.view(batch, hp // ws, ws, wp // ws, ws, 3 * c).permute(0, 1, 3, 2, 4, 5).reshape(batch, -1, 49, 3, heads, d)
        q, k, v = (win[:, :, :, i].permute(0, 1, 3, 2, 4) for i in range(3))                                 # [B, nW, heads, N, d]
        m = rel_bias.float()[None, None]
        if mask is not None:
            m = m + mask
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=m.expand(batch, -1, -1, -1, -1))
        o = o.permute(0, 1, 3, 2, 4).reshape(batch, hp // ws, wp // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(batch, hp, wp, c)
        if shift > 0:
            o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
        out.copy_(o[:, :h, :w].reshape(batch * h * w, c).to(out.dtype))
        return 0


def _bf16_twin(f16_name, tensor_args):
    """`_bf16` entry point = the statement of its f16 sibling (written on .float() copies, storing in the output's own 16-bit
    type) with the SAME parameter list, after checking that the 16-bit tensors really are bf16."""
    import inspect
    f = getattr(_FusionLib, f16_name)
    names = list(inspect.signature(f).parameters)[1:]

    def twin(self, *args):
        for n in tensor_args:
            t = args[names.index(n)]
            assert t is None or t.dtype in (torch.bfloat16, torch.float32), '%s_bf16: %s holds %s' % (f16_name, n, t.dtype)
        return f(self, *args)
    twin.__signature__ = inspect.signature(f)
    return twin


_FusionLib.isf_gemm_bf16 = _bf16_twin('isf_gemm_f16', ('a', 'w', 'residual', 'out_f16'))
_FusionLib.isf_conv3x3_nhwc_bf16 = _bf16_twin('isf_conv3x3_nhwc_f16', ('x', 'wpk', 'out'))
_FusionLib.isf_window_attention_bf16 = _bf16_twin('isf_window_attention', ('qkv', 'out'))
_FusionLib.isf_cam2bev_bf16 = _bf16_twin('isf_cam2bev', ('feat', 'out'))


class FakeLib(_FusionLib, SparseMixin, VoxelMixin, RowsAroundMixin):
    _KINDS = None
    HOST_ONLY = ('isf_spconv_tile_order_bits', 'isf_debug_nbr_rows_host', 'isf_debug_swin_window_map', 'isf_debug_msda_backward_host', 'isf_debug_merge_sweeps_host', 'isf_debug_image_aug_host', 'isf_set_option',
                 'isf_version')

    def __init__(self, real=None):
        self._real = real

    def __getattribute__(self, name):       # every emulated call is checked against the header's parameter kinds first
        attr = object.__getattribute__(self, name)
        if name.startswith('isf_') and callable(attr):
            kinds = FakeLib._KINDS.get(name) if FakeLib._KINDS is not None else None
            if FakeLib._KINDS is None:
                FakeLib._KINDS = _header_params()
                kinds = FakeLib._KINDS.get(name)
            if kinds is not None:
                def checked(*args, _f=attr, _k=kinds, _n=name):
                    assert len(args) == len(_k), '%s: %d arguments, the header declares %d' % (_n, len(args), len(_k))
                    _check_kinds(_n, _k, args)
                    return _f(*args)
                return checked
        return attr

    def __getattr__(self, name):           # host-only entry points need no GPU: the real library serves them
        if name in FakeLib.HOST_ONLY and self.__dict__.get('_real') is not None:
            return getattr(self._real, name)
        raise AttributeError('abi_emulator: %s is not emulated' % name)


def _conv_cpu(x_nhwc, packed, stride=1, relu=True):
    w, b = packed[0], packed[1]
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2).float(), w.float(), b.float(), stride=stride, padding=1)
    return (torch.relu(y) if relu else y).permute(0, 2, 3, 1).half().contiguous()


def _ptr_checked(t):
    """L.ptr on the emulator hands the tensor itself over.  A raw pointer carries no strides, so what the kernel would read
    equals what the emulated entry point reads only if the tensor is dense along its last dimension and its leading
    dimensions are either contiguous or a plain row-strided 2-D view (the `ld*` arguments of the ABI describe exactly that)."""
    if t is None or t.numel() <= 1:
        return t
    assert t.stride(-1) == 1 or t.size(-1) == 1, 'pointer to a tensor with a strided last dimension: %s / %s' % (tuple(t.shape), t.stride())
    if not t.is_contiguous():
        lead = t.reshape(-1, t.size(-1)) if t.dim() != 2 else t          # reshape of a non-collapsible view would copy:
        assert lead.data_ptr() == t.data_ptr(), 'pointer to a view that is not row-strided: %s / %s' % (tuple(t.shape), t.stride())
    return t


def _require_contiguous(*tensors):
    """L.require_cuda minus the device test: the contiguity contract still holds on the emulator."""
    import isfusion_b200._lib as L
    for t in tensors:
        if t is not None and not t.is_contiguous():
            raise L.IsfError('expected a contiguous tensor')


ORIGINAL_CONV = None      # fusion_encoder._conv as shipped (set by install; lets a test re-enable it for the own-conv path)


def install(monkeypatch):
    """patch isfusion_b200 so that its modules run on CPU tensors against FakeLib; returns the FakeLib."""
    import isfusion_b200._lib as L
    import isfusion_b200.fusion_encoder as FE
    global ORIGINAL_CONV
    ORIGINAL_CONV = FE._conv
    fake = FakeLib(L.lib())
    monkeypatch.setattr(L, 'lib', lambda: fake)
    monkeypatch.setattr(L, 'ptr', _ptr_checked)
    monkeypatch.setattr(L, 'stream', lambda: None)
    monkeypatch.setattr(L, 'check', lambda rc: None if rc == 0 else (_ for _ in ()).throw(L.IsfError('emulator rc %r' % (rc,))))
    monkeypatch.setattr(L, 'require_cuda', _require_contiguous)
    monkeypatch.setattr(L, 'is_device', lambda t: True)
    monkeypatch.setattr(L, 'workspace', lambda nbytes, device: torch.empty(max(int(nbytes), 1), dtype=torch.uint8))
    return fake
