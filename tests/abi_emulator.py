"""A software stand-in for libisfusion_b200.so, for CPU tests of the product's HOST logic.

TEST INFRASTRUCTURE ONLY.  `install(monkeypatch)` replaces `isfusion_b200._lib.lib()` by an object whose `isf_*` methods
are torch statements of what each entry point computes according to include/isfusion_b200.h, makes `_lib.ptr` hand the
tensor itself over (instead of its device address) and neutralises the CUDA-only plumbing (streams, workspaces, the
`is_cuda` checks, cuDNN's fused conv).  The product's Python modules then run UNMODIFIED on the CPU: packing, layouts,
token orders, the transposed-kernel trick, epilogue arguments, row scatters ... everything except the kernels themselves,
which the GPU suite covers.  Only the entry points of the fusion encoder / SECONDV2 / SECONDFPN path are emulated.
"""
import ctypes

import torch
import torch.nn.functional as F


def _val(x):
    return x.value if hasattr(x, 'value') else x


def _set_size(byref_obj, n):
    byref_obj._obj.value = n


class FakeLib:
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
        assert num_dev is None and out.shape[-1] == ldo and feat.shape == (bs, ncam, fh, fw, c)
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


def _conv_cpu(x_nhwc, packed, stride=1, relu=True):
    w, b = packed
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2).float(), w.float(), b.float(), stride=stride, padding=1)
    return (torch.relu(y) if relu else y).permute(0, 2, 3, 1).half().contiguous()


def install(monkeypatch):
    """patch isfusion_b200 so that its modules run on CPU tensors against FakeLib; returns the FakeLib."""
    import isfusion_b200._lib as L
    import isfusion_b200.fusion_encoder as FE
    fake = FakeLib()
    monkeypatch.setattr(L, 'lib', lambda: fake)
    monkeypatch.setattr(L, 'ptr', lambda t: t)
    monkeypatch.setattr(L, 'stream', lambda: None)
    monkeypatch.setattr(L, 'check', lambda rc: None if rc == 0 else (_ for _ in ()).throw(L.IsfError('emulator rc %r' % (rc,))))
    monkeypatch.setattr(L, 'require_cuda', lambda *a: None)
    monkeypatch.setattr(L, 'is_device', lambda t: True)
    monkeypatch.setattr(L, 'workspace', lambda nbytes, device: torch.empty(max(int(nbytes), 1), dtype=torch.uint8))
    monkeypatch.setattr(FE, '_conv', _conv_cpu)
    return fake
