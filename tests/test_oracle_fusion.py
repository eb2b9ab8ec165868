"""CPU: the fusion-encoder oracle against the reference's golden vectors and its own invariants."""
import os

import numpy as np
import torch

import synth
from oracle import fusion_ops as fo
from oracle import sst_ops as ss


def test_msda_oracle_matches_reference_golden(golden_dir):
    """ms_deform_attn_core_pytorch restatement == the reference function executed from the reference file
    (tests/golden/make_golden.py), on the ops/test.py:20-36 shapes and the IGF head layout."""
    g = np.load(os.path.join(golden_dir, 'msda_ref.npz'))
    out = fo.ms_deform_attn_core_pytorch(torch.from_numpy(g['a_value']).double(), torch.from_numpy(g['a_shapes']),
                                         torch.from_numpy(g['a_loc']).double(), torch.from_numpy(g['a_aw']).double())
    assert torch.allclose(out.float(), torch.from_numpy(g['a_out']), atol=1e-7)
    out = fo.ms_deform_attn_core_pytorch(torch.from_numpy(g['b_value_f16']).double(), torch.tensor([[40, 44]]),
                                         torch.from_numpy(g['b_loc']).double(), torch.from_numpy(g['b_aw']).double())
    assert torch.allclose(out.float(), torch.from_numpy(g['b_out']), atol=1e-5)


def test_msda_oracle_matches_cuda_kernel_spec():
    """the grid_sample form == the explicit bilinear loop of ms_deform_im2col_cuda.cuh:34-83,238-298."""
    torch.manual_seed(0)
    B, H, W, M, D, Q, P = 1, 5, 7, 2, 4, 6, 3
    value = torch.randn(B, H * W, M, D, dtype=torch.float64)
    loc = torch.rand(B, Q, M, 1, P, 2, dtype=torch.float64) * 1.4 - 0.2
    aw = torch.softmax(torch.randn(B, Q, M, P, dtype=torch.float64), -1).view(B, Q, M, 1, P)
    ref = fo.ms_deform_attn_core_pytorch(value, torch.tensor([[H, W]]), loc, aw)
    out = torch.zeros(B, Q, M * D, dtype=torch.float64)
    for q in range(Q):
        for m in range(M):
            for p in range(P):
                w_im, h_im = loc[0, q, m, 0, p, 0] * W - 0.5, loc[0, q, m, 0, p, 1] * H - 0.5
                if not (h_im > -1 and w_im > -1 and h_im < H and w_im < W):
                    continue
                hl, wl = int(np.floor(float(h_im))), int(np.floor(float(w_im)))
                lh, lw = h_im - hl, w_im - wl
                acc = torch.zeros(D, dtype=torch.float64)
                for (hh, ww, wt) in [(hl, wl, (1 - lh) * (1 - lw)), (hl, wl + 1, (1 - lh) * lw), (hl + 1, wl, lh * (1 - lw)),
                                     (hl + 1, wl + 1, lh * lw)]:
                    if 0 <= hh < H and 0 <= ww < W:
                        acc += wt * value[0, hh * W + ww, m]
                out[0, q, m * D:(m + 1) * D] += aw[0, q, m, 0, p] * acc
    assert torch.allclose(out, ref, atol=1e-10)


def test_sst_oracle_is_shuffle_invariant_and_masks_partial_windows():
    torch.manual_seed(0)
    B, S, C = 2, 21, 32          # 21 is not a multiple of 6: partial windows in both partitions
    di = {0: {'max_tokens': 36, 'drop_range': (0, 100000)}}
    lay = ss.SSTInputLayerV2(di, (6, 6, 1), [S, S, 1], shuffle_voxels=True, pos_temperature=1000, pos_embed=C)
    ys, xs = torch.meshgrid(torch.arange(S), torch.arange(S), indexing='ij')
    coords = torch.cat([torch.stack([torch.full((S * S,), b), torch.zeros(S * S, dtype=torch.long), ys.flatten(), xs.flatten()], 1)
                        for b in range(B)]).int()
    feats = torch.randn(B * S * S, C)
    net = ss.SSTv2([C] * 4, [8] * 4, 1, [C] * 4, output_shape=[S, S], in_channel=C).eval()
    a = net(lay(feats, coords, B))[0]
    lay.shuffle_voxels = False
    info = lay(feats, coords, B)
    b = net(info)[0]
    assert (a - b).abs().max() < 1e-5
    # closed forms of SURVEY Appendix D for the window partition
    n = int(np.ceil(S / 6)) + 1
    bw0, ciw0 = ss.get_window_coors(coords.long(), [S, S, 1], (6, 6, 1), False)
    assert torch.equal(bw0, coords[:, 0] * (n * n * 2) + ((coords[:, 3] + 6) // 6) * (n * 2) + ((coords[:, 2] + 6) // 6) * 2)
    assert torch.equal(ciw0[:, 1], (coords[:, 2].long() + 6) % 6) and torch.equal(ciw0[:, 2], (coords[:, 3].long() + 6) % 6)
    assert info['key_mask_shift1'][0].any()            # partial windows -> padded key slots exist


def test_cam2bev_oracle_against_direct_formula():
    """img_fv_to_bev == summing, per pillar, the bilinear samples of its 12 (zero-padded) points in all 6 cameras
    (SURVEY Appendix D 'Camera projection'), computed here point by point in float64."""
    torch.manual_seed(0)
    cfg = dict(num_points_in_pillar=12, embed_dims=256, bev_size=180, num_views=6, region_shape=[(6, 6, 1), (6, 6, 1)],
               grid_size=[[180, 180, 1], [90, 90, 1]],
               region_drop_info=[{0: {'max_tokens': 36, 'drop_range': (0, 100000)}}] * 2, instance_num=200)
    enc = fo.ISFusionEncoder(**cfg)
    l2i, aug, laug = synth.make_calib(1)
    feat = synth.make_img_feats(1)[1]
    pillars = torch.zeros(3, 12, 5)
    pillars[0, :3, :3] = torch.tensor([[10.0, 2.0, -1.0], [10.2, 2.1, -0.5], [9.9, 1.8, 0.2]])
    pillars[1, :1, :3] = torch.tensor([[-20.0, 5.0, -1.5]])
    pillars[2, :2, :3] = torch.tensor([[3.0, -15.0, 0.0], [3.1, -15.2, 0.5]])
    pc = torch.tensor([[0, 0, 93, 106], [0, 0, 98, 56], [0, 0, 64, 95]], dtype=torch.int32)
    out = enc.img_fv_to_bev([feat], 1, pts_metas=dict(pillars=pillars, pillar_coors=pc), img_metas=[dict(input_shape=(256, 704))],
                            lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    H, W = feat.shape[2], feat.shape[3]
    for k in range(3):
        acc = torch.zeros(256, dtype=torch.float64)
        for p in pillars[k, :, :3].double():
            for cam in range(6):
                q = l2i[0, cam, :3, :3].double() @ p + l2i[0, cam, :3, 3].double()
                z = q[2].clamp(1e-5, 1e5)
                uvz = torch.stack([q[0] / z, q[1] / z, z])
                a = aug[0, cam, :3, :3].double() @ uvz + aug[0, cam, :3, 3].double()
                gx, gy = (a[0] / 704 - 0.5) * 2, (a[1] / 256 - 0.5) * 2
                ix, iy = ((gx + 1) * W - 1) / 2, ((gy + 1) * H - 1) / 2
                x0, y0 = int(np.floor(float(ix))), int(np.floor(float(iy)))
                for (xx, yy, wt) in [(x0, y0, (1 - (ix - x0)) * (1 - (iy - y0))), (x0 + 1, y0, (ix - x0) * (1 - (iy - y0))),
                                     (x0, y0 + 1, (1 - (ix - x0)) * (iy - y0)), (x0 + 1, y0 + 1, (ix - x0) * (iy - y0))]:
                    if 0 <= xx < W and 0 <= yy < H:
                        acc += wt * feat[cam, :, yy, xx].double()
        got = out[0, :, pc[k, 2], pc[k, 3]].double()
        assert torch.allclose(got, acc, atol=1e-4), k
    assert out.abs().sum() > 0 and (out != 0).any(1).sum() == 3
