#!/usr/bin/env python
"""Whole-config forward (BASELINE configs[1] read literally: "full isfusion_0075voxel.py forward"): raw camera images +
points -> SwinTransformer -> GeneralizedLSSFPN -> fusion neck -> SECONDFPN -> TransFusionHeadV2 -> decoded boxes, through
`ISFusionDetector.simple_test` (eager; the neck's CUDA-graph / frames-in-flight machinery of bench.py is not used here).
Inputs start in pinned host memory; the timed region includes their H2D copy and the D2H read of the boxes.  Reports the
end-to-end frames/s and the split image branch / neck / FPN + head from CUDA events.  Round-2 bring-up tool — bench.py
stays the contract benchmark of the neck.
    python scripts/bench_full_forward.py [--steps 20 --warmup 3 --points 300000 --img 384x1056]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--points', type=int, default=300000)
    ap.add_argument('--img', default='384x1056')
    ap.add_argument('--frames', type=int, default=4, help='distinct frames rotated through')
    args = ap.parse_args()
    import synth
    from __graft_entry__ import make_full_net
    from isfusion_b200 import _lib as L
    hw = tuple(int(v) for v in args.img.split('x'))
    dev = torch.device(os.environ.get('ISF_TOOL_DEVICE', 'cuda:0'))      # 'cpu' only for the emulator dry run in tests/
    if dev.type == 'cuda':
        torch.cuda.set_device(dev)
    net = make_full_net(0).to(dev).eval()
    frames = []
    for i in range(args.frames):
        pts = synth.make_points(args.points, 'S', i).pin_memory()
        img = torch.randn(1, 6, 3, hw[0], hw[1], generator=torch.Generator().manual_seed(100 + i)).half().pin_memory()
        frames.append((pts, img))
    l2i, aug, laug = (t.to(dev) for t in synth.make_calib(1, img_hw=hw))
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)

    def step(i, ev=None):
        pts, img = frames[i % len(frames)]
        p = pts.to(dev, non_blocking=True)
        im = img.to(dev, non_blocking=True)
        metas = [dict()]
        if ev:
            ev[0].record()
        feats = net.extract_img_feat(im, metas)
        if ev:
            ev[1].record()
        x, _ = net.extract_pts_feat([p], feats, metas, **kw)
        if ev:
            ev[2].record()
        res = net.simple_test_pts(x, feats, metas)
        boxes = res[0]['boxes_3d'].float().cpu()          # D2H read of the result
        if ev:
            ev[3].record()
        return boxes

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        boxes = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (L.launch_count() - n0) // args.steps
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    step(0, ev)
    torch.cuda.synchronize()
    pts, img = frames[0]
    print(json.dumps({
        'metric': 'frames/sec full isfusion_0075voxel.py forward (images + points -> boxes)', 'value': 1000.0 / ms, 'unit': 'frames/s',
        'ms_per_step': ms, 'steps': args.steps, 'warmup': args.warmup, 'dtype': 'f16', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1], whole config: Swin-T + GeneralizedLSSFPN + fusion neck + SECONDFPN + TransFusionHeadV2',
                   'points_per_frame': args.points, 'cams': 6, 'img': list(hw), 'batch_per_gpu': 1, 'mode': 'eager simple_test'},
        'split_ms': {'image_branch': ev[0].elapsed_time(ev[1]), 'neck': ev[1].elapsed_time(ev[2]), 'fpn_head_decode_d2h': ev[2].elapsed_time(ev[3])},
        'h2d_bytes_per_step': pts.numel() * 4 + img.numel() * 2, 'd2h_bytes_per_step': int(boxes.numel()) * 4,
        'own_launches_per_step': int(launches), 'boxes': list(boxes.shape)}))


if __name__ == '__main__':
    main()
