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
    ap.add_argument('--raw', action='store_true', help='start from RAW sensor data: 6 uint8 camera frames (1600x900, or --raw-hw) + key frame and 10 raw '
                    'sweeps in pinned host memory; resize / crop / normalise and the sweep merge run on the device (input contract kernels)')
    ap.add_argument('--raw-hw', default='900x1600', help='camera frame size for --raw')
    args = ap.parse_args()
    import synth
    from __graft_entry__ import make_full_net
    from isfusion_b200 import _lib as L
    hw = tuple(int(v) for v in args.img.split('x'))
    dev = torch.device(os.environ.get('ISF_TOOL_DEVICE', 'cuda:0'))      # 'cpu' only for the emulator dry run in tests/
    if dev.type == 'cuda':
        torch.cuda.set_device(dev)
    net = make_full_net(0).to(dev).eval()
    l2i, aug, laug = (t.to(dev) for t in synth.make_calib(1, img_hw=hw))
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    frames = []
    if args.raw:
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
        import golden_common as gc
        from isfusion_b200.image_pipeline import ImageAug3D, image_aug_normalize
        from isfusion_b200.sweeps import LoadPointsFromMultiSweeps, merge_sweeps
        rh, rw = (int(v) for v in args.raw_hw.split('x'))
        scale = max(hw[0] / rh, hw[1] / rw) * 1.0001                 # the config's 0.72 for 900x1600 -> 384x1056 (up to rounding)
        scale = 0.72 if (rh, rw, hw) == (900, 1600, (384, 1056)) else scale
        augm = ImageAug3D(final_dim=hw, resize_lim=[scale, scale], bot_pct_lim=[0.0, 0.0], rot_lim=[0.0, 0.0], rand_flip=False, is_train=False)
        _, dims, crop, _, _ = augm.sample_augmentation(dict(ori_shape=(rw, rh)))
        loader = LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], device=str(dev))
        per_sweep = max(args.points // 11, 1)
        for i in range(args.frames):
            cloud = synth.make_points(args.points, 'S', i).numpy()            # LiDAR-shaped returns, cut into key frame + 10 sweeps
            _, ts, sweeps = gc.sweep_inputs(seed=20 + i, n_key=20, n_sweeps=10, n_sweep=20)      # ego motion / time stamps of the sweeps
            key = cloud[:args.points - 10 * per_sweep]
            for j, sw in enumerate(sweeps):
                lo = args.points - (10 - j) * per_sweep
                sw['points'] = cloud[lo:lo + per_sweep]
            blocks, seg_start, flags, lag, rot, trans = loader.plan(dict(points=key, timestamp=ts, sweeps=sweeps))
            raw_pts = torch.from_numpy(np.concatenate(blocks, 0)).pin_memory()
            cams = torch.from_numpy(np.random.RandomState(i).randint(0, 256, (6, rh, rw, 3)).astype(np.uint8)).pin_memory()
            frames.append((raw_pts, cams, (seg_start, flags, lag, rot, trans)))
    else:
        for i in range(args.frames):
            pts = synth.make_points(args.points, 'S', i).pin_memory()
            img = torch.randn(1, 6, 3, hw[0], hw[1], generator=torch.Generator().manual_seed(100 + i)).half().pin_memory()
            frames.append((pts, img))

    def step(i, ev=None):
        if args.raw:
            raw_pts, cams, (seg_start, flags, lag, rot, trans) = frames[i % len(frames)]
            # the config's PointsRangeFilter rides in the merge kernel (ego motion moves some returns out of the range)
            p, cnt = merge_sweeps(raw_pts.to(dev, non_blocking=True), seg_start, flags, lag, rot, trans, point_cloud_range=synth.PC_RANGE)
            p = p[:int(cnt.item())]
            im = image_aug_normalize(cams.to(dev, non_blocking=True), dims, crop, False, gc.IMG_NORM['mean'], gc.IMG_NORM['std'],
                                     out_dtype=torch.float16).unsqueeze(0)
        else:
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
    pts, img = frames[0][0], frames[0][1]
    print(json.dumps({
        'metric': 'frames/sec full isfusion_0075voxel.py forward (images + points -> boxes)', 'value': 1000.0 / ms, 'unit': 'frames/s',
        'ms_per_step': ms, 'steps': args.steps, 'warmup': args.warmup, 'dtype': 'f16', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1], whole config: Swin-T + GeneralizedLSSFPN + fusion neck + SECONDFPN + TransFusionHeadV2',
                   'points_per_frame': args.points, 'cams': 6, 'img': list(hw), 'batch_per_gpu': 1, 'mode': 'eager simple_test'},
        'split_ms': {'image_branch': ev[0].elapsed_time(ev[1]), 'neck': ev[1].elapsed_time(ev[2]), 'fpn_head_decode_d2h': ev[2].elapsed_time(ev[3])},
        'h2d_bytes_per_step': pts.numel() * pts.element_size() + img.numel() * img.element_size(), 'raw_sensor_input': bool(args.raw), 'd2h_bytes_per_step': int(boxes.numel()) * 4,
        'own_launches_per_step': int(launches), 'boxes': list(boxes.shape)}))


if __name__ == '__main__':
    main()
