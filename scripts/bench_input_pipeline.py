#!/usr/bin/env python
"""Input-contract kernels at the config's real sizes (round-2 bring-up tool, GPU only): CUDA-event time and achieved GB/s of
  * isf_merge_sweeps          key frame + 10 sweeps, ~300 k points (20 B read + 20 B write per point; + 8 B with the close filter)
  * isf_image_aug_normalize   6 cameras 1600x900 uint8 -> resize 0.72 -> crop 384x1056 -> normalised fp16 / fp32
against the algorithmic bytes (inputs resident in HBM; 8 distinct inputs rotated).  NOTE: a call is a few microseconds of kernel
time behind ~50 us of Python / ctypes argument marshalling, so `ms` is per CALL (host-paced) and `gbps` a lower bound; the kernel
times proper come from the ncu launch list (`ncu -k regex:merge_|image_ ...`).  Plus
the same through the host-facing pipeline classes with pinned host buffers (H2D inside the timed region).
    python scripts/bench_input_pipeline.py [--reps 50]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def timed(fn, reps, warmup=5):
    for i in range(min(warmup, reps)):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=50)
    ap.add_argument('--small', action='store_true', help='tiny sizes (dry run of this script on the CPU emulator)')
    args = ap.parse_args()
    import golden_common as gc
    from isfusion_b200.image_pipeline import ImageAug3D, ImageNormalize, image_aug_normalize
    from isfusion_b200.sweeps import LoadPointsFromMultiSweeps, merge_sweeps
    dev = torch.device(os.environ.get('ISF_TOOL_DEVICE', 'cuda:0'))      # 'cpu' only for the emulator dry run in tests/
    H, W, NCAM = (90, 160, 2) if args.small else (900, 1600, 6)
    RS, CROP, FD = ((115, 64), (9, 16, 105, 64), (48, 96)) if args.small else ((1152, 648), (48, 264, 1104, 648), (384, 1056))
    out = {}
    # ---- sweeps
    key, ts, sweeps = gc.sweep_inputs(seed=3, n_key=300 if args.small else 30000, n_sweeps=10, n_sweep=270 if args.small else 27000)
    for rc in (False, True):
        loader = LoadPointsFromMultiSweeps(use_dim=[0, 1, 2, 3, 4], remove_close=rc, device=str(dev))
        blocks, seg_start, flags, lag, rot, trans = loader.plan(dict(points=key, timestamp=ts, sweeps=sweeps))
        raws = [torch.from_numpy(np.concatenate(blocks, 0) + np.float32(i)).to(dev) for i in range(8)]      # 8 distinct inputs
        m = raws[0].shape[0]
        ms = timed(lambda i: merge_sweeps(raws[i % 8], seg_start, flags, lag, rot, trans), args.reps)
        nbytes = m * (40 + (8 if rc else 0))
        ms_e2e = timed(lambda i: loader(dict(points=key, timestamp=ts, sweeps=sweeps)), max(args.reps // 5, 1 if args.small else 5))
        out['merge_sweeps' + ('_close' if rc else '')] = dict(points=m, ms=ms, gbps=nbytes / ms / 1e6, algorithmic_bytes=nbytes,
                                                               ms_host_interface=ms_e2e)
    # ---- cameras
    rs = np.random.RandomState(0)
    imgs = [torch.from_numpy(rs.randint(0, 256, (NCAM, H, W, 3)).astype(np.uint8)).to(dev) for _ in range(2 if args.small else 8)]
    for dt in (torch.float16, torch.float32):
        ms = timed(lambda i: image_aug_normalize(imgs[i % len(imgs)], RS, CROP, False, gc.IMG_NORM['mean'], gc.IMG_NORM['std'],
                                                 out_dtype=dt), args.reps)
        rows = H - int(CROP[1] * H / RS[1])                   # input rows below the crop line that the vertical pass reads (approx.)
        ow, oh = CROP[2] - CROP[0], CROP[3] - CROP[1]
        nbytes = NCAM * (rows * W * 3 + 2 * rows * ow * 3 + oh * ow * 3 * (2 if dt == torch.float16 else 4))
        out['image_aug_normalize_' + ('f16' if dt == torch.float16 else 'f32')] = dict(ms=ms, gbps=nbytes / ms / 1e6, algorithmic_bytes=nbytes)
    host = [rs.randint(0, 256, (H, W, 3)).astype(np.uint8) for _ in range(NCAM)]
    aug = ImageAug3D(final_dim=FD, resize_lim=[0.72, 0.72], bot_pct_lim=[0.0, 0.0], rot_lim=[0.0, 0.0], rand_flip=False, is_train=False)
    norm = ImageNormalize(out_dtype=torch.float16, device=str(dev), **gc.IMG_NORM)
    out['image_pipeline_host_interface_ms'] = timed(lambda i: norm(aug(dict(img=list(host), ori_shape=(W, H)))), max(args.reps // 5, 1 if args.small else 5))
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        out['hbm_peak_gbps'] = peaks.get('hbm_gbs')
    except Exception:
        out['hbm_peak_gbps'] = None
    print(json.dumps(out))


if __name__ == '__main__':
    main()
