#!/usr/bin/env python
"""BASELINE configs[4] on one GPU (round-2 bring-up tool): dense-scene stress — 1 M points / frame, 0.05 m voxels
(+-51.2 m -> 2048 x 2048 x 40 grid), 256 x 256 BEV — through the fusion neck in eager mode, with per-stage CUDA-event times
and the achieved HBM GB/s of the HBM-bound front end (dynamic voxelization + DynamicVFE: 36 B/point + 144 B/voxel, SURVEY
§8d) and the useful / issued TFLOP/s of the sparse-conv GEMM launches.  bench.py stays the contract benchmark (configs[1]).
    python scripts/bench_config5.py [--points 1000000 --steps 10 --warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_config5.py
Under torchrun every rank runs its own frames on its own GPU (frames are independent: replicas, no data-path collective);
the step time is the MAX over ranks (NCCL all-reduce of the CUDA-event time), `value` the whole-job frames/s, and the
roofline block lists, per HBM-bound stage, algorithmic bytes / CUDA-event time against the measured HBM peak.
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
    ap.add_argument('--points', type=int, default=1000000)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    args = ap.parse_args()
    import synth
    import isfusion_b200.detector as det
    import isfusion_b200.spconv as spmod
    from isfusion_b200 import profiling
    torch.manual_seed(0)
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    dev = torch.device(os.environ.get('ISF_TOOL_DEVICE', 'cuda:%d' % int(os.environ.get('LOCAL_RANK', '0'))))   # 'cpu': emulator dry run
    pc_range = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    cfg = synth.neck_config(voxel_size=[0.05, 0.05, 0.2], pc_range=pc_range, out_size_factor=8)
    net = synth.randomize_bn(det.ISFusionDetector(
        pts_voxel_layer=dict(point_cloud_range=pc_range, max_num_points=-1, voxel_size=cfg['voxel_size'], max_voxels=(-1, -1)),
        pts_voxel_encoder=dict(type='DynamicVFE', **cfg['pts_voxel_encoder']),
        pts_middle_encoder=dict(type='SparseEncoder', **cfg['pts_middle_encoder']),
        fusion_encoder=dict(type='ISFusionEncoder', **cfg['fusion_encoder']),
        pts_backbone=dict(type='SECONDV2', **cfg['pts_backbone']),
        detach=True, pc_range=pc_range, voxel_size=cfg['voxel_size'], out_size_factor=8)).to(dev).eval()
    frames = [synth.make_points(args.points, 'S', rank * 16 + s, pc_range).to(dev) for s in range(4)]
    l2i, aug, laug = (t.to(dev) for t in synth.make_calib(1, img_hw=(384, 1056)))
    feats = [f.to(dev) for f in synth.make_img_feats(1, img_hw=(384, 1056))]
    metas = [dict(input_shape=(384, 1056))]
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)

    def step(i):
        return net([frames[i % len(frames)]], feats, metas, **kw)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    profiling.ENABLED = True
    profiling.reset()
    prof = []
    spmod.PROFILE = prof
    step(0)
    spmod.PROFILE = None
    stages = {k: round(v, 4) for k, v in profiling.summary().items()}
    profiling.ENABLED = False
    line = {'metric': 'frames/sec fusion-neck forward, BASELINE configs[4] geometry (eager)', 'value': world * 1000.0 / ms, 'unit': 'frames/s', 'n_gpus': world, 'scaling': 'weak',
            'ms_per_step': ms, 'steps': args.steps, 'warmup': args.warmup, 'dtype': 'f16', 'data': 'synthetic',
            'config': {'workload': 'configs[4]: 0.05 m voxels, 2048^2 x 40 grid, 256x256 BEV', 'points_per_frame': args.points, 'cams': 6,
                       'img': [384, 1056], 'batch_per_gpu': 1, 'mode': 'eager, exact sizes (host syncs included)'},
            'stage_ms': stages, 'feat_shapes': [list(o.shape) for o in out[0]]}
    try:
        recs = spmod.time_launches(prof, reps=3)
        t = sum(r[0] for r in recs)
        line['sparse_conv'] = {'launches': len(recs), 'ms': t, 'rows_first_layer': recs[0][5],
                               'useful_tflops': sum(r[2] for r in recs) / (t * 1e-3) / 1e12,
                               'issued_tflops': sum(r[1] for r in recs) / (t * 1e-3) / 1e12}
        v = recs[0][5]
        fe_ms = sum(val for k, val in stages.items() if 'vox' in k.lower() or 'vfe' in k.lower())
        if fe_ms > 0:
            nbytes = 36.0 * args.points + 144.0 * v
            line['front_end'] = {'stages': [k for k in stages if 'vox' in k.lower() or 'vfe' in k.lower()], 'ms': fe_ms, 'voxels': v,
                                 'algorithmic_bytes': nbytes, 'gbps': nbytes / fe_ms / 1e6}
    except Exception as e:
        line['sparse_conv'] = 'failed: %r' % (e,)
    try:
        line['hbm_peak_gbps'] = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs')
    except Exception:
        line['hbm_peak_gbps'] = 6650.0       # fallback of the profiling recipe
    # HBM-bound stages: algorithmic bytes (rank 0's frame) / stage time -> GB/s and fraction of the measured peak
    try:
        g = cfg['fusion_encoder'].get('bev_size', 256)
        v = recs[0][5]
        pillars = g * g          # upper bound: every BEV cell holds a pillar on the dense synthetic cloud
        alg = {'a1 dynamic_voxelize': 32.0 * args.points, 'a3 dynamic_vfe': 36.0 * args.points + 144.0 * v,
               'a5 pillarize': 20.0 * args.points + pillars * 12 * 20.0,
               'a6 cam2bev': pillars * 12 * 12.0 + 6 * 256 * 24 * 66 * 2.0 + 256.0 * pillars * 2}
        roof = {}
        for k, nb in alg.items():
            if stages.get(k):
                gbps = nb / stages[k] / 1e6
                roof[k] = {'ms': stages[k], 'algorithmic_MB': round(nb / 1e6, 2), 'GBps': round(gbps, 1),
                           'frac_of_hbm_peak': round(gbps / line['hbm_peak_gbps'], 4)}
        line['roofline'] = {'bound': 'hbm', 'peak': line['hbm_peak_gbps'], 'unit': 'GB/s', 'stages': roof,
                            'note': 'eager stage times include launch gaps; the sparse encoder (tensor-bound) is under sparse_conv'}
    except Exception as e:
        line['roofline'] = 'failed: %r' % (e,)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
