#!/usr/bin/env python
"""Training step of the LiDAR branch (the trainable part of BASELINE configs[3]: "batch = 32 synthetic frames data-parallel across
8 x B200, NCCL grad allreduce, training step timing"), run under torchrun on N >= 1 GPUs:

    voxelize -> DynamicVFE (train: batch-statistics naiveSyncBN1d over the ranks, own scatter forward / backward)
             -> SparseEncoder (train: implicit-GEMM kernels under autograd, spconv.ImplicitGemmFunction) -> BEV map
    loss = <BEV, fixed random weights> (a stand-in: targets / assigner / losses of the head are not built)
    backward -> ONE flat NCCL all-reduce of the gradients (data parallel) -> SGD update of the fp32 master weights

Prints one JSON line: step time (CUDA events, max over ranks), the share of the gradient all-reduce, gradient bytes.  What is NOT in
the step: the camera branch, the fusion encoder, the head and its loss (no backward of those modules exists);
the weight gradient of the sparse convolutions is isf_spconv_wgrad (ISF_WGRAD=torch: gathered copies + library GEMMs).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train_lidar.py --frames 4"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=4, help='frames per GPU and step (configs[3]: 32 over 8 GPUs = 4)')
    ap.add_argument('--points', type=int, default=100000)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        dist.init_process_group('nccl')
    import synth
    from __graft_entry__ import make_net
    import isfusion_b200.spconv as spconv_mod
    net = make_net(0).cuda()
    vfe, enc = net.pts_voxel_encoder.train(), net.pts_middle_encoder.train()
    params = [p for m in (vfe, enc) for p in m.parameters()]
    nbytes = sum(p.numel() for p in params) * 4
    pts = [synth.make_points(args.points, 'S', 1000 * rank + i).cuda() for i in range(args.frames)]
    with torch.no_grad():
        voxels, coors = net.dynamic_voxelize(pts)              # [N, 5], [N, 4] (b, z, y, x); out-of-range points dropped here
        keep = (coors[:, 1:] >= 0).all(1)
        voxels, coors = voxels[keep].contiguous(), coors[keep].contiguous()
    torch.manual_seed(7 + rank)
    target = None
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_step, t_ar = [], []
    for it in range(args.warmup + args.steps):
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for p in params:
            p.grad = None
        vf, vc = vfe(voxels, coors)
        bev, _, _ = enc(vf, vc, args.frames, sorted_coors=True)
        if target is None:
            target = torch.randn_like(bev)
        loss = (bev * target).mean()
        loss.backward()
        e1.record()
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
        if world > 1:
            dist.all_reduce(flat)
            flat /= world
        e2.record()
        off = 0
        with torch.no_grad():
            for p in params:
                p.add_(flat[off:off + p.numel()].view_as(p), alpha=-1e-4)
                off += p.numel()
        e3.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            t_step.append(e0.elapsed_time(e3))
            t_ar.append(e1.elapsed_time(e2))
    ms = torch.tensor([sum(t_step) / len(t_step), sum(t_ar) / len(t_ar)], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({'metric': 'LiDAR-branch training step (DynamicVFE + SparseEncoder, data parallel)', 'n_gpus': world,
                          'frames_per_gpu': args.frames, 'global_batch': args.frames * world, 'points_per_frame': args.points,
                          'voxels_rank0': int(vc.shape[0]), 'ms_per_step_max_over_ranks': float(ms[0]), 'frames_per_s': args.frames * world / float(ms[0]) * 1e3,
                          'grad_allreduce_ms': float(ms[1]), 'grad_allreduce_share': float(ms[1] / ms[0]), 'gradient_bytes': nbytes,
                          'loss': float(loss.detach()), 'steps': args.steps, 'warmup': args.warmup,
                          'not_in_the_step': 'camera branch, fusion encoder, head + loss (no backward of those modules)',
                          'weight_gradient': spconv_mod.WGRAD + ' (ISF_WGRAD: own = csrc/wgrad.cu, torch = gathered copies + library GEMMs)'}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
