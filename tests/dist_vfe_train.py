#!/usr/bin/env python
"""a15 over NCCL (run under torchrun on >= 2 GPUs): one training step of the product DynamicVFE in TRAIN mode on every rank
(own scatter kernels forward + backward, naiveSyncBN1d exchanging its [mean, mean-of-squares] vector over NCCL), the gradient
all-reduce of a data-parallel step, and the same step of the CPU oracle over gloo in the same process group.  Asserts that the
outputs, the rank-averaged gradients and the running statistics agree; prints one JSON line with the step time.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_vfe_train.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class RefAllReduce(torch.autograd.Function):
    """the reference's `AllReduce` autograd function (mmdet3d/ops/norm.py:8-24): forward = all_gather + sum, backward = all_reduce."""

    @staticmethod
    def forward(ctx, x):
        parts = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, x.contiguous())
        return torch.stack(parts).sum(0)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g)
        return g


class RefSyncBN1d(torch.nn.BatchNorm1d):
    """restatement of the reference's naiveSyncBN1d.forward (mmdet3d/ops/norm.py:160-202) for the checker side: per-rank
    [mean, mean of squares], averaged over ranks with EQUAL weight (the reference's AllReduce function), biased variance."""

    def forward(self, x):
        c = x.shape[1]
        vec = torch.cat([x.mean(0), (x * x).mean(0)])
        vec = RefAllReduce.apply(vec) * (1.0 / dist.get_world_size())
        mean, meansqr = vec.split(c)
        var = meansqr - mean * mean
        with torch.no_grad():
            self.running_mean += self.momentum * (mean.detach() - self.running_mean)
            self.running_var += self.momentum * (var.detach() - self.running_var)
        scale = self.weight * torch.rsqrt(var + self.eps)
        return x * scale + (self.bias - mean * scale)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group('cpu:gloo,cuda:nccl')
    import synth
    from oracle import voxel_ops as vo            # checker side only (this is a test script)
    from isfusion_b200.voxel_encoder import DynamicVFE
    vs, pr = [0.075, 0.075, 0.2], synth.PC_RANGE
    torch.manual_seed(0)
    ref = synth.randomize_bn(vo.DynamicVFE(5, [64, 64], False, True, True, vs, pr, dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01))).train()
    mod = DynamicVFE(5, [64, 64], False, True, True, vs, pr, dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01))
    mod.load_state_dict(ref.state_dict())
    mod = mod.cuda().train()
    for layer in ref.vfe_layers:                   # the oracle's BN layers become the reference's rank-synchronised BN (gloo)
        bn = RefSyncBN1d(layer.norm.num_features, eps=layer.norm.eps, momentum=layer.norm.momentum)
        bn.load_state_dict(layer.norm.state_dict())
        layer.norm = bn.train()
    n = 40000 + 7000 * rank                        # ranks see different point counts (the equal-weight rank average matters)
    pts = synth.make_points(n, 'S', 100 + rank)
    c3 = vo.dynamic_voxelize(pts, vs, pr)
    keep = (c3 >= 0).all(1)
    pts, c3 = pts[keep], c3[keep]
    coors = torch.nn.functional.pad(c3, (1, 0), value=0)
    torch.manual_seed(5 + rank)
    rf, rc = vo.dynamic_vfe_train_forward(ref, pts, coors)
    wgt = torch.randn(rf.shape)
    (rf * wgt).sum().backward()
    for p in ref.parameters():
        dist.all_reduce(p.grad)                    # gloo
        p.grad /= world
    dpts, dcoors, dw = pts.cuda(), coors.cuda(), wgt.cuda()
    for it in range(3):                            # step 0 = checked, the rest = timing
        if it == 1:
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        mod.zero_grad(set_to_none=True)
        f, c = mod(dpts, dcoors)
        (f * dw).sum().backward()
        flat = torch.cat([p.grad.reshape(-1) for p in mod.parameters()])
        dist.all_reduce(flat)                      # NCCL: the data-parallel gradient exchange
        flat /= world
        if it == 0:
            torch.cuda.synchronize()
            assert torch.equal(c.cpu(), rc)
            assert torch.allclose(f.detach().cpu(), rf.detach(), rtol=1e-4, atol=1e-4), float((f.detach().cpu() - rf).abs().max())
            off = 0
            for (name, p), q in zip(mod.named_parameters(), ref.parameters()):
                g = flat[off:off + p.numel()].view_as(p).cpu()
                off += p.numel()
                r = ((g - q.grad).norm() / q.grad.norm().clamp(min=1e-12)).item()
                assert r < 1e-3, (name, r)
            for (name, b), q in zip(mod.named_buffers(), ref.buffers()):
                assert torch.allclose(b.cpu().float(), q.float(), rtol=1e-4, atol=1e-5), name
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 2], device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({'test': 'DynamicVFE train step over NCCL (a15 naiveSyncBN1d + gradient all-reduce)', 'world': world,
                          'points_rank0': int(pts.shape[0]), 'ms_per_step_max_over_ranks': float(ms.item()), 'parity': 'ok'}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
