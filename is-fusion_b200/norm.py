"""naiveSyncBN1d / naiveSyncBN2d (ref: mmdet3d/ops/norm.py:8-24 AllReduce, :136-202 1-D, :205-271 2-D).

The only exchange step on the whole path, and only while TRAINING on more than one rank: each rank's per-channel
[mean, mean-of-squares] (2·C floats, 512 B for C = 64) is all-gathered and the rank values are averaged with EQUAL
weight per rank (not per row — the reference quirk that keeps ranks with different point counts stable); the
gradient of that average is an all-reduce.  Eval, a single process, or an uninitialised process group behave as
plain BatchNorm (`norm.py:173-175`), which is what the fused inference kernels fold into scale/shift.

One class for both ranks of tensor: the channel dimension is 1, every other dimension is reduced.
"""
import torch
import torch.distributed as dist
from torch import nn


class _RankAverage(torch.autograd.Function):
    """forward: mean over ranks of a small vector (all_gather + mean, as the reference does instead of an in-place
    all_reduce); backward: all_reduce(sum) of the incoming gradient, scaled by 1/world like the forward."""

    @staticmethod
    def forward(ctx, vec):
        world = dist.get_world_size()
        parts = [torch.empty_like(vec) for _ in range(world)]
        dist.all_gather(parts, vec.contiguous())
        ctx.world = world
        return torch.stack(parts).sum(0) / world

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous().clone()
        dist.all_reduce(grad)
        return grad / ctx.world


def _multi_rank():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class _RankSyncForward:
    """mixin in front of nn.BatchNorm1d / nn.BatchNorm2d (so isinstance checks and state_dict keys stay those of BN)."""

    def forward(self, x):
        self._check_input_dim(x)
        if not (self.training and _multi_rank()):
            return super().forward(x)
        if x.dtype != torch.float32:
            raise TypeError(f'input should be in float32 type, got {x.dtype}')       # norm.py:170-171
        if x.shape[0] == 0:
            raise ValueError('SyncBN does not support empty inputs')                  # norm.py:176
        c = x.shape[1]
        red = [d for d in range(x.dim()) if d != 1]
        moments = torch.cat([x.mean(dim=red), (x * x).mean(dim=red)])
        mean, meansqr = _RankAverage.apply(moments).split(c)
        var = meansqr - mean * mean
        with torch.no_grad():                                                         # norm.py:190-192
            self.running_mean += self.momentum * (mean - self.running_mean)
            self.running_var += self.momentum * (var - self.running_var)
        scale = self.weight * torch.rsqrt(var + self.eps)
        shift = self.bias - mean * scale
        bshape = [1, c] + [1] * (x.dim() - 2)
        return x * scale.view(bshape) + shift.view(bshape)


class NaiveSyncBatchNorm1d(_RankSyncForward, nn.BatchNorm1d):
    """(N, C) or (N, C, L) — DynamicVFE's norm (`config:70`)."""


class NaiveSyncBatchNorm2d(_RankSyncForward, nn.BatchNorm2d):
    """(N, C, H, W)."""
