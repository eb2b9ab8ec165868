"""world_size-2 gloo test (CPU) of the multi-GPU host logic of bench.py: frames are sharded across ranks with
disjoint seeds (no data-path collective), the job time is the MAX over ranks, value = all ranks' frames / that time."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import bench
    dist.init_process_group('gloo', rank=rank, world_size=world)
    seeds = [bench.frame_seed(rank, f) for f in range(4)]
    ms_local = 10.0 + 5.0 * rank                       # rank 1 is the slow one
    dist.barrier()
    ms = bench.max_over_ranks(ms_local, world, dist, None)
    gathered = [None] * world
    dist.all_gather_object(gathered, seeds)
    q.put((rank, ms, gathered, bench.job_fps(world, 1, 20, ms * 20)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    for rank, ms, gathered, fps in res:
        assert ms == 15.0                                                  # max over ranks on every rank
        assert len(set(gathered[0]) & set(gathered[1])) == 0               # disjoint frames per rank
        assert abs(fps - 2 * 1 * 20 / (15.0 * 20 / 1000.0)) < 1e-6         # whole-job frames / slowest rank's time


def test_frame_generators_are_rank_disjoint_and_seeded():
    sys.path.insert(0, ROOT)
    import bench
    import synth
    a = synth.make_points(2000, 'S', bench.frame_seed(0, 1) * 8)
    b = synth.make_points(2000, 'S', bench.frame_seed(1, 1) * 8)
    a2 = synth.make_points(2000, 'S', bench.frame_seed(0, 1) * 8)
    assert torch.equal(a, a2) and not torch.equal(a, b)


def _bn_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from isfusion_b200.norm import NaiveSyncBatchNorm1d, NaiveSyncBatchNorm2d
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    out = {}
    for name, cls, shape in (('1d', NaiveSyncBatchNorm1d, (50 + 30 * rank, 8)), ('2d', NaiveSyncBatchNorm2d, (2, 8, 5, 3 + rank))):
        torch.manual_seed(0)
        bn = cls(8, eps=1e-3, momentum=0.01)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.5, 0.5)
        x = (torch.randn(shape, generator=g) * (1 + rank) + rank).requires_grad_(True)
        bn.train()
        y = bn(x)
        y.square().sum().backward()
        bn.eval()
        y_eval = bn(x.detach())
        out[name] = {k: v.detach().clone().numpy() for k, v in dict(
            x=x, y=y, gx=x.grad, rm=bn.running_mean, rv=bn.running_var, w=bn.weight, b=bn.bias, y_eval=y_eval).items()}
    q.put((rank, out))
    dist.destroy_process_group()


def test_naive_sync_bn_two_ranks():
    """a15 (ref ops/norm.py:160-202): ranks are averaged with equal weight; checked against the closed form computed
    in one process from both ranks' inputs, gradients against autograd through that closed form."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bn_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    res = {r: {n: {k: torch.from_numpy(v) for k, v in d.items()} for n, d in o.items()} for r, o in res.items()}
    for name in ('1d', '2d'):
        xs = [res[r][name]['x'].clone().requires_grad_(True) for r in range(2)]
        w = res[0][name]['w'].clone().requires_grad_(True)
        b = res[0][name]['b']
        red = [d for d in range(xs[0].dim()) if d != 1]
        mean = sum(x.mean(dim=red) for x in xs) / 2
        meansqr = sum((x * x).mean(dim=red) for x in xs) / 2
        var = meansqr - mean * mean
        bshape = [1, -1] + [1] * (xs[0].dim() - 2)
        scale = w * torch.rsqrt(var + 1e-3)
        ys = [x * scale.view(bshape) + (b - mean * scale).view(bshape) for x in xs]
        # every rank back-propagates its own loss; DDP-style total = sum of the per-rank losses
        sum(y.square().sum() for y in ys).backward()
        for r in range(2):
            o = res[r][name]
            assert torch.allclose(o['y'], ys[r].detach(), atol=1e-5)
            assert torch.allclose(o['gx'], xs[r].grad, atol=1e-4, rtol=1e-4)
            assert torch.allclose(o['rm'], 0.01 * mean.detach(), atol=1e-6)
            assert torch.allclose(o['rv'], 1 + 0.01 * (var.detach() - 1), atol=1e-6)
            # eval = plain BatchNorm on the running statistics
            ref = torch.nn.functional.batch_norm(o['x'], o['rm'], o['rv'], o['w'], o['b'], False, 0.0, 1e-3)
            assert torch.allclose(o['y_eval'], ref, atol=1e-5)


def _sparse_train_worker(rank, world, port, q):
    """one data-parallel training step of the product SparseEncoder (train mode, naiveSyncBN1d over gloo) on the ABI emulator."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import pytest
    import abi_emulator
    import test_sparse_train_gpu as T
    from isfusion_b200.sparse_encoder import SparseEncoder
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mp_ = pytest.MonkeyPatch()
    abi_emulator.install(mp_)
    try:
        torch.manual_seed(0)                                         # same initial weights on every rank
        cfg = dict(in_channels=64, sparse_shape=[41, 64, 64], base_channels=32, output_channels=256,
                   norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01),
                   encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
                   encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')
        net = SparseEncoder(**cfg).train()
        feats, idx = T._inputs(2500 + 700 * rank, 64, [41, 64, 64], 2, 11 + rank)      # ranks see different clouds
        x = feats.clone().requires_grad_(True)
        out, _, _ = net(x, idx, 2)
        out.square().mean().backward()
        flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        own = flat.clone()
        dist.all_reduce(flat)                                        # the data-parallel gradient exchange
        flat /= world
        stats = torch.cat([b.float().reshape(-1) for n, b in net.named_buffers() if 'running' in n])
        # plain Python objects through the queue (a tensor would need this process alive until the parent has read it)
        q.put((rank, own.norm().item(), flat.double().sum().item(), flat[::997].tolist(), stats.tolist(), bool(torch.isfinite(flat).all())))
    finally:
        mp_.undo()
        dist.destroy_process_group()


def test_sparse_encoder_train_step_two_ranks():
    """N > 1 path of the training step on the CPU: the sparse encoder in train mode on both ranks (ABI emulator), its
    naiveSyncBN1d layers exchanging moments over gloo, one flat gradient all-reduce: the ranks end up with the same running
    statistics and the same averaged gradient although their inputs (and own gradients) differ."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 777) % 2000
    procs = [ctx.Process(target=_sparse_train_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    (r0, n0, t0, g0, s0, ok0), (r1, n1, t1, g1, s1, ok1) = res
    assert ok0 and ok1 and n0 > 0 and n1 > 0 and abs(n0 - n1) > 1e-9          # different data, different own gradients
    assert t0 == t1 and g0 == g1 and len(g0) > 1000                           # one averaged gradient, bit for bit
    assert s0 == s1 and len(s0) > 100                                         # synchronised BN: identical running statistics
