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
