#!/usr/bin/env python
"""bench.py — frames/s of the IS-Fusion fusion-neck forward (0.075 m voxel, 6 cameras) on N B200s.

A "step" = one pass of the hot path (dynamic voxelize -> DynamicVFE -> SparseEncoder -> pillarize ->
ISFusionEncoder incl. SECONDV2 stage1/2; BASELINE.md "timed region") over one batch of synthetic frames.
Workload at N=1: BASELINE.json configs[1] (single B200, ~300k-point nuScenes-shaped frame, 6 cameras, fp16).
  value : frames/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the public module call with pinned HOST inputs (H2D in the timed region) and a D2H
          read of the two output feature maps
  --impl reference : the CPU oracle (restated reference, the reference itself cannot run without mmcv/spconv and
          has GPU-only ops) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'frames/sec fusion-neck forward (0.075 voxel, 6-cam)'


def frame_seed(rank, f):
    """frames are the sharded unit: rank r owns seeds 1000 r + f, disjoint across ranks (weak scaling)."""
    return 1000 * rank + f


def max_over_ranks(ms, world, dist=None, device=None):
    """whole-job time = slowest rank (device time of each rank, MAX all-reduce)."""
    if world <= 1:
        return ms
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_fps(world, batch, steps, ms):
    return world * batch * steps / (ms / 1000.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--points', type=int, default=300000)
    ap.add_argument('--cloud', default='S', choices=['S', 'U'])
    ap.add_argument('--img', default='384x1056', help='network input size of the 6 cameras (config img_scale; 256x704 = BASELINE configs[0])')
    ap.add_argument('--batch', type=int, default=1, help='frames per step per GPU')
    ap.add_argument('--frames', type=int, default=16, help='distinct resident frames rotated through (L2 defeat)')
    ap.add_argument('--cpu-points', type=int, default=0, help='points of the CPU-baseline sample frame (0 = same as --points)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--inflight', type=int, default=3, help='frames in flight (graph lanes replayed round-robin on their own streams)')
    ap.add_argument('--eager-only', action='store_true', help='profiling aid: time the eager exact-mode path only (no graph)')
    ap.add_argument('--stages', action='store_true', help='extra pass with per-stage CUDA-event timers')
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(',')]
                if len(f) >= 8:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(int(float(r[1])) for r in self.rows)
        reasons = []
        for i, name in enumerate(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']):
            if any('Active' == r[4 + i] for r in self.rows):
                reasons.append(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': int(float(self.rows[0][2])), 'samples': len(sm),
                'reasons': reasons}


def log(msg):
    sys.stderr.write('[bench %.1fs] %s\n' % (time.time() - T0, msg))
    sys.stderr.flush()
    try:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'bench_progress.log'), 'a') as f:
            f.write('[%.1fs] %s\n' % (time.time() - T0, msg))
    except Exception:
        pass


T0 = time.time()


def host_threads():
    """threads the CPU arm may use: the cores this process is allowed on (cgroup/affinity aware), capped at 64
    (the oracle's ops are small; more threads only add OpenMP overhead)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:   # cgroup v2 cpu.max quota
        q, p = open('/sys/fs/cgroup/cpu.max').read().split()
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(p))))
    except Exception:
        pass
    return max(1, min(n, 64))


def img_hw_of(args):
    h, w = args.img.lower().split('x')
    return (int(h), int(w))


def oracle_fps(points, cloud, steps, warmup, threads=None, img_hw=(384, 1056)):
    """CPU arm: the oracle neck forward on the host cores."""
    import torch
    import synth
    from oracle import model as omodel
    threads = threads or host_threads()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = synth.randomize_bn(omodel.ISFusionNeck(shuffle_voxels=True)).eval()
    l2i, aug, laug = synth.make_calib(1, img_hw)
    feats = synth.make_img_feats(1, img_hw)
    metas = [dict(input_shape=img_hw)]
    times, stages = [], {}
    for i in range(warmup + steps):
        pts = [synth.make_points(points if i >= warmup else 5000, cloud, i)]      # warm-up on a small frame
        tm = {}
        t0 = time.perf_counter()
        net(pts, feats, metas, timings=tm, lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
        dt = time.perf_counter() - t0
        log('cpu oracle frame %d (%d pts, %d threads): %.1f s' % (i, pts[0].shape[0], threads, dt))
        if i >= warmup:
            times.append(dt)
            for k, v in tm.items():
                stages[k] = stages.get(k, 0.0) + v / steps
    return steps / sum(times), threads, {k: round(v, 3) for k, v in stages.items()}


def workload_config(args, points):
    """`config` of the JSON line, identical for the B200 arm and the reference arm (how each arm runs it goes under `run`)."""
    return {'workload': 'BASELINE configs[1]: isfusion_0075voxel.py fusion-neck forward, synthetic nuScenes-shaped frame',
            'points_per_frame': points, 'cloud': args.cloud, 'cams': 6, 'img': list(img_hw_of(args)), 'bev': 180, 'batch_per_gpu': args.batch}


def cpu_baseline_subprocess(args, deadline=200):
    """the CPU arm in a child process with a hard deadline (a mis-sized OpenMP pool must not stall the bench)."""
    for pts_n in [args.cpu_points or args.points, 20000]:
        cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1', '--warmup', '1',
               '--points', str(args.points), '--cpu-points', str(pts_n), '--cloud', args.cloud, '--img', args.img]
        try:
            env = dict(os.environ)
            for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
                env.pop(k, None)
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=deadline, env=env).stdout.strip().splitlines()
            line = json.loads(out[-1])
            return line['cpu_baseline']
        except Exception as e:
            log('cpu baseline with %d points failed or timed out (%s); falling back' % (pts_n, type(e).__name__))
    return {'value': None, 'unit': 'frames/s', 'cores': host_threads(), 'kind': 'port', 'sample': 'timed out'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))
    warm = 1 if args.warmup > 0 else 0
    cpu_points = args.cpu_points or args.points
    fps, threads, stages = oracle_fps(cpu_points, args.cloud, steps, warm, img_hw=img_hw_of(args))
    sample = ('%d frame(s) of %d %s-cloud points + 6 cameras through the CPU oracle (1 warm-up on a 5k-point frame); '
              '%s (B200 arm: %d points)' % (steps, cpu_points, args.cloud,
                                            'same workload' if cpu_points == args.points else 'smaller bounded sample', args.points))
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warm, 'ms_per_step': 1000.0 / fps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, cpu_points),
            'run': {'parallelism': 'host cores of the box (CPU oracle = restated reference, fp32)'},
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port', 'sample': sample,
                             'stage_seconds': stages},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import synth
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from __graft_entry__ import make_net
    from isfusion_b200 import _lib as L
    import isfusion_b200.spconv as spmod
    net = make_net(0)              # product modules only: the oracle is not touched by the B200 arm
    net = net.to(dev).eval()
    B, F_ = args.batch, args.frames
    IMG = img_hw_of(args)
    metas = [dict(input_shape=IMG)] * B
    # resident inputs: F distinct frames per rank, rotated so that consecutive steps never reuse L2-resident inputs
    host_sets = []
    for f in range(F_):
        seed = frame_seed(rank, f)
        pts = [synth.make_points(args.points, args.cloud, seed * 8 + b).pin_memory() for b in range(B)]
        feats = [t.pin_memory() for t in synth.make_img_feats(B, IMG, seed=seed)]
        host_sets.append((pts, feats))
    l2i, aug, laug = synth.make_calib(B, IMG)
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    dev_sets = [([p.to(dev) for p in pts], [t.to(dev) for t in feats]) for pts, feats in host_sets]
    in_bytes = sum(p.numel() * 4 for p in host_sets[0][0]) + host_sets[0][1][1].numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps, warmup, base=0, join=None):
        for i in range(warmup):
            fn(base + i)
        torch.cuda.synchronize()
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(base + warmup + i)
        if join is not None:
            join()                  # the timing stream waits for every frame still in flight on the lane streams
        e1.record()
        torch.cuda.synchronize()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), world, dist, dev)

    def step_eager_pre(i):
        pts, feats = dev_sets[i % F_]
        return net(pts, feats, metas, **kw)

    if args.eager_only:
        ms = timed(step_eager_pre, args.steps, args.warmup)
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'mode': 'eager-only (profiling aid, not a bench value)',
                              'ms_per_step': ms / args.steps}))
        return
    from isfusion_b200.detector import FramePipeline, GraphedNeck
    LANES = max(1, args.inflight)
    log('calibrating buffer capacities on 2 frames (exact mode)')
    caps = GraphedNeck.calibrate(net, [dev_sets[0][0], dev_sets[1 % F_][0]], dev_sets[0][1], metas, margin=1.35, **kw)
    npts_cap = max(p.shape[0] for pts, _ in host_sets for p in pts)
    t_pre = timed(step_eager_pre, 5, 2) / 5
    log('eager before capture: %.3f ms/step' % t_pre)

    def build_pipeline(caps):
        return FramePipeline(net, LANES, B, npts_cap, img_hw=IMG, calib=kw, capacities=caps,
                             device=dev).capture(dev_sets[0][0], dev_sets[0][1][1])

    pipe = build_pipeline(caps)
    log('graph captured (%d lane(s)); strided-conv capacities %s' % (LANES, caps))

    def step_eager(i):
        pts, feats = dev_sets[i % F_]
        return net(pts, feats, metas, **kw)

    def step_resident(i):
        pts, feats = dev_sets[i % F_]
        return pipe.submit(pts, feats[1])   # D2D copy into the lane's static inputs + one graph replay

    out_host = [None] * LANES
    tickets = [None] * LANES

    def step_e2e(i):
        """public pipelined API: pinned host frame i -> H2D -> graph replay -> D2H of both output maps.  The H2D of
        the lane's next frame and the D2H of its previous frame run on a copy stream and overlap the replays; every
        byte still moves inside the timed region (it ends with a join + device synchronise)."""
        k = i % LANES
        if out_host[k] is None:
            o, _ = pipe.lanes[k].outputs
            out_host[k] = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in o]
        if tickets[k] is None:
            pts, feats = host_sets[i % F_]
            tickets[k] = pipe.prefetch(k, pts, feats[1])
        ticket = tickets[k]
        pts, feats = host_sets[(i + LANES) % F_]
        tickets[k] = pipe.prefetch(k, pts, feats[1])          # the lane's next frame: its H2D overlaps this replay
        return pipe.submit_prefetched(k, ticket, out_host[k])

    sampler = ClockSampler(local)
    if rank == 0:           # one sampler per job (rank 0's GPU): N concurrent nvidia-smi loops would perturb the run
        sampler.start()
    ms_eager = timed(step_eager, max(3, args.steps // 4), 2)
    ms_eager /= max(3, args.steps // 4)
    log('eager (exact mode, host syncs, after capture): %.3f ms/step' % ms_eager)
    L.launch_count(reset=True)
    for attempt in range(3):
        ms = timed(step_resident, args.steps, args.warmup, join=pipe.join)
        over = pipe.check()
        log('graph replay: %.3f ms/step  overflow=%s' % (ms / args.steps, over))
        # the decision is collective: every rank re-captures if any rank overflowed (keeps the barriers paired)
        if max_over_ranks(1.0 if over else 0.0, world, dist, dev) == 0.0:
            break
        # a frame produced more strided-conv output sites than calibrated: grow the buffers and re-capture
        caps = [int(c * 1.5) // 128 * 128 + 128 for c in caps]
        pipe = build_pipeline(caps)
        log('re-captured with capacities %s' % caps)
    else:
        raise SystemExit('bench.py: strided-conv capacity overflow persists: %s' % over)
    sampler.stop_flag = True
    # one frame alone (a single lane, nothing else in flight): the latency of a frame and the denominator of the
    # dominant kernel's share
    n_lat = max(5, args.steps // 4)
    ms_frame = timed(lambda i: pipe.lanes[0](dev_sets[i % F_][0], dev_sets[i % F_][1][1]), n_lat, 2) / n_lat
    log('single frame in flight: %.3f ms/frame' % ms_frame)
    L.launch_count(reset=True)
    prof = []
    spmod.PROFILE = prof
    step_eager(0)
    spmod.PROFILE = None
    own_per_step = L.launch_count(reset=True)      # kernels of one frame (a graph replay launches the same set)
    fps = job_fps(world, B, args.steps, ms)
    ms_e2e = timed(step_e2e, args.steps, max(3, args.warmup) // LANES * LANES + LANES, base=7 * LANES, join=pipe.join)
    fps_e2e = job_fps(world, B, args.steps, ms_e2e)
    log('e2e: %.3f ms/step' % (ms_e2e / args.steps))
    out_bytes = sum(h.numel() * h.element_size() for h in out_host[0])

    stage_ms = None
    if args.stages:
        from isfusion_b200 import profiling
        profiling.ENABLED = True
        profiling.reset()
        for i in range(args.steps):
            step_eager(100 + i)
        stage_ms = {k: round(v / args.steps, 4) for k, v in profiling.summary().items()}
        profiling.ENABLED = False
    roof = None
    if prof:
        # dominant kernel: the tcgen05 implicit GEMM of the sparse encoder; every launch of one frame is re-issued
        # back to back between CUDA events on its launching stream (same tensors, same shapes)
        try:
            recs = spmod.time_launches(prof, reps=5)
        except Exception as e:       # the headline numbers are already measured: never lose the line to the roofline pass
            log('roofline pass failed: %r' % (e,))
            recs = None
        if recs:
            t_ms = sum(r[0] for r in recs)
            issued = sum(r[1] for r in recs)
            useful = sum(r[2] for r in recs)
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
            except Exception:
                pass
            # the launches are re-issued in isolation (not inside a long step): the recipe's denominator is the BURST peak
            peak = peaks.get('bf16_tflops', 1600.0)
            top = max(recs, key=lambda r: r[0])
            # DRAM traffic of the largest launch: not measurable inside this run (needs ncu); the figure of the committed
            # `ncu --set full` capture of the same launch is quoted with its provenance, null when there is none
            traffic, traffic_src = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
                traffic, traffic_src = float(tj['dram_bytes_per_launch']), tj['source']
            except Exception:
                pass
            roof = {'kernel': 'tc::gemm5_kernel (CTA pairs, >= 128 columns) / tc::gemm4_kernel (tcgen05 sparse-conv implicit GEMM; %d launches per frame)' % len(recs),
                    'bound': 'tensor', 'achieved': useful / (t_ms * 1e-3) / 1e12, 'achieved_issued': issued / (t_ms * 1e-3) / 1e12,
                    'peak': peak, 'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst; launches timed in isolation)' if peaks else 'fallback 1.6 PF/s burst',
                    'frac_of_sustained': useful / (t_ms * 1e-3) / 1e12 / peaks.get('bf16_tflops_sustained', 1400.0),
                    'unit': 'TFLOP/s', 'frac': useful / (t_ms * 1e-3) / 1e12 / peak, 'traffic': traffic,
                    'traffic_note': ('dram read+write bytes of the largest launch, quoted from %s (ncu capture, not measured in this run)'
                                     % traffic_src) if traffic_src else 'no ncu capture of this build committed', 'launches_per_frame': len(recs),
                    'kernel_ms_per_step': t_ms, 'share_of_step': t_ms / ms_frame,
                    'share_note': 'kernel time of one frame / latency of one frame alone (%.3f ms); with %d frames in flight the '
                                  'steps overlap, so ms_per_step is below that latency' % (ms_frame, LANES),
                    'largest_launch': {'cin': top[3], 'cout': top[4], 'rows': top[5], 'ms': top[0],
                                       'useful_tflops': top[2] / (top[0] * 1e-3) / 1e12, 'issued_tflops': top[1] / (top[0] * 1e-3) / 1e12},
                    'note': 'achieved = sum(2*pairs*Cin*Cout) / sum(CUDA-event time per launch) over the 21 sparse-conv launches of '
                            'one frame, each re-issued 5x back to back; achieved_issued = the 128x64xCout MMA blocks the kernel '
                            'really walks (chunks skipped per tile excluded, zero-filled no-neighbour slots included)'}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(args)
    if rank == 0:
        line = {'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'frame_latency_ms': ms_frame, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16',
                'data': 'synthetic',
                'config': workload_config(args, args.points),
                'run': {'frames_in_flight': LANES, 'parallelism': 'dp%d (independent replicas, no data-path collective)' % world,
                        'l2': 'rotating %d distinct resident frames (%.0f MB of inputs) > 126 MB L2' % (F_, F_ * in_bytes / 1e6),
                        'value_note': '`value` = throughput with %d frames in flight (graph lanes); one frame alone takes '
                                      'frame_latency_ms' % LANES},
                'clocks': sampler.summary(),
                'e2e': {'value': fps_e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': in_bytes, 'd2h_bytes_per_step': out_bytes,
                        'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': int(round(own_per_step * args.steps)), 'eager_ms_per_step': ms_eager,
                'gpu_launches_note': 'kernels of libisfusion_b200.so (incl. its CUB passes) in the timed region; the dense 3x3 '
                                     'convs are among them (isf_conv3x3_nhwc_f16)',
                'roofline': roof, 'cpu_baseline': cpu}
        if stage_ms:
            line['stage_ms'] = stage_ms
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
