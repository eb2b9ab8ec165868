"""Per-layer timing of the sparse-conv implicit GEMMs of one frame (round-2 iteration aid, GPU only):
   python scripts/bench_gemm_layers.py [--points 300000] [--v2]
prints, for each of the 21 launches, rows / Cin / Cout / ms and useful + issued TFLOP/s, for the default kernel and
(with --v2) the opt-in warp-specialised schedule side by side."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import synth  # noqa: E402
from __graft_entry__ import make_pair  # noqa: E402
from isfusion_b200 import _lib as L  # noqa: E402
import isfusion_b200.spconv as spmod  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=300000)
    ap.add_argument('--cloud', default='S')
    ap.add_argument('--v2', action='store_true')
    ap.add_argument('--reps', type=int, default=10)
    args = ap.parse_args()
    _, net = make_pair(0)
    net = net.cuda().eval()
    pts = [synth.make_points(args.points, args.cloud, 8).cuda()]
    l2i, aug, laug = synth.make_calib(1)
    feats = [f.cuda() for f in synth.make_img_feats(1)]
    metas = [dict(input_shape=(256, 704))]
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    net(pts, feats, metas, **kw)
    prof = []
    spmod.PROFILE = prof
    net(pts, feats, metas, **kw)
    spmod.PROFILE = None
    cols = {}
    for name, opt in (('v1', 0),) + ((('v2', 1),) if args.v2 else ()):
        L.set_option('gemm_v2', opt)
        cols[name] = spmod.time_launches(prof, reps=args.reps)
    L.set_option('gemm_v2', 0)
    print('%3s %8s %4s %4s | %s' % ('#', 'rows', 'cin', 'cout', ' | '.join('%s ms  useful/issued TF/s' % n for n in cols)))
    tot = {n: 0.0 for n in cols}
    for i in range(len(prof)):
        r0 = cols['v1'][i]
        cells = []
        for n, recs in cols.items():
            ms, issued, useful = recs[i][0], recs[i][1], recs[i][2]
            tot[n] += ms
            cells.append('%7.3f  %6.0f /%6.0f' % (ms, useful / ms / 1e9, issued / ms / 1e9))
        print('%3d %8d %4d %4d | %s' % (i, r0[5], r0[3], r0[4], ' | '.join(cells)))
    print('total ms: ' + '  '.join('%s %.3f' % (n, t) for n, t in tot.items()))


if __name__ == '__main__':
    main()
