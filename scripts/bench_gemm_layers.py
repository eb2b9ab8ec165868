"""Per-layer timing of the sparse-conv implicit GEMMs of one frame (round-2 iteration aid, GPU only):
   python scripts/bench_gemm_layers.py [--points 300000] [--v2] [--tile-order 12,16,31]
prints, for each of the 21 launches, rows / Cin / Cout / ms and useful TFLOP/s for the default kernel and, side by
side, the opt-in variants: the warp-specialised schedule (--v2, csrc/tc_gemm2.cuh) and the neighbour-pattern tile
order (--tile-order, csrc/spconv.cu isf_spconv_tile_order; chunk log2 per variant, 31 = one global sort), alone and
combined.  Also checks that every variant's encoder output is bit-identical to the default one."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import synth  # noqa: E402
from __graft_entry__ import make_net  # noqa: E402
from isfusion_b200 import _lib as L  # noqa: E402
import isfusion_b200.spconv as spmod  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=300000)
    ap.add_argument('--cloud', default='S')
    ap.add_argument('--v2', action='store_true')
    ap.add_argument('--tile-order', default='', help='comma-separated chunk log2 values, e.g. 12,16,31')
    ap.add_argument('--reps', type=int, default=10)
    args = ap.parse_args()
    dev = torch.device(os.environ.get('ISF_TOOL_DEVICE', 'cuda:0'))      # 'cpu' only for the emulator dry run in tests/
    net = make_net(0)
    net = net.to(dev).eval()
    pts = [synth.make_points(args.points, args.cloud, 8).to(dev)]
    l2i, aug, laug = synth.make_calib(1)
    feats = [f.to(dev) for f in synth.make_img_feats(1)]
    metas = [dict(input_shape=(256, 704))]
    kw = dict(lidar2img=l2i, img_aug_matrix=aug, lidar_aug_matrix=laug)
    orders = [None] + [int(t) for t in args.tile_order.split(',') if t]
    variants = [('v1' + ('' if o is None else '+ord%d' % o), 0, o) for o in orders]
    if args.v2:
        variants += [('v2' + ('' if o is None else '+ord%d' % o), 1, o) for o in orders]
    cols, ref_out = {}, None
    for name, v2, order in variants:
        L.set_option('gemm_v2', v2)
        spmod.TILE_ORDER = order
        try:
            net(pts, feats, metas, **kw)
            prof = []
            spmod.PROFILE = prof
            out, _ = net(pts, feats, metas, **kw)
            spmod.PROFILE = None
            torch.cuda.synchronize()
            if ref_out is None:
                ref_out = [o.clone() for o in out]
            else:
                same = all(torch.equal(a, b) for a, b in zip(ref_out, out))
                print('%s: neck output %s the default path' % (name, 'bit-identical to' if same else 'DIFFERS from'))
            cols[name] = spmod.time_launches(prof, reps=args.reps)
        finally:
            spmod.PROFILE = None
            spmod.TILE_ORDER = None
            L.set_option('gemm_v2', 0)
    names = list(cols)
    print('%3s %8s %4s %4s | %s' % ('#', 'rows', 'cin', 'cout', ' | '.join('%-16s' % (n + ' ms/TF') for n in names)))
    tot = {n: 0.0 for n in names}
    base = cols[names[0]]
    for i in range(len(base)):
        cells = []
        for n in names:
            ms, useful = cols[n][i][0], cols[n][i][2]
            tot[n] += ms
            cells.append('%7.3f %7.0f ' % (ms, useful / ms / 1e9))
        print('%3d %8d %4d %4d | %s' % (i, base[i][5], base[i][3], base[i][4], ' | '.join(cells)))
    print('total ms: ' + '  '.join('%s %.3f' % (n, t) for n, t in tot.items()))


if __name__ == '__main__':
    main()
