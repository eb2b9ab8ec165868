"""Per-layer timing of the sparse-conv implicit GEMMs of one frame (round-2 iteration aid, GPU only):
   python scripts/bench_gemm_layers.py [--points 300000] [--opt gemm_mt=2,gemm_debug=1] [--tile-order none,16,31]
prints, for each of the 21 launches, rows / Cin / Cout / ms and useful TFLOP/s for the default kernel and, side by
side, variants: library options (--opt name=value, isf_set_option; each listed option is one variant) and the
neighbour-pattern tile order (--tile-order, csrc/spconv.cu isf_spconv_tile_order; chunk log2 per variant, 31 = one
global sort = the default, none = storage order).  Also checks whether every variant's neck output is bit-identical
to the first one (gemm_debug variants are timing experiments and are expected to differ)."""
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
    ap.add_argument('--opt', default='', help='comma-separated library options, each one a variant, e.g. gemm_mt=2,gemm_debug=1')
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
    default_order = spmod.TILE_ORDER
    orders = [default_order] + [None if t == 'none' else int(t) for t in args.tile_order.split(',') if t]
    oname = lambda o: 'storage' if o is None else 'ord%d' % o
    variants = [(oname(o), None, o) for o in orders]
    variants += [(t, t.split('='), default_order) for t in args.opt.split(',') if t]
    defaults = {'gemm_mt': 3, 'gemm_debug': 0, 'gemm_cluster': 1, 'nbr_rows': 1, 'gemm_ln_fast': 1, 'gemm_v4': 2, 'gemm_zskip': 0, 'gemm_wpc': 0, 'gemm_rev': 0, 'gemm_v4_32': 1, 'gemm_pair': 128, 'gemm_mma2': 1, 'gemm_trim': 0}
    cols, ref_out = {}, None
    for name, opt, order in variants:
        if opt:
            L.set_option(opt[0], int(opt[1]))
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
                print('%s: neck output %s the first variant' % (name, 'bit-identical to' if same else 'DIFFERS from'))
            cols[name] = spmod.time_launches(prof, reps=args.reps)
        finally:
            spmod.PROFILE = None
            spmod.TILE_ORDER = default_order
            if opt:
                L.set_option(opt[0], defaults.get(opt[0], 0))
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
