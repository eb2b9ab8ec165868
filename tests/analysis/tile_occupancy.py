"""CPU analysis (no GPU): how many zero-filled slots does the output-stationary implicit GEMM multiply, per stage,
and how does that change with the assignment of rows to 128-row tiles?

  python tests/analysis/tile_occupancy.py [--points 300000] [--cloud S]

For every SubM stage of the SparseEncoder on the synthetic frame it builds the neighbour table with the oracle and
reports, for each candidate tile assignment (a permutation of the rows; storage order is untouched, the kernel
would read nbr[k][perm[i]] and scatter row perm[i] in its epilogue):
  useful = pairs, issued = 128 * sum_tiles #(kernel offsets with >= 1 neighbour in the tile)   (what the kernel gathers
  and multiplies: chunks whose offset has no neighbour in the whole tile are skipped), fill = useful / issued.
TEST / ANALYSIS INFRASTRUCTURE (lives under tests/: it imports oracle/ for the rulebooks)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import synth  # noqa: E402
from oracle import spconv_ops as so, voxel_ops  # noqa: E402

TILE = 128


def issued_slots(nbr_valid, perm):
    """nbr_valid [K,V] bool; perm [V] -> issued slots with 128-row tiles over the permuted rows."""
    k, v = nbr_valid.shape
    pad = (-v) % TILE
    m = nbr_valid[:, perm]
    if pad:
        m = np.concatenate([m, np.zeros((k, pad), bool)], 1)
    any_ = m.reshape(k, -1, TILE).any(2)          # [K, tiles]
    return int(any_.sum()) * TILE, any_.shape[1]


def morton3(z, y, x, bits=11):
    def spread(a):
        a = a.astype(np.uint64)
        r = np.zeros_like(a)
        for i in range(bits):
            r |= ((a >> np.uint64(i)) & np.uint64(1)) << np.uint64(3 * i)
        return r
    return (spread(z) << np.uint64(2)) | (spread(y) << np.uint64(1)) | spread(x)


def library_tap_bits(ksize=(3, 3, 3)):
    """bit position of every kernel offset in the sort key of isf_spconv_tile_order (host-only ABI call, no GPU)."""
    import ctypes
    from isfusion_b200 import _lib as L
    taps = ksize[0] * ksize[1] * ksize[2]
    pos = (ctypes.c_ubyte * 32)()
    L.check(L.lib().isf_spconv_tile_order_bits(L.host_i32(ksize), taps, pos))
    return [int(pos[k]) for k in range(taps)]


def occupancy_masks(valid, ksize=(3, 3, 3)):
    """occupancy mask of every row with kernel offset k at bit library_tap_bits()[k]."""
    pos = library_tap_bits(ksize)
    mask = np.zeros(valid.shape[1], np.uint64)
    for k in range(valid.shape[0]):
        mask |= valid[k].astype(np.uint64) << np.uint64(pos[k])
    return mask


def mask_rank(mask):
    """tile_order_rank of csrc/spconv.cu: the mask read as a binary-reflected Gray code (neighbouring keys differ in one offset)."""
    b = mask.copy()
    for s in (1, 2, 4, 8, 16):
        b ^= b >> np.uint64(s)
    return b


def tile_order_model(valid, chunk_log2=31, ksize=(3, 3, 3), drop_bits=0):
    """CPU model of isf_spconv_tile_order: stable sort of the rows by (row >> chunk_log2, occupancy mask), the
    `drop_bits` least significant key bits left out (isf_set_option("tile_order_drop_bits"))."""
    key = mask_rank(occupancy_masks(valid, ksize))
    if chunk_log2 < 31:
        key |= (np.arange(valid.shape[1], dtype=np.uint64) >> np.uint64(chunk_log2)) << np.uint64(valid.shape[0])
    return np.argsort(key >> np.uint64(drop_bits), kind='stable')


def orders(idx):
    b, z, y, x = [idx[:, i].astype(np.int64) for i in range(4)]
    out = {'sorted (b,z,y,x) [current]': np.arange(len(idx))}
    for by, bx in ((4, 32), (8, 16), (16, 8), (8, 8), (4, 4)):
        key = ((((b * 64 + z) * 4096 + y // by) * 4096 + x // bx) * by + y % by) * bx + x % bx
        out['z-slice blocks %dx%d (y,x)' % (by, bx)] = np.argsort(key, kind='stable')
    for bz, by, bx in ((2, 8, 8), (4, 8, 4), (2, 4, 16), (3, 8, 8)):
        key = (((((b * 64 + z // bz) * 4096 + y // by) * 4096 + x // bx) * bz + z % bz) * by + y % by) * bx + x % bx
        out['3-D blocks %dx%dx%d (z,y,x)' % (bz, by, bx)] = np.argsort(key, kind='stable')
    out['morton (z,y,x)'] = np.argsort((b.astype(np.uint64) << np.uint64(40)) | morton3(z, y, x), kind='stable')
    # column-major: all z of one (y,x) column together, columns in 8x8 blocks
    key = ((((b * 4096 + y // 8) * 4096 + x // 8) * 8 + y % 8) * 8 + x % 8) * 64 + z
    out['columns (z fastest) in 8x8 (y,x) blocks'] = np.argsort(key, kind='stable')
    return out


def gathered_slots_4row(valid, perm):
    """slots actually copied when a 4-row warp copy is skipped if none of its rows has a neighbour (v2 zero-skip)."""
    k, v = valid.shape
    pad = (-v) % TILE
    m = valid[:, perm]
    if pad:
        m = np.concatenate([m, np.zeros((k, pad), bool)], 1)
    return int(m.reshape(k, -1, TILE // 4, 4).any(3).sum()) * 4


SUMMARY = []      # (name, useful slots, issued slots now, issued slots with the global tile order)
VALID = []        # (name, valid [K, V], ksize) of every rulebook, in launch order

# measured per-launch durations (us) of the 21 sparse-conv launches of this very frame (seed 8), from the ncu table
# profiles/own_kernels_r01.txt (ids 15-19, 23, 25-28, 32, 34-37, 41, 43-46, 50): (rulebook index, Cin, Cout, us)
MEASURED_R01 = [(0, 64, 32, 100.5), (0, 32, 32, 80.1), (0, 32, 32, 84.8), (0, 32, 32, 79.5), (0, 32, 32, 84.7), (1, 32, 64, 103.0),
                (2, 64, 64, 160.6), (2, 64, 64, 167.0), (2, 64, 64, 161.5), (2, 64, 64, 170.6), (3, 64, 128, 178.5),
                (4, 128, 128, 285.2), (4, 128, 128, 303.1), (4, 128, 128, 283.4), (4, 128, 128, 302.6), (5, 128, 256, 177.9),
                (6, 256, 256, 257.6), (6, 256, 256, 274.0), (6, 256, 256, 259.2), (6, 256, 256, 274.9), (7, 256, 256, 44.7)]


def chunk_blocks(valid, perm, cin):
    """128-row x 64-K MMA blocks the kernel walks: a 64-wide K chunk covers 64/cin kernel offsets (cin < 64) or a
    cin/64-th of one (cin >= 64) and is skipped for a tile when none of its offsets has a neighbour there."""
    k, v = valid.shape
    m = valid[:, perm]
    pad = (-v) % TILE
    if pad:
        m = np.concatenate([m, np.zeros((k, pad), bool)], 1)
    t = m.reshape(k, -1, TILE).any(2)
    if cin < 64:
        g = 64 // cin
        kp = (-k) % g
        if kp:
            t = np.concatenate([t, np.zeros((kp, t.shape[1]), bool)], 0)
        return int(t.reshape(-1, g, t.shape[1]).any(1).sum())
    return int(t.sum()) * (cin // 64)


def predict():
    """ties the round-1 measurement to the pacing model of DESIGN.md 3.1: cycles per walked block per SM of every launch
    (1965 MHz, 148 SMs), and the launch times the same cost per block predicts under the global tile order."""
    print('\nper-launch cost model from the round-1 ncu durations (cold-cache, serialised: upper bounds):')
    print('   %-4s %-10s %9s %8s %14s | %9s %8s' % ('#', 'Cin->Cout', 'blocks', 'us', 'cyc/block/SM', 'blocks*', 'us*'))
    tot = tot_new = 0.0
    for i, (rb, cin, cout, us) in enumerate(MEASURED_R01):
        name, valid, ksize = VALID[rb]
        b0 = chunk_blocks(valid, np.arange(valid.shape[1]), cin)
        b1 = chunk_blocks(valid, tile_order_model(valid, 31, ksize), cin)
        cyc = us * 1965.0 * 148 / b0
        tot += us
        tot_new += us * b1 / b0
        print('   %-4d %-10s %9d %8.1f %14.0f | %9d %8.1f' % (i, '%d->%d' % (cin, cout), b0, us, cyc, b1, us * b1 / b0))
    print('   sum of the 21 launches: %.0f us measured -> %.0f us predicted with the global tile order (x%.3f) if the cost per '
          'walked block stays what it is' % (tot, tot_new, tot_new / tot))


def report(name, idx_np, valid, ksize=(3, 3, 3), spatial=True):
    useful = int(valid.sum())
    VALID.append((name, valid, ksize))
    SUMMARY.append((name, useful, issued_slots(valid, np.arange(valid.shape[1]))[0],
                    issued_slots(valid, tile_order_model(valid, 31, ksize))[0]))
    v = valid.shape[1]
    print('\n%s: rows = %d, pairs = %d (%.1f per row)' % (name, v, useful, useful / max(v, 1)))
    cand = orders(idx_np) if spatial else {'sorted (b,z,y,x) [current]': np.arange(v)}
    for cl in (12, 14, 16, 31):
        cand['neighbour-pattern sort (isf_spconv_tile_order), chunk 2^%d' % cl] = tile_order_model(valid, cl, ksize)
    base = None
    for oname, perm in cand.items():
        iss, tiles = issued_slots(valid, perm)
        base = base or iss
        print('   %-62s issued %9d  fill %.3f  offsets/tile %5.2f  issued vs current %.3f  | 4-row zero-skip copies vs current issued %.3f'
              % (oname, iss, useful / iss, iss / TILE / tiles, iss / base, gathered_slots_4row(valid, perm) / base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=300000)
    ap.add_argument('--cloud', default='S')
    ap.add_argument('--seed', type=int, default=8)
    ap.add_argument('--all-orders', action='store_true', help='also list the spatial re-blockings (they do not help)')
    ap.add_argument('--config5', action='store_true', help='BASELINE configs[4] geometry: 0.05 m voxels, +-51.2 m, 2048^2 x 41 grid (use --points 1000000)')
    args = ap.parse_args()
    vs, pr = ([0.05, 0.05, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]) if args.config5 else ([0.075, 0.075, 0.2], synth.PC_RANGE)
    pts = synth.make_points(args.points, args.cloud, args.seed, pr)
    c = voxel_ops.dynamic_voxelize(pts, vs, pr)
    c = c[(c >= 0).all(1)]
    idx = torch.unique(torch.cat([torch.zeros(len(c), 1, dtype=c.dtype), c], 1), dim=0).int()
    shape = [41, 2048, 2048] if args.config5 else [41, 1440, 1440]
    stages = [('stage1 SubM 32', None), ('stage2 SubM 64', (1, 1, 1)), ('stage3 SubM 128', (1, 1, 1)), ('stage4 SubM 256', (0, 1, 1))]
    for name, pad in stages:
        if pad is not None:
            idx, shape, nbr_s = so.get_indice_pairs(idx, 1, shape, 3, 2, list(pad), 1, False)
            report('strided SparseConv3d into ' + name.split()[0] + ' (padding %s)' % (pad,), idx.numpy(), (nbr_s >= 0).numpy(), spatial=False)
        _, _, nbr = so.get_indice_pairs(idx, 1, shape, 3, 1, 1, 1, True)
        report(name + ' (grid %s)' % shape, idx.numpy(), (nbr >= 0).numpy(), spatial=args.all_orders)
    oidx, oshape, nbr_o = so.get_indice_pairs(idx, 1, shape, (3, 1, 1), (2, 1, 1), [0, 0, 0], 1, False)
    report('conv_out SparseConv3d (3,1,1)/(2,1,1)', oidx.numpy(), (nbr_o >= 0).numpy(), ksize=(3, 1, 1), spatial=False)
    # whole encoder: every launch weighted by its Cin x Cout (config:72-82): conv_input 64->32 and 4 x 32->32 on the stage-1
    # rulebook, strided 32->64, 4 x 64->64, strided 64->128, 4 x 128->128, strided 128->256, 4 x 256->256, conv_out 256->256
    weights = [64 * 32 + 4 * 32 * 32, 32 * 64, 4 * 64 * 64, 64 * 128, 4 * 128 * 128, 128 * 256, 4 * 256 * 256, 256 * 256]
    tot = np.array([[u * w, a * w, b * w] for (_, u, a, b), w in zip(SUMMARY, weights)], dtype=np.float64).sum(0) * 2
    print('\nwhole sparse encoder (21 launches): useful %.0f GFLOP, issued now %.0f GFLOP (fill %.3f), issued with the global '
          'tile order %.0f GFLOP (fill %.3f, x%.3f)' % (tot[0] / 1e9, tot[1] / 1e9, tot[0] / tot[1], tot[2] / 1e9, tot[0] / tot[2],
                                                         tot[2] / tot[1]))
    if args.points == 300000 and args.cloud == 'S' and args.seed == 8 and not args.config5:
        predict()


if __name__ == '__main__':
    main()
