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


def tile_order_model(valid, chunk_log2=31, ksize=(3, 3, 3), drop_bits=0):
    """CPU model of isf_spconv_tile_order: stable sort of the rows by (row >> chunk_log2, occupancy mask), the
    `drop_bits` least significant key bits left out (isf_set_option("tile_order_drop_bits"))."""
    pos = library_tap_bits(ksize)
    key = np.zeros(valid.shape[1], np.uint64)
    for k in range(valid.shape[0]):
        key |= valid[k].astype(np.uint64) << np.uint64(pos[k])
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=300000)
    ap.add_argument('--cloud', default='S')
    ap.add_argument('--seed', type=int, default=8)
    args = ap.parse_args()
    pts = synth.make_points(args.points, args.cloud, args.seed)
    vs, pr = [0.075, 0.075, 0.2], synth.PC_RANGE
    c = voxel_ops.dynamic_voxelize(pts, vs, pr)
    c = c[(c >= 0).all(1)]
    idx = torch.unique(torch.cat([torch.zeros(len(c), 1, dtype=c.dtype), c], 1), dim=0).int()
    shape = [41, 1440, 1440]
    stages = [('stage1 SubM 32', None), ('stage2 SubM 64', (1, 1, 1)), ('stage3 SubM 128', (1, 1, 1)),
              ('stage4 SubM 256', (0, 1, 1))]
    chans = [32, 64, 128, 256]
    for (name, pad), ch in zip(stages, chans):
        if pad is not None:
            idx, shape, _ = so.get_indice_pairs(idx, 1, shape, 3, 2, list(pad), 1, False)
        _, _, nbr = so.get_indice_pairs(idx, 1, shape, 3, 1, 1, 1, True)
        valid = (nbr >= 0).numpy()
        useful = int(valid.sum())
        print('\n%s: V = %d, pairs = %d (%.1f per row), grid %s' % (name, len(idx), useful, useful / len(idx), shape))
        base = None
        cand = orders(idx.numpy())
        for cl in (12, 14, 16, 31):
            cand['neighbour-pattern sort (isf_spconv_tile_order), chunk 2^%d' % cl] = tile_order_model(valid, cl)
        for oname, perm in cand.items():
            iss, tiles = issued_slots(valid, perm)
            base = base or iss
            print('   %-62s issued %9d  fill %.3f  offsets/tile %5.2f  issued vs current %.3f'
                  % (oname, iss, useful / iss, iss / TILE / tiles, iss / base))


if __name__ == '__main__':
    main()
