"""Stage timing probe of the LiDAR branch (voxelize -> VFE -> SparseEncoder) on one B200 (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth
from isfusion_b200.voxel_layer import Voxelization
from isfusion_b200.voxel_encoder import DynamicVFE
from isfusion_b200.sparse_encoder import SparseEncoder
from isfusion_b200 import _lib as L

VS, PR = [0.075, 0.075, 0.2], synth.PC_RANGE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
if len(sys.argv) > 3:
    L.set_option('gemm_cluster', int(sys.argv[3]))
kind = sys.argv[2] if len(sys.argv) > 2 else 'S'
torch.manual_seed(0)
pts = synth.make_points(n, kind, 0).cuda()
vox = Voxelization(VS, PR, -1, (-1, -1))
pil = Voxelization([0.6, 0.6, 8.0], PR, 12, (30000, 60000)).eval()
vfe = synth.randomize_bn(DynamicVFE(5, [64, 64], False, True, True, VS, PR, dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01))).cuda().eval()
enc = synth.randomize_bn(SparseEncoder(in_channels=64, sparse_shape=[41, 1440, 1440], base_channels=32, output_channels=256,
                    encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
                    encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')).cuda().eval()

def timed(fn, name, reps=5):
    for _ in range(2): out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print('%-28s %8.3f ms' % (name, e0.elapsed_time(e1) / reps), flush=True)
    return out

c3 = timed(lambda: vox(pts), 'dynamic_voxelize')
coors = torch.nn.functional.pad(c3, (1, 0), value=0)
vf, vc = timed(lambda: vfe(pts, coors, batch_size=1), 'dynamic VFE (fused)')
print('points', pts.shape[0], 'voxels', vf.shape[0])
timed(lambda: pil(pts), 'hard_voxelize (pillars)')
sp, feats, _ = timed(lambda: enc(vf, vc, 1), 'SparseEncoder total')
for i, f in enumerate(feats):
    print(' stage', i, 'V =', f.features.shape[0], 'C =', f.features.shape[1], 'shape', f.spatial_shape)
# per-layer: time each conv of the encoder separately with its rulebook prebuilt
import isfusion_b200.spconv as sp_mod
x = sp_mod.SparseConvTensor(vf.contiguous(), vc.contiguous(), [41, 1440, 1440], 1)
def walk(mods, x):
    for m in mods:
        if isinstance(m, sp_mod.SparseConvolution) or m.__class__.__name__ == 'SparseBasicBlock':
            pass
    return x
tot_flops = 0
cur = x
def conv_time(conv, cur, name):
    global tot_flops
    rb = timed(lambda: sp_mod.build_rulebook(cur, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.subm), name + ' rulebook', 3)
    pairs = int((rb.nbr[:, :rb.num_out] >= 0).sum().item())
    tile_occ = (rb.nbr >= 0).view(rb.taps, -1, 128).any(-1).float().mean().item()
    w = conv.packed_weight()
    f = timed(lambda: sp_mod.implicit_gemm(cur.features, rb, w, conv.in_channels, conv.out_channels, relu=True), name + ' gemm', 5)
    flops = 2.0 * pairs * conv.in_channels * conv.out_channels
    dense_flops = 2.0 * rb.num_out * rb.taps * conv.in_channels * conv.out_channels
    print('    V_in %d V_out %d pairs %d (%.0f%% of slots, %.0f%% of (tile,tap) non-empty)  useful %.1f GF  issued %.1f GF' % (cur.features.shape[0], rb.num_out, pairs, 100.0 * pairs / max(rb.num_out * rb.taps, 1), 100 * tile_occ, flops / 1e9, dense_flops / 1e9))
    out = sp_mod.SparseConvTensor(f, rb.out_indices, rb.out_shape, cur.batch_size)
    return out
cur = conv_time(enc.conv_input[0], cur, 'conv_input 64->32')
for si, stage in enumerate(enc.encoder_layers):
    for bi, blk in enumerate(stage):
        if blk.__class__.__name__ == 'SparseBasicBlock':
            if bi == 0:
                conv_time(blk.conv1, cur, 'stage%d subm %d->%d' % (si + 1, blk.conv1.in_channels, blk.conv1.out_channels))
        else:
            cur = conv_time(blk[0], cur, 'stage%d down %d->%d' % (si + 1, blk[0].in_channels, blk[0].out_channels))
cur = conv_time(enc.conv_out[0], cur, 'conv_out')
