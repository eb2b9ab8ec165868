"""ncu target: two SparseEncoder forwards on one 300k-point frame (21 tc::gemm_kernel launches each, fixed order:
conv_input, stage1 4x32->32 + down, stage2 4x64->64 + down, stage3 4x128->128 + down, stage4 4x256->256, conv_out)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth
from isfusion_b200.voxel_layer import Voxelization
from isfusion_b200.voxel_encoder import DynamicVFE
from isfusion_b200.sparse_encoder import SparseEncoder
from isfusion_b200 import _lib as L
if len(sys.argv) > 1:
    L.set_option('gemm_cluster', int(sys.argv[1]))
VS, PR = [0.075, 0.075, 0.2], synth.PC_RANGE
torch.manual_seed(0)
pts = synth.make_points(300000, 'S', 0).cuda()
vox = Voxelization(VS, PR, -1, (-1, -1))
vfe = synth.randomize_bn(DynamicVFE(5, [64, 64], False, True, True, VS, PR, dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01))).cuda().eval()
enc = synth.randomize_bn(SparseEncoder(in_channels=64, sparse_shape=[41, 1440, 1440], base_channels=32, output_channels=256,
                    encoder_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256)),
                    encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')).cuda().eval()
coors = torch.nn.functional.pad(vox(pts), (1, 0), value=0)
vf, vc = vfe(pts, coors, batch_size=1)
for _ in range(2):
    enc(vf, vc, 1, dense=False)
torch.cuda.synchronize()
print('done')
