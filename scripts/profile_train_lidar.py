"""torch.profiler table of one LiDAR-branch training step (scripts/bench_train_lidar.py): where the step's time goes."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth  # noqa: E402
from __graft_entry__ import make_net  # noqa: E402

net = make_net(0).cuda()
vfe, enc = net.pts_voxel_encoder.train(), net.pts_middle_encoder.train()
pts = [synth.make_points(100000, 'S', i).cuda() for i in range(4)]
with torch.no_grad():
    voxels, coors = net.dynamic_voxelize(pts)
    keep = (coors[:, 1:] >= 0).all(1)
    voxels, coors = voxels[keep].contiguous(), coors[keep].contiguous()


def step():
    for p in list(vfe.parameters()) + list(enc.parameters()):
        p.grad = None
    vf, vc = vfe(voxels, coors)
    bev, _, _ = enc(vf, vc, 4, sorted_coors=True)
    (bev * 0.001).sum().backward()


step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=18, max_name_column_width=60))
print(prof.key_averages().table(sort_by='cpu_time_total', row_limit=12, max_name_column_width=60))
