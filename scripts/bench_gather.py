"""Micro-benchmark of the A-operand gather paths (csrc/debug_tma.cu gather_bench_kernel): cycles per 128-row x 128-B chunk
per SM for TMA tile::gather4 (three issue patterns), a plain TMA tile load and the LDGSTS producer pattern, on a
stage-3-sized feature matrix (272 k rows x 128 ch) with rows drawn from a real neighbour-like window.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from isfusion_b200 import _lib as L  # noqa: E402

torch.manual_seed(0)
rows, cols = 272826, 128
a = torch.randn(rows, cols, device='cuda').half()
n_idx = 1 << 20
base = torch.arange(n_idx, device='cuda') // 4 % (rows - 4096)          # neighbour tables are locally coherent: rows of a tile
idx_local = (base + torch.randint(0, 4096, (n_idx,), device='cuda')).int()  # come from a window of a few thousand rows
idx_rand = torch.randint(0, rows, (n_idx,), dtype=torch.int32, device='cuda')
idx_half = idx_local.clone()
idx_half[torch.rand(n_idx, device='cuda') < 0.4] = -1                    # 40 % missing neighbours (zero fill)
names = {0: 'gather4 32 lanes', 1: 'gather4 1 lane', 2: 'gather4 4 warps x 8', 3: 'TMA tile (no gather)', 4: 'LDGSTS 256 thr'}
print('%-22s %-7s %6s | %10s %10s %10s  (cycles per 16 KB chunk per SM; 256 = what a N=128 MMA chunk takes)' % ('mode', 'stages', 'blocks', 'local', 'random', '40% miss'))
for blocks in (148, 296):
    for stages in (4, 8, 12):
        if blocks == 296 and stages > 4:
            continue
        for mode in (0, 1, 2, 3, 4):
            res = []
            for idx in (idx_local, idx_rand, idx_half):
                out = torch.zeros(blocks, device='cuda')
                for _ in range(2):
                    L.check(L.lib().isf_debug_gather_bench(L.ptr(a), rows, cols, L.ptr(idx), n_idx, mode, stages, 2000, blocks, L.ptr(out), L.stream()))
                torch.cuda.synchronize()
                res.append(out.mean().item())
            print('%-22s %-7d %6d | %10.0f %10.0f %10.0f' % (names[mode], stages, blocks, *res), flush=True)
