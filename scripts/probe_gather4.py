import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isfusion_b200 import _lib as L
torch.manual_seed(0)
rows, cols = 1000, 128
a = torch.randn(rows, cols, device='cuda').half()
idx = torch.randint(0, rows, (128,), dtype=torch.int32, device='cuda')
idx[5] = rows + 7      # out of bounds -> expect zeros
idx[9] = -1            # negative -> expect zeros?
for box_rows in ([int(sys.argv[1])] if len(sys.argv) > 1 else (1, 4)):
    for col0 in (0, 64):
        out = torch.full((128, 64), 7.0, dtype=torch.float16, device='cuda')
        st = torch.zeros(1, dtype=torch.int32, device='cuda')
        try:
            L.check(L.lib().isf_debug_tma_gather4(L.ptr(a), rows, cols, L.ptr(idx), col0, box_rows, L.ptr(out), L.ptr(st), L.stream()))
            torch.cuda.synchronize()
        except Exception as e:
            print('box_rows', box_rows, 'col0', col0, 'ERROR', e); continue
        raw = out.view(128, 8, 8)                      # [row][16B chunk][8 halfs] as stored
        exp = torch.zeros(128, 8, 8, dtype=torch.float16, device='cuda')
        for r in range(128):
            i = int(idx[r])
            src = a[i, col0:col0 + 64].view(8, 8) if 0 <= i < rows else torch.zeros(8, 8, dtype=torch.float16, device='cuda')
            for c in range(8):
                exp[r, c ^ (r & 7)] = src[c]
        ok = torch.equal(raw, exp)
        bad = (raw != exp).any(-1).any(-1).nonzero().flatten().tolist()
        print('box_rows', box_rows, 'col0', col0, 'status', int(st), 'match', ok, 'bad rows', bad[:12])
