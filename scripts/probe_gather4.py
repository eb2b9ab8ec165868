"""Bring-up probe of TMA tile::gather4 (csrc/debug_tma.cu): every (variant, out-of-bounds rows) case in its own process,
because a faulting TMA kills the CUDA context.   python scripts/probe_gather4.py            (GPU only)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(variant, oob):
    sys.path.insert(0, ROOT)
    import torch
    from isfusion_b200 import _lib as L
    torch.manual_seed(0)
    rows, cols = 1000, 128
    a = torch.randn(rows, cols, device='cuda').half()
    idx = torch.randint(0, rows, (128,), dtype=torch.int32, device='cuda')
    if oob & 1:
        idx[5] = rows + 7
    if oob & 2:
        idx[9] = -1
    stride = 1024 if variant & 2 else 512
    for col0 in (0, 64):
        out = torch.zeros((32768 // 2,), dtype=torch.float16, device='cuda')
        st = torch.zeros(1, dtype=torch.int32, device='cuda')
        L.check(L.lib().isf_debug_tma_gather4(L.ptr(a), rows, cols, L.ptr(idx), col0, variant, L.ptr(out), L.ptr(st), L.stream()))
        torch.cuda.synchronize()
        raw = out.view(-1, 8, 8)                     # [128-byte row][16 B piece][8 halfs] as stored
        nbad, first = 0, []
        for g in range(32):
            for j in range(4):
                r = 4 * g + j
                i = int(idx[r])
                src = a[i, col0:col0 + 64].view(8, 8) if 0 <= i < rows else torch.zeros(8, 8, dtype=torch.float16, device='cuda')
                srow = (g * stride) // 128 + j
                exp = torch.empty_like(src)
                for c in range(8):
                    exp[(c ^ (srow & 7)) if not variant & 4 else c] = src[c]
                if not torch.equal(raw[srow], exp):
                    nbad += 1
                    if len(first) < 6:
                        first.append(r)
        print('variant', variant, 'oob', oob, 'col0', col0, 'status', int(st), 'bad rows', nbad, first, flush=True)


if __name__ == '__main__':
    if len(sys.argv) == 3:
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for variant in range(8):
            good = False
            for oob in (0, 1, 2):
                if oob and not good:
                    break
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(variant), str(oob)], capture_output=True, text=True, timeout=120)
                tail = (r.stdout.strip() or '') + ('' if r.returncode == 0 else ' | rc=%d %s' % (r.returncode, r.stderr.strip().splitlines()[-1][:200] if r.stderr.strip() else ''))
                print(tail, flush=True)
                good = r.returncode == 0 and 'bad rows 0 ' in r.stdout and 'status 1' in r.stdout
