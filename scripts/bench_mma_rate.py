"""tcgen05.mma issue rate by instruction shape / operand source / CTA group on this GPU (csrc/debug_tma.cu isf_debug_mma_rate):
   python scripts/bench_mma_rate.py
prints cycles per instruction per CTA and the chip-wide TFLOP/s that rate means, then the accumulator layout of M = 64 and M = 128
instructions (which TMEM lane holds which row).  Iteration aid / evidence for DESIGN.md 3.1c, not on the product path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from isfusion_b200 import _lib as L  # noqa: E402


def rate(m, n, cg, a_tmem, blocks, chains=1, issuers=1, iters=4000):
    out = torch.zeros(blocks, device='cuda')
    L.check(L.lib().isf_debug_mma_rate(m, n, cg, a_tmem, chains, issuers, blocks, iters, L.ptr(out), None, L.stream()))
    torch.cuda.synchronize()
    v = out[out > 0]
    return v.mean().item(), v.min().item(), v.max().item()


def main():
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = 1965.0
    print('%-64s %9s %9s %9s %12s' % ('variant', 'cyc/instr', 'min', 'max', 'chip TFLOP/s'))
    cases = []      # (m, n, cta_group, a_tmem, chains, issuers, CTAs per SM)
    for n in (32, 64, 128, 256):
        cases.append((128, n, 1, 0, 1, 1, 1))
    for n, ch in ((256, 2), (128, 2), (128, 4), (64, 4), (64, 8)):
        cases.append((128, n, 1, 0, ch, 1, 1))
    for n, iss in ((256, 2), (128, 2), (128, 4), (64, 2), (64, 4), (32, 4)):
        cases.append((128, n, 1, 0, 1, iss, 1))
    for n, ctas in ((256, 2), (128, 2), (128, 3), (128, 4), (64, 4)):
        cases.append((128, n, 1, 0, 1, 1, ctas))
    cases += [(128, 128, 1, 0, 1, 2, 2), (128, 64, 1, 0, 1, 4, 2)]
    cases += [(64, 256, 1, 0, 1, 1, 1), (64, 256, 1, 0, 1, 2, 1)]
    cases += [(128, 256, 2, 0, 1, 1, 1), (128, 256, 2, 0, 1, 2, 1), (128, 128, 2, 0, 1, 2, 1), (128, 128, 2, 0, 1, 1, 2)]
    cases += [(128, 128, 1, 1, 1, 1, 1), (128, 128, 1, 1, 1, 2, 1), (128, 64, 1, 1, 1, 4, 1)]
    # a_tmem = 2: whole warp converged, instruction predicated by elect.sync
    for n in (32, 64, 128, 256):
        cases.append((128, n, 1, 2, 1, 1, 1))
    cases += [(128, 128, 1, 2, 1, 2, 1), (128, 64, 1, 2, 1, 4, 1), (128, 128, 1, 2, 1, 1, 2), (128, 256, 1, 2, 1, 2, 1),
              (128, 256, 2, 2, 1, 1, 1), (128, 128, 2, 2, 1, 1, 2), (64, 256, 1, 2, 1, 1, 1)]
    # a_tmem = 4: one elect.sync per group of four instructions + commit (tc::umma_chunk)
    for n in (32, 64, 128, 256):
        cases.append((128, n, 1, 4, 1, 1, 1))
    cases += [(128, 128, 1, 4, 1, 2, 1), (128, 64, 1, 4, 1, 4, 1), (128, 64, 1, 4, 1, 2, 2), (128, 128, 1, 4, 1, 1, 2), (128, 256, 2, 4, 1, 1, 1),
              (128, 128, 2, 4, 1, 1, 2), (128, 64, 2, 4, 1, 1, 2)]
    for m, n, cg, at, ch, iss, ctas in cases:
        name = 'M=%d N=%d %s cta_group::%d, %d chain(s) x %d issuer(s), %d CTA(s)/SM' % (cg * m, n, {0: 'SS', 1: 'TS', 2: 'SS elect', 4: 'SS elect/4'}[at], cg, ch, iss, ctas)
        blocks = sms * ctas
        blocks -= blocks % cg
        try:
            mean, lo, hi = rate(m, n, cg, at, blocks, ch, iss)
        except Exception as e:      # noqa: BLE001
            print('%-64s failed: %s' % (name, str(e)[:80]))
            continue
        streams = blocks / cg * iss
        flops = 2.0 * (cg * m) * n * 16
        print('%-64s %9.1f %9.1f %9.1f %12.0f' % (name, mean, lo, hi, flops * streams / mean * mhz * 1e6 / 1e12))
    for m in (64, 128):
        out = torch.zeros(1, device='cuda')
        probe = torch.full((256,), -1.0, device='cuda')
        L.check(L.lib().isf_debug_mma_rate(m, 32, 1, 0, 1, 1, 1, 1, L.ptr(out), L.ptr(probe), L.stream()))
        torch.cuda.synchronize()
        p = probe.cpu().tolist()
        print('accumulator layout M=%d: TMEM lane -> row + 1 (column 0): %s' % (m, ' '.join('%d' % int(x) for x in p[:128])))


if __name__ == '__main__':
    main()
