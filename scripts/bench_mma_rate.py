"""tcgen05.mma issue rate by instruction shape / operand source / CTA group on this GPU (csrc/debug_tma.cu isf_debug_mma_rate):
   python scripts/bench_mma_rate.py
prints cycles per instruction per CTA and the chip-wide TFLOP/s that rate means, then the accumulator layout of M = 64 and M = 128
instructions (which TMEM lane holds which row).  Iteration aid / evidence for DESIGN.md 3.1c, not on the product path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from isfusion_b200 import _lib as L  # noqa: E402


def rate(m, n, cg, a_tmem, blocks, iters=4000):
    out = torch.zeros(blocks, device='cuda')
    L.check(L.lib().isf_debug_mma_rate(m, n, cg, a_tmem, blocks, iters, L.ptr(out), None, L.stream()))
    torch.cuda.synchronize()
    v = out[out > 0]
    return v.mean().item(), v.min().item(), v.max().item()


def main():
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = 1965.0
    print('%-42s %9s %9s %9s %12s' % ('variant', 'cyc/instr', 'min', 'max', 'chip TFLOP/s'))
    cases = []
    for n in (32, 64, 128, 256):
        cases.append(('M=128 N=%d SS cta_group::1, 1 CTA/SM' % n, 128, n, 1, 0, sms))
    for n in (64, 128):
        cases.append(('M=128 N=%d SS cta_group::1, 2 CTAs/SM' % n, 128, n, 1, 0, 2 * sms))
    for n in (64, 128, 256):
        cases.append(('M=64  N=%d SS cta_group::1, 1 CTA/SM' % n, 64, n, 1, 0, sms))
    for n in (64, 128, 256):
        cases.append(('M=256 N=%d SS cta_group::2, 1 pair / 2 SMs' % n, 128, n, 2, 0, sms - sms % 2))
    for n in (32, 64, 128):
        cases.append(('M=128 N=%d TS (A in TMEM), 1 CTA/SM' % n, 128, n, 1, 1, sms))
    for name, m, n, cg, at, blocks in cases:
        try:
            mean, lo, hi = rate(m, n, cg, at, blocks)
        except Exception as e:      # noqa: BLE001
            print('%-42s failed: %s' % (name, str(e)[:80]))
            continue
        issuers = blocks / cg
        flops = 2.0 * (cg * m) * n * 16
        print('%-42s %9.1f %9.1f %9.1f %12.0f' % (name, mean, lo, hi, flops * issuers / mean * mhz * 1e6 / 1e12))
    for m in (64, 128):
        out = torch.zeros(1, device='cuda')
        probe = torch.full((256,), -1.0, device='cuda')
        L.check(L.lib().isf_debug_mma_rate(m, 32, 1, 0, 1, 1, L.ptr(out), L.ptr(probe), L.stream()))
        torch.cuda.synchronize()
        p = probe.cpu().tolist()
        print('accumulator layout M=%d: TMEM lane -> row + 1 (column 0): %s' % (m, ' '.join('%d' % int(x) for x in p[:128])))


if __name__ == '__main__':
    main()
