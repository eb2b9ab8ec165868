"""Top CUDA source lines of an ncu report by warp-stall samples (needs -lineinfo and --import-source on):
   python scripts/ncu_source_top.py gpurun_out/gemm128_r02.ncu-rep [N]      (runs `ncu -i ... --page source --print-source cuda,sass --csv`)"""
import csv
import io
import subprocess
import sys


def main():
    rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(r for r in rows if r and r[0] == 'Line No')
    si = hdr.index('# Samples')
    stalls = [(j, h[6:]) for j, h in enumerate(hdr) if h.startswith('stall_') and '(Not' not in h]
    out, fname = [], ''
    for r in rows:
        if r and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        if len(r) <= si or not r[0].isdigit() or not r[si].isdigit():
            continue
        s = int(r[si])
        if s:
            st = sorted(((int(r[j]), name) for j, name in stalls if r[j].isdigit() and int(r[j])), reverse=True)[:3]
            out.append((s, fname, r[0], r[1].strip()[:105], ' '.join('%s=%d' % (nm, v) for v, nm in st)))
    tot = sum(o[0] for o in out)
    print('%s: %d samples' % (rep, tot))
    for s, f, l, src, st in sorted(out, reverse=True)[:n]:
        print('%5.1f%% %s:%-4s %-105s | %s' % (100.0 * s / tot, f, l, src, st))


if __name__ == '__main__':
    main()
