#!/usr/bin/env python
"""Static report of every kernel in the built library (no GPU needed): registers, shared / local memory, spills as
`cuobjdump -res-usage` gives them, and the count of the SASS mnemonics that identify the Blackwell paths
(tcgen05 = UTCHMMA / LDTM / UTCBAR, TMA = UTMALDG, async gathers = LDGSTS, 256-bit global IO = .ENL2.256, mma.sync = HMMA, ldmatrix = LDSM, red.global = REDG).
    python scripts/sass_report.py > profiles/sass_rNN.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'is-fusion_b200', 'libisfusion_b200.so')
MNEMONICS = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTCBAR', 'UTCBAR.2CTA.MULTICAST', 'UTMALDG', 'UTMASTG', 'STAS', 'UTCATOMSWS', 'LDGSTS', 'SYNCS', 'HMMA', 'LDSM', 'REDG', 'ENL2.256', 'STL', 'LDL',
             'BAR.SYNC']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)           # drop the argument list
    return name


def main():
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        if cur and 'REG:' in line:
            usage[cur] = dict((k, int(v)) for k, v in re.findall(r'(\w+):(\d+)', line))
            cur = None
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    n_instr = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or '/*' not in line:
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(.*?);', line)
        if not m:
            continue
        ins = m.group(1)
        n_instr[cur] += 1
        op = ins.split()[1] if ins.startswith('@') and len(ins.split()) > 1 else ins.split()[0]
        for mn in MNEMONICS:
            if op.startswith(mn) or (mn == 'ENL2.256' and mn in op):
                counts[cur][mn] += 1
    names = demangle(sorted(usage))
    print('# static SASS / resource report of %s (sm_100a; `python scripts/sass_report.py`)' % os.path.relpath(LIB, ROOT))
    print('# columns: registers, static smem bytes, local (stack) bytes, SASS instructions, then mnemonic counts')
    rows = []
    for fn in sorted(usage, key=lambda f: short(names[f])):
        u = usage[fn]
        c = counts[fn]
        tags = ' '.join('%s=%d' % (mn, c[mn]) for mn in MNEMONICS if c[mn])
        rows.append((short(names[fn]), u.get('REG', 0), u.get('SHARED', 0), u.get('STACK', 0) + u.get('LOCAL', 0), n_instr[fn], tags))
    w = min(110, max(len(r[0]) for r in rows))
    for r in rows:
        print('%-*s reg=%-3d smem=%-6d local=%-4d instr=%-6d %s' % (w, r[0][:w], r[1], r[2], r[3], r[4], r[5]))
    spills = [r[0] for r in rows if r[3]]
    print('# kernels with local memory (spills / stack): %s' % (', '.join(spills) if spills else 'none'))
    tc = [r[0] for r in rows if 'UTCHMMA' in r[5]]
    print('# %d kernels, %d with tcgen05.mma (UTCHMMA)' % (len(rows), len(tc)))


if __name__ == '__main__':
    sys.exit(main())
