"""Build libisfusion_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch dependency: plain `nvcc -shared`, static cudart.  Objects are cached by
mtime under is-fusion_b200/csrc/_build/ (git-ignored); the .so sits next to this
file so that it travels with the gpurun snapshot.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_build')
LIB = os.path.join(HERE, 'libisfusion_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _newest_header():
    hs = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(CSRC, '*.h')) + \
        glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return max([os.path.getmtime(h) for h in hs] + [0.0])


def _compile(src, obj, log):
    cmd = [NVCC] + FLAGS + ['-I', os.path.join(HERE, '..', 'include'), '-c', src, '-o', obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, 'w') as f:
        f.write(' '.join(cmd) + '\n' + r.stdout)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stdout[-6000:]))
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    hdr = _newest_header()
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr):
            jobs.append((s, o, o[:-2] + '.log'))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for f in [ex.submit(_compile, *j) for j in jobs]:
                o = f.result()
                if verbose:
                    print('compiled', o)
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs
        subprocess.check_call(cmd)
        if verbose:
            print('linked', LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
