"""Build oracle/_ref/libref_voxel.so from the REFERENCE's own C++ (where it lies).

TEST INFRASTRUCTURE ONLY.  Runs only where /root/reference exists (the build
container); the GPU box uses the prebuilt .so that travels with the snapshot.
Only the reference's CPU voxelization compiles stand-alone: the
scatter / MSDA / ingroup ops have no CPU path and spconv/mmcv are not vendored
in buildable form, so those rows stay on the restated oracle.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/mmdet3d/ops/voxel/src/voxelization_cpu.cpp'
OUT_DIR = os.path.join(HERE, '_ref')
OUT = os.path.join(OUT_DIR, 'libref_voxel.so')


def build(force=False):
    if not os.path.exists(REF_SRC):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    wrap = os.path.join(HERE, 'ref_voxel_wrap.cpp')
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= max(os.path.getmtime(wrap), os.path.getmtime(REF_SRC))):
        return OUT
    import torch
    from torch.utils import cpp_extension as ce
    tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    cmd = ['g++', '-O2', '-shared', '-fPIC', '-std=c++17', '-w', wrap, REF_SRC]
    for inc in ce.include_paths():
        cmd += ['-I', inc]
    cmd += ['-I', sysconfig.get_paths()['include'], '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI),
            '-L', tlib, '-ltorch', '-ltorch_cpu', '-lc10', '-Wl,-rpath,' + tlib, '-o', OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
