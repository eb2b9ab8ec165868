"""Build oracle/_ref/libref_voxel.so from the REFERENCE's own C++ (where it lies).

TEST INFRASTRUCTURE ONLY.  Runs only where /root/reference exists (the build
container); the GPU box uses the prebuilt .so that travels with the snapshot.
Two pieces of the reference compile stand-alone from a few files: its CPU voxelization
(libref_voxel.so) and the vendored spconv-1.x CPU rulebook generator, a header-only template
(libref_spconv.so, build_spconv below).  The scatter / MSDA / ingroup ops have no CPU path and the rest of
spconv/mmcv needs their own build systems; those rows are pinned by outputs of the reference's Python files
instead (tests/golden/make_golden_ref.py).
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


SPCONV_INC = '/root/reference/mmdet3d/ops/bevfusion-ops/spconv/include'
OUT_SPCONV = os.path.join(OUT_DIR, 'libref_spconv.so')


def build_spconv(force=False):
    """oracle/_ref/libref_spconv.so: the vendored spconv-1.x CPU rulebook generator (header-only templates in
    spconv/geometry.h) behind the C shim oracle/ref_spconv_wrap.cpp.  Plain g++, no torch."""
    hdr = os.path.join(SPCONV_INC, 'spconv', 'geometry.h')
    if not os.path.exists(hdr):
        return OUT_SPCONV if os.path.exists(OUT_SPCONV) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    wrap = os.path.join(HERE, 'ref_spconv_wrap.cpp')
    if (not force and os.path.exists(OUT_SPCONV)
            and os.path.getmtime(OUT_SPCONV) >= max(os.path.getmtime(wrap), os.path.getmtime(hdr))):
        return OUT_SPCONV
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')   # tensorview.h wants cuda_runtime_api.h
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-std=c++14', '-w', '-I', SPCONV_INC, '-I', cuda_inc, wrap,
                           '-o', OUT_SPCONV])
    return OUT_SPCONV


def build_all(force=False):
    return [build(force), build_spconv(force)]


if __name__ == '__main__':
    print(build_all(force='--force' in sys.argv))
