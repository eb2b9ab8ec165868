"""The C-ABI library loads and exports every symbol include/isfusion_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, 'include', 'isfusion_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = sorted(set(re.findall(r'\b(isf_\w+)\s*\(', hdr)))
    assert len(names) >= 10
    import isfusion_b200._lib as L
    lib = L.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.isf_version() >= 100


def test_errors_are_reported_without_gpu():
    import isfusion_b200._lib as L
    lib = L.lib()
    n = ctypes.c_size_t(0)
    assert lib.isf_dynamic_scatter_workspace_size(1000, ctypes.byref(n)) == 0 and n.value > 0
    rc = lib.isf_dynamic_scatter_workspace_size(-1, ctypes.byref(n))
    assert rc == 1 and b'invalid argument' in lib.isf_last_error()


def test_product_has_no_cpu_fallback():
    import pytest
    import torch
    from isfusion_b200 import _lib as L
    from isfusion_b200.voxel_layer import dynamic_voxelize
    pts = torch.rand(10, 5)
    coors = torch.zeros(10, 3, dtype=torch.int32)
    with pytest.raises(L.IsfError):
        dynamic_voxelize(pts, coors, [0.1, 0.1, 0.1], [0, 0, 0, 1, 1, 1])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'is-fusion_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py') or f.endswith('.cu') or f.endswith('.cuh'):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(d, f)


def test_only_the_checkers_import_the_oracle():
    """tier rule: oracle/ is test infrastructure — outside tests/ only __graft_entry__ (build of the checker, smoke(),
    make_pair for the parity tests) and bench.py's CPU legs (oracle_fps, reached by --impl reference / cpu_baseline) use it."""
    import ast
    allowed = {'__graft_entry__.py': {'build', 'make_pair', 'smoke'}, 'bench.py': {'oracle_fps'}}
    for d, dirs, files in os.walk(ROOT):
        dirs[:] = [x for x in dirs if x not in ('.git', 'tests', 'oracle', 'gpurun_out', '__pycache__', 'baseline')]
        for f in files:
            if not f.endswith('.py'):
                continue
            path = os.path.join(d, f)
            tree = ast.parse(open(path).read())
            for fn in ast.walk(tree):
                if isinstance(fn, (ast.FunctionDef, ast.Module)):
                    for node in (fn.body if isinstance(fn, ast.Module) else ast.walk(fn)):
                        hit = (isinstance(node, ast.ImportFrom) and (node.module or '').split('.')[0] == 'oracle') or \
                              (isinstance(node, ast.Import) and any(a.name.split('.')[0] == 'oracle' for a in node.names))
                        if hit:
                            name = fn.name if isinstance(fn, ast.FunctionDef) else '<module>'
                            assert name in allowed.get(os.path.relpath(path, ROOT), set()), (path, name)


def test_new_host_side_entry_points_validate_arguments():
    import isfusion_b200._lib as L
    lib = L.lib()
    n = ctypes.c_size_t(0)
    assert lib.isf_spconv_tile_order_workspace_size(300000, ctypes.byref(n)) == 0 and n.value > 300000 * 24
    assert lib.isf_spconv_tile_order_workspace_size(-1, ctypes.byref(n)) == 1
    for name, good, bad in (('tile_order_drop_bits', 11, 27), ('nbr_rows', 1, 2), ('gemm_mt', 3, 4), ('gemm_debug', 0, 8)):
        assert lib.isf_set_option(name.encode(), good) == 0, name
        assert lib.isf_set_option(name.encode(), bad) == 1 and b'unknown option or bad value' in lib.isf_last_error(), name
    for name, default in (('tile_order_drop_bits', 0), ('nbr_rows', 1), ('gemm_mt', 3), ('gemm_debug', 0)):
        assert lib.isf_set_option(name.encode(), default) == 0
    assert lib.isf_set_option(b'no_such_option', 1) == 1
    # the tile-order entry points refuse bad geometry before touching the device
    assert lib.isf_spconv_tile_order(None, ctypes.c_longlong(128), 27, None, 10, None, 31, None, None, None, ctypes.c_size_t(0), None) == 1


def test_emulator_signatures_match_the_header():
    """ctypes does not check argument counts, so a call site with a missing / extra argument would only show on the GPU.
    The CPU host-logic tests run the product against tests/abi_emulator.py, whose Python methods DO check arity — provided
    they have exactly the parameters the header declares.  This test ties the two together."""
    import inspect
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import abi_emulator
    hdr = open(os.path.join(ROOT, 'include', 'isfusion_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    decl = {}
    for m in re.finditer(r'\b(?:int|long long|unsigned long long|const char\*)\s+(isf_\w+)\s*\(([^;]*?)\)\s*;', hdr, flags=re.S):
        params = m.group(2).strip()
        decl[m.group(1)] = 0 if params in ('', 'void') else params.count(',') + 1
    emulated = [n for n in dir(abi_emulator.FakeLib) if n.startswith('isf_')]
    assert len(emulated) >= 35
    for name in emulated:
        assert name in decl, name + ' is emulated but not declared in the header'
        n_emul = len(inspect.signature(getattr(abi_emulator.FakeLib, name)).parameters) - 1      # minus self
        assert n_emul == decl[name], (name, n_emul, decl[name])


def test_emulator_pointer_contract_rejects_strided_views():
    """the emulator hands tensors where the product passes raw pointers: views a pointer cannot describe are refused, so a
    CPU run of the product code cannot pass with an argument the CUDA library would misread."""
    import pytest
    import torch
    import abi_emulator as E
    x = torch.zeros(6, 8)
    assert E._ptr_checked(x) is x and E._ptr_checked(x[:, :4]) is not None and E._ptr_checked(None) is None
    assert E._ptr_checked(x[2:5]) is not None                       # row slice: still one dense block
    with pytest.raises(AssertionError):
        E._ptr_checked(x.t())                                       # strided last dimension
    with pytest.raises(AssertionError):
        E._ptr_checked(torch.zeros(2, 3, 4, 8)[:, :, :2, :])        # 4-D view that is not row-strided
    with pytest.raises(Exception):
        E._require_contiguous(x[:, :4])


def test_backward_and_input_entry_points_validate_arguments():
    """argument checks of the entry points added for the backward kernels and the input contract run before any CUDA call,
    so they can be exercised here: bad arguments give ISF_ERR_ARG (1) and a message, never a crash."""
    import isfusion_b200._lib as L
    lib = L.lib()
    n = ctypes.c_size_t(0)
    assert lib.isf_dynamic_scatter_backward_workspace_size(1000, 64, ctypes.byref(n)) == 0 and n.value >= 1000 * 64 * 4
    assert lib.isf_dynamic_scatter_backward_workspace_size(-1, 64, ctypes.byref(n)) == 1
    assert lib.isf_dynamic_point_to_voxel_backward(None, None, None, None, None, None, 10, 5, 4, 3, None, ctypes.c_size_t(0), None) == 1   # reduce_type 3
    assert lib.isf_dynamic_point_to_voxel_backward(None, None, None, None, None, None, 0, 0, 4, 2, None, ctypes.c_size_t(0), None) == 0    # empty input: no-op
    assert lib.isf_dynamic_point_to_voxel_backward(None, None, None, None, None, None, 10, 5, 4, 2, None, ctypes.c_size_t(0), None) == 1   # null pointers
    assert lib.isf_ms_deform_attn_backward(None, 0, None, None, None, None, None, 1, 4, 2, 2, 1, 2, 2, None, None, None, None) == 1
    assert lib.isf_merge_sweeps_workspace_size(300000, ctypes.byref(n)) == 0 and n.value >= 2 * 300000 * 4
    cnt = ctypes.c_int(0)
    raw = (ctypes.c_float * 50)()
    out = (ctypes.c_float * 50)()
    one = (ctypes.c_double * 1)(0.0)
    # segments must tile [0, m] exactly; at most 24 segments; load_dim >= 5
    args = lambda start, nseg=1, load_dim=5: (raw, 10, load_dim, nseg, L.host_i32(start), None, None, one, L.host_i32([0] * max(nseg, 1)),
                                               ctypes.c_float(1.0), None, out, ctypes.byref(cnt))
    assert lib.isf_debug_merge_sweeps_host(*args([0, 10])) == 0 and cnt.value == 10
    assert lib.isf_debug_merge_sweeps_host(*args([0, 9])) == 1
    assert lib.isf_debug_merge_sweeps_host(*args([1, 10])) == 1
    assert lib.isf_debug_merge_sweeps_host(*args([0, 10], load_dim=4)) == 1
    assert lib.isf_debug_merge_sweeps_host(*args([0] * 26, nseg=25)) == 1
    # a transforming segment needs its calibration
    assert lib.isf_debug_merge_sweeps_host(raw, 10, 5, 1, L.host_i32([0, 10]), None, None, one, L.host_i32([1]), ctypes.c_float(1.0), None, out,
                                           ctypes.byref(cnt)) == 1


def test_bf16_entry_points_and_dtype_switch():
    """the `_bf16` twins are exported, resolve through _lib.fn16 by tensor type, validate their arguments without a GPU, and the
    library's options for the round-2 schedules accept / reject values as documented."""
    import pytest
    import torch
    from isfusion_b200 import _lib as L
    lib = L.lib()
    for name in ('isf_gemm_bf16', 'isf_conv3x3_nhwc_bf16', 'isf_window_attention_bf16', 'isf_cam2bev_bf16'):
        assert hasattr(lib, name), name
    assert L.fn16('isf_gemm_f16', torch.float16).__name__ == 'isf_gemm_f16'
    assert L.fn16('isf_gemm_f16', torch.bfloat16).__name__ == 'isf_gemm_bf16'
    assert L.fn16('isf_cam2bev', torch.bfloat16).__name__ == 'isf_cam2bev_bf16'
    with pytest.raises(L.IsfError):
        L.fn16('isf_gemm_f16', torch.float32)
    with pytest.raises(L.IsfError):
        L.set_dtype(torch.float32)
    assert L.DTYPE == torch.float16
    L.set_dtype(torch.bfloat16)
    try:
        assert L.DTYPE == torch.bfloat16
    finally:
        L.set_dtype(torch.float16)
    # argument validation happens before any CUDA call (k not a multiple of 8; misaligned leading dimension)
    assert lib.isf_gemm_bf16(None, None, 4, 12, 64, 64, None, None, None, 0, 0, None, None, ctypes.c_float(0), None, None, 64, None, None,
                             None) != 0
    for name, good, bad in (('gemm_pair', (0, 1, 64, 128, 256), (2, 512)), ('gemm_mma2', (0, 1), (2,)), ('gemm_wpc', (0, 1, 2), (3,)),
                            ('gemm_v4_32', (0, 1, 2), (3,)), ('i2s_tc', (0, 1), (2,)), ('gemm_rev', (0, 1), (2,))):
        for v in bad:
            assert lib.isf_set_option(name.encode(), v) != 0, (name, v)
        for v in good:
            assert lib.isf_set_option(name.encode(), v) == 0, (name, v)
    # back to the shipped defaults
    for name, v in (('gemm_pair', 128), ('gemm_mma2', 1), ('gemm_wpc', 0), ('gemm_v4_32', 1), ('i2s_tc', 0), ('gemm_rev', 0)):
        assert lib.isf_set_option(name.encode(), v) == 0
