"""Dry run of the round-2 GPU tool scripts on the CPU: their `main()` runs unmodified against the ABI emulator with the CUDA
timing calls stubbed (events -> wall clock), at tiny sizes.  A typo or a wrong call signature in a tool would otherwise cost
a GPU call to find.  What is checked is that each tool runs to its JSON line and that the line has the fields it promises —
never the numbers."""
import importlib.util
import json
import os
import sys
import time

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)


class _Event:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)

    def synchronize(self):
        pass


def _run_tool(name, argv, monkeypatch, capsys):
    import abi_emulator
    import isfusion_b200.head as H
    import isfusion_b200.img_neck as N
    import isfusion_b200.image_pipeline as IP
    import isfusion_b200._lib as L
    abi_emulator.install(monkeypatch)
    monkeypatch.setattr(H, '_conv', abi_emulator._conv_cpu)
    monkeypatch.setattr(N, '_conv', abi_emulator._conv_cpu)
    monkeypatch.setattr(IP, '_PLANS', {})
    monkeypatch.setattr(L, 'launch_count', lambda reset=False: 0)
    monkeypatch.setenv('ISF_TOOL_DEVICE', 'cpu')
    monkeypatch.setattr(torch.cuda, 'Event', _Event)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self, *a, **k: self)
    monkeypatch.setattr(sys, 'argv', [name] + argv)
    spec = importlib.util.spec_from_file_location('tool_' + name, os.path.join(ROOT, 'scripts', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith('{')]
    assert lines, 'the tool printed no JSON line'
    return json.loads(lines[-1])


def test_bench_full_forward_dry_run(monkeypatch, capsys):
    line = _run_tool('bench_full_forward', ['--steps', '1', '--warmup', '0', '--points', '3000', '--img', '64x96', '--frames', '1'],
                     monkeypatch, capsys)
    assert line['boxes'] == [200, 9] and set(line['split_ms']) == {'image_branch', 'neck', 'fpn_head_decode_d2h'}
    assert line['h2d_bytes_per_step'] == 3000 * 5 * 4 + 6 * 3 * 64 * 96 * 2 or line['h2d_bytes_per_step'] > 0


def test_bench_full_forward_raw_sensor_dry_run(monkeypatch, capsys):
    """--raw: uint8 camera frames + raw sweeps in, the input-contract kernels in front of the detector."""
    line = _run_tool('bench_full_forward', ['--steps', '1', '--warmup', '0', '--points', '3300', '--img', '64x96', '--frames', '1', '--raw',
                                            '--raw-hw', '90x160'], monkeypatch, capsys)
    assert line['boxes'] == [200, 9] and line['raw_sensor_input'] is True
    assert line['h2d_bytes_per_step'] == 3300 * 5 * 4 + 6 * 90 * 160 * 3


def test_bench_config5_dry_run(monkeypatch, capsys):
    line = _run_tool('bench_config5', ['--steps', '1', '--warmup', '0', '--points', '4000'], monkeypatch, capsys)
    assert line['feat_shapes'] == [[1, 128, 256, 256], [1, 256, 128, 128]]
    assert any('vfe' in k for k in line['stage_ms']) and any('sparse_encoder' in k for k in line['stage_ms'])
    assert isinstance(line['sparse_conv'], dict), line['sparse_conv']           # the per-launch re-issue + FLOP accounting ran
    assert line['sparse_conv']['launches'] == 21 and line['front_end']['voxels'] > 0


def test_bench_input_pipeline_dry_run(monkeypatch, capsys):
    line = _run_tool('bench_input_pipeline', ['--reps', '1', '--small'], monkeypatch, capsys)
    for key in ('merge_sweeps', 'merge_sweeps_close', 'image_aug_normalize_f16', 'image_aug_normalize_f32', 'image_pipeline_host_interface_ms'):
        assert key in line, key
    assert line['merge_sweeps']['points'] == 300 + 10 * 270



def test_bench_gemm_layers_dry_run(monkeypatch, capsys):
    """the per-layer A/B tool (default kernel, library options, tile orders): variants switch on and off, the
    bit-identity check and the table print.  (On the emulator every variant is the same arithmetic.)"""
    import abi_emulator
    import isfusion_b200._lib as L
    import isfusion_b200.spconv as sp
    abi_emulator.install(monkeypatch)
    monkeypatch.setenv('ISF_TOOL_DEVICE', 'cpu')
    monkeypatch.setattr(torch.cuda, 'Event', _Event)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(sys, 'argv', ['bench_gemm_layers', '--points', '2000', '--opt', 'gemm_mt=2', '--tile-order', 'none,16', '--reps', '1'])
    spec = importlib.util.spec_from_file_location('tool_gemm_layers', os.path.join(ROOT, 'scripts', 'bench_gemm_layers.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
    out = capsys.readouterr().out
    assert out.count('bit-identical to the first variant') == 3 and 'DIFFERS' not in out       # storage, ord16, gemm_mt=2
    assert 'total ms:' in out and sp.TILE_ORDER == 31
    assert L.lib().isf_set_option(b'gemm_mt', 3) == 0


def test_bench_hsf_bf16_dry_run(monkeypatch, capsys):
    """configs[2] on the bf16 entry points (isf_*_bf16): host plumbing of _lib.set_dtype through the HSF modules on the emulator."""
    import torch
    from isfusion_b200 import _lib as L
    try:
        line = _run_tool('bench_hsf', ['--steps', '1', '--warmup', '0', '--points', '3000', '--dtype', 'bf16'], monkeypatch, capsys)
    finally:
        L.set_dtype(torch.float16)
    assert line['dtype'] == 'bf16' and line['config']['region_levels'] == 1


def test_bench_hsf_dry_run(monkeypatch, capsys):
    line = _run_tool('bench_hsf', ['--steps', '1', '--warmup', '0', '--points', '3000', '--levels', '2'], monkeypatch, capsys)
    assert line['config']['region_levels'] == 2 and len(line['out_shapes']) == 2
    for key in ('a6 point-to-grid (cam2bev)', 'a7 conv_fusion', 'a8+a9 grid-to-region level 0', 'a8+a9 grid-to-region level 1'):
        assert key in line['stage_ms_eager'], key
    assert line['roofline']['bound'] == 'tensor' and 'point_to_grid' in line['roofline']
