"""Optional per-stage CUDA-event timers (off by default; bench.py --stages / scripts enable them)."""
import contextlib

import torch

ENABLED = False
_events = []


@contextlib.contextmanager
def stage(name):
    if not ENABLED:
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        yield
    finally:
        e1.record()
        _events.append((name, e0, e1))


def reset():
    _events.clear()


def summary():
    """{stage: total ms} of everything recorded since reset() (synchronises)."""
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _events:
        out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
    return out
