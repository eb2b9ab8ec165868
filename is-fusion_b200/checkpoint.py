"""Checkpoint loading without mmcv (ref: mmcv.runner.load_checkpoint as tools/test.py:186 calls it on the released
IS-Fusion weights): accepts a raw state_dict or `{'state_dict': ..., 'meta': ...}`, strips the DataParallel `module.`
prefix, optional `revise_keys`, reports missing / unexpected keys.  The modules themselves accept both sparse-conv weight
layouts (spconv1 / mmcv (kD,kH,kW,Cin,Cout) and spconv2 (Cout,kD,kH,kW,Cin), write_spconv2.py:62-94)."""
import re

import torch


def load_checkpoint(model, filename, map_location='cpu', strict=False, logger=None, revise_keys=((r'^module\.', ''),)):
    ckpt = torch.load(filename, map_location=map_location, weights_only=False) if isinstance(filename, str) else filename
    if not isinstance(ckpt, dict):
        raise RuntimeError(f'No state_dict found in checkpoint file {filename}')
    state = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
    for pat, rep in revise_keys:
        state = {re.sub(pat, rep, k): v for k, v in state.items()}
    own = model.state_dict()
    unexpected = [k for k in state if k not in own]
    missing = [k for k in own if k not in state]
    mismatched = []
    for k in list(state):
        if k in own and tuple(state[k].shape) != tuple(own[k].shape):
            a, b = tuple(state[k].shape), tuple(own[k].shape)
            if not (len(a) == 5 and len(b) == 5 and a == (b[4], b[0], b[1], b[2], b[3])):     # spconv2 layout: permuted on load
                mismatched.append((k, a, b))
                del state[k]
    model.load_state_dict({k: v for k, v in state.items() if k in own}, strict=False)
    msg = []
    if unexpected:
        msg.append('unexpected key in source state_dict: ' + ', '.join(unexpected))
    if missing:
        msg.append('missing keys in source state_dict: ' + ', '.join(missing))
    if mismatched:
        msg.append('size mismatch: ' + ', '.join('%s %s vs %s' % m for m in mismatched))
    if msg:
        text = 'The model and loaded state dict do not match exactly\n' + '\n'.join(msg)
        if strict:
            raise RuntimeError(text)
        (logger.warning if logger is not None else print)(text)
    return ckpt
