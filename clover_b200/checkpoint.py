"""Checkpoint I/O in the reference's on-disk format.

The reference saves ``{'meta': {...}, 'state_dict': OrderedDict, ['optimizer': ...]}`` through mmcv's ``save_checkpoint``
(core/runner/epoch_based_runner.py:25-58) and loads with ``load_checkpoint(model, path, strict=False)``
(swin_transformer_3d.py:210, tools/train.py), which unwraps ``state_dict``, strips a DataParallel ``module.`` prefix and
reports missing / unexpected keys instead of failing.  Parameter names, shapes and dtypes of the clover_b200 modules
equal the reference's (tests/test_modules_cpu.py), so the files are interchangeable in both directions; derived bf16
operand caches are keyed on parameter versions and refresh themselves after a load.
"""
import time
from collections import OrderedDict

import torch


def _unwrap(model):
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


def save_checkpoint(model, filename, optimizer=None, meta=None):
    """Reference layout: CPU tensors, no ``module.`` prefix, integer buffers included (the loaders drop
    relative_position_index / attn_mask themselves, swin_transformer_3d.py:146-154)."""
    sd = OrderedDict((k, v.detach().cpu()) for k, v in _unwrap(model).state_dict().items())
    ck = {"meta": dict(meta or {}, time=time.asctime()), "state_dict": sd}
    if optimizer is not None:
        ck["optimizer"] = optimizer.state_dict()
    torch.save(ck, filename)
    return ck


def load_checkpoint(model, filename, map_location="cpu", strict=False, revise_keys=((r"^module\.", ""),)):
    """Returns the checkpoint dict.  Keys are rewritten by ``revise_keys`` (regex, replacement) like mmcv does; the
    Swin integer buffers saved by older files are ignored; with strict=False mismatches are returned in
    ``checkpoint['missing_keys'] / ['unexpected_keys']`` instead of raising."""
    import re
    ck = torch.load(filename, map_location=map_location)
    if not isinstance(ck, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {filename}")
    sd = ck.get("state_dict", ck)
    out = OrderedDict()
    for k, v in sd.items():
        for pat, rep in revise_keys:
            k = re.sub(pat, rep, k)
        if "relative_position_index" in k or "attn_mask" in k:
            continue
        out[k] = v
    res = _unwrap(model).load_state_dict(out, strict=strict)
    ck["missing_keys"] = [k for k in res.missing_keys if "relative_position_index" not in k]
    ck["unexpected_keys"] = list(res.unexpected_keys)
    return ck
