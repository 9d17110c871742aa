"""Optimizer step of the Clover training path on one multi-tensor kernel.

``FusedAdamW`` is a ``torch.optim.Optimizer`` (same param-group keys as ``torch.optim.AdamW``: lr, betas, eps,
weight_decay) whose ``step()`` is ONE launch of ``clv_adamw_step`` over every parameter: AdamW on the fp32 master weights,
optional gradient unscale (``grad_scale`` = 1 / loss scale), global-norm clipping (``max_grad_norm`` -- the reference's
``optimizer_config.grad_clip.max_norm`` 15 / 5 / 50), device-side skip of a step whose gradient norm is not finite
(core/hooks/mmcv_Fp16OptimizerHook.py:96-149) and the refresh of the bf16 operand copies the GEMMs read
(``functional.w16``), so no cast kernels run after the step.

``param_groups_from_cfg`` reproduces mmcv's DefaultOptimizerConstructor for the shipped ``paramwise_cfg``
(configs/exp_local/pretrain_webvid_cc3m.py:129-134, finetune_msrvttQA.py:90-97): no weight decay on norm layers, biases and
the listed custom keys, ``lr_mult`` per custom key (qa_head x10).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import functional as Fn

CHUNK = 16384
_DT = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("w16", "<u8"), ("n", "<i8"), ("lr", "<f4"), ("wd", "<f4")])
assert _DT.itemsize == 56


def param_groups_from_cfg(model, base_lr, weight_decay, paramwise_cfg=None):
    """mmcv DefaultOptimizerConstructor.add_params semantics for the keys Clover's configs use: custom_keys (longest
    matching substring wins; lr_mult / decay_mult), then bias_decay_mult for '*.bias', norm_decay_mult for parameters of
    normalisation layers.  Returns one param group per parameter (as mmcv does)."""
    cfg = dict(paramwise_cfg or {})
    custom = cfg.get("custom_keys", {})
    keys = sorted(sorted(custom.keys()), key=len, reverse=True)
    bias_mult, norm_mult = cfg.get("bias_decay_mult", 1.0), cfg.get("norm_decay_mult", 1.0)
    norm_types = (torch.nn.LayerNorm, torch.nn.GroupNorm, torch.nn.modules.batchnorm._BatchNorm)
    norm_params = {id(p) for m in model.modules() if isinstance(m, norm_types) for p in m.parameters(recurse=False)}
    groups = []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        g = {"params": [p], "lr": base_lr, "weight_decay": weight_decay, "name": name}
        hit = next((k for k in keys if k in name), None)
        if hit is not None:
            g["lr"] = base_lr * custom[hit].get("lr_mult", 1.0)
            g["weight_decay"] = weight_decay * custom[hit].get("decay_mult", 1.0)
        elif id(p) in norm_params:
            g["weight_decay"] = weight_decay * norm_mult
        elif name.endswith(".bias"):
            g["weight_decay"] = weight_decay * bias_mult
        groups.append(g)
    return groups


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None,
                 check_finite=True, emit_bf16=True):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.check_finite, self.emit_bf16 = bool(check_finite), bool(emit_bf16)
        self.grad_scale = 1.0                    # multiply gradients by this (1 / loss scale) before clipping
        self._step = 0
        self._static = None                      # (param ids, chunk tables) for the current set of parameters with gradients
        self.status = None                       # device fp32[3]: gradient norm, skipped flag, scratch

    # ---- tables ----------------------------------------------------------------------------------------------------
    def _chunks(self, plist, device):
        key = tuple(id(p) for p in plist)
        if self._static is None or self._static[0] != key:
            ct, co = [], []
            for i, p in enumerate(plist):
                offs = np.arange(0, p.numel(), CHUNK, dtype=np.int64)
                ct.append(np.full(len(offs), i, dtype=np.int32))
                co.append(offs)
            ct = torch.from_numpy(np.concatenate(ct)).to(device)
            co = torch.from_numpy(np.concatenate(co)).to(device)
            host = torch.empty(len(plist) * _DT.itemsize, dtype=torch.uint8).pin_memory()
            self._static = (key, ct, co, host, torch.empty(len(plist) * _DT.itemsize, dtype=torch.uint8, device=device))
        return self._static[1:]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        plist, lrs, wds = [], [], []
        betas = eps = None
        for g in self.param_groups:
            if betas is None:
                betas, eps = g["betas"], g["eps"]
            elif (g["betas"], g["eps"]) != (betas, eps):
                raise ValueError("FusedAdamW: betas / eps must be the same in every param group")
            for p in g["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() \
                        or not p.grad.is_contiguous():
                    raise TypeError("FusedAdamW: contiguous fp32 CUDA parameters and gradients only (no CPU path)")
                plist.append(p); lrs.append(g["lr"]); wds.append(g["weight_decay"])
        if not plist:
            return loss
        dev = plist[0].device
        ct, co, host, table = self._chunks(plist, dev)
        rec = np.zeros(len(plist), dtype=_DT)
        dests = []
        for i, p in enumerate(plist):
            st = self.state[p]
            if not st:
                st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
            d = Fn.bf16_destination(p) if self.emit_bf16 else None
            dests.append(d)
            rec[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                      d[0] if d else 0, p.numel(), lrs[i], wds[i])
        host.numpy()[:] = np.frombuffer(rec.tobytes(), dtype=np.uint8)
        table.copy_(host, non_blocking=True)
        if self.status is None:
            self.status = torch.zeros(3, dtype=torch.float32, device=dev)
        self._step += 1
        lib = _lib.load()
        rc = lib.clv_adamw_step(C.c_void_p(table.data_ptr()), C.c_void_p(ct.data_ptr()), C.c_void_p(co.data_ptr()), ct.numel(), CHUNK,
                                float(betas[0]), float(betas[1]), float(eps), self._step, float(self.grad_scale),
                                float(self.max_grad_norm), int(self.check_finite), C.c_void_p(self.status.data_ptr()),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "clv_adamw_step")
        # the kernel wrote the parameters through raw pointers: bump their version counters (every derived cache keyed on
        # them re-validates) and re-stamp the bf16 copies that were refreshed in the same pass
        vers = tuple(p._version + 1 for p in plist)
        torch._C._autograd._unsafe_set_version_counter(tuple(plist), vers)
        for p, d in zip(plist, dests):
            if d:
                Fn.bf16_restamp(p, d)
        return loss

    def grad_norm(self):
        """Gradient norm of the last step (after grad_scale) and whether that step was skipped -- one device->host read."""
        s = self.status.tolist()
        return s[0], bool(s[1])
