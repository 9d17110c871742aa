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
                 check_finite=True, emit_bf16=True, reuse_grad_buffers=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        # reuse_grad_buffers: zero_grad() hands the gradient tensors of the finished step to the backward functions as
        # destinations for the next one (clover_b200.functional.stash_grad_sinks).  Under DistributedDataParallel with
        # gradient_as_bucket_view=True those tensors are the all-reduce bucket views, so the weight gradients are produced
        # in place and DDP's per-parameter copy into the bucket disappears.
        self.reuse_grad_buffers = bool(reuse_grad_buffers)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.check_finite, self.emit_bf16 = bool(check_finite), bool(emit_bf16)
        self.grad_scale = 1.0                    # multiply gradients by this (1 / loss scale) before clipping
        # Count of APPLIED updates, resident on the device (an overflow-skipped step does not advance it) and shared by every
        # parameter as state[p]['step'] -- the torch.optim.AdamW state layout (step, exp_avg, exp_avg_sq), so state_dict()
        # carries it and checkpoints are interchangeable with the reference's optimizer.
        self._step_dev = None
        self._static = None                      # (param ids, chunk tables) for the current set of parameters with gradients
        self._host = [None, None]                # double-buffered pinned descriptor tables + the events guarding their reuse
        self._host_ev = [None, None]
        self._calls = 0
        self.status = None                       # device fp32[3]: gradient norm, skipped flag, scratch

    @property
    def applied_steps(self):
        """Number of updates applied so far (one device -> host read)."""
        return 0 if self._step_dev is None else int(self._step_dev.item())

    def state_dict(self):
        """torch.optim.AdamW's layout: every parameter's state holds its OWN cpu scalar 'step' (the shared device counter is
        read once), so the dict loads into the reference's optimizer unchanged."""
        sd = super().state_dict()
        if self._step_dev is not None:
            step = float(self._step_dev.item())
            sd["state"] = {k: dict(st, step=torch.tensor(step, dtype=torch.float32)) if "step" in st else st
                           for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # re-link the shared device counter: every parameter was saved with the same 'step'
        steps = [st["step"] for st in self.state.values() if "step" in st]
        self._step_dev = None
        if steps:
            dev = next(iter(self.state.keys())).device
            self._step_dev = torch.as_tensor(steps[0], dtype=torch.float32).reshape(1).to(dev).clone()
            for st in self.state.values():
                st["step"] = self._step_dev

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
            for ev in self._host_ev:             # the old pinned tables may still be in flight
                if ev is not None:
                    ev.synchronize()
            self._host = [torch.empty(len(plist) * _DT.itemsize, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._host_ev = [None, None]
            self._static = (key, ct, co, torch.empty(len(plist) * _DT.itemsize, dtype=torch.uint8, device=device))
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
        ct, co, table = self._chunks(plist, dev)
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        rec = np.zeros(len(plist), dtype=_DT)
        dests = []
        for i, p in enumerate(plist):
            st = self.state[p]
            if not st:
                st["step"] = self._step_dev
                st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
            d = Fn.bf16_destination(p) if self.emit_bf16 else None
            dests.append(d)
            rec[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                      d[0] if d else 0, p.numel(), lrs[i], wds[i])
        # The pinned table is read by an asynchronous copy: never rewrite a buffer whose previous upload may still be in
        # flight (a loop without a per-step host sync lets the CPU run ahead).  Two buffers alternate, each guarded by the
        # event recorded after its copy.
        slot = self._calls & 1
        self._calls += 1
        if self._host_ev[slot] is not None:
            self._host_ev[slot].synchronize()
        host = self._host[slot]
        host.numpy()[:] = np.frombuffer(rec.tobytes(), dtype=np.uint8)
        table.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._host_ev[slot] = ev
        if self.status is None:
            self.status = torch.zeros(3, dtype=torch.float32, device=dev)
        lib = _lib.load()
        from . import ops
        ev = ops._prof_open()
        rc = lib.clv_adamw_step(C.c_void_p(table.data_ptr()), C.c_void_p(ct.data_ptr()), C.c_void_p(co.data_ptr()), ct.numel(), CHUNK,
                                float(betas[0]), float(betas[1]), float(eps), 0, float(self.grad_scale),
                                float(self.max_grad_norm), int(self.check_finite), C.c_void_p(self.status.data_ptr()),
                                C.c_void_p(self._step_dev.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "clv_adamw_step")
        # algorithmic bytes: grad read twice (norm pass + update), p / m / v read + written, bf16 copy written
        ops._prof_close(ev, "adamw", 0.0, float(sum(p.numel() for p in plist)) * (4 * 2 + 4 * 6 + 2))
        # the kernel wrote the parameters through raw pointers: bump their version counters (every derived cache keyed on
        # them re-validates) and re-stamp the bf16 copies that were refreshed in the same pass
        vers = tuple(p._version + 1 for p in plist)
        torch._C._autograd._unsafe_set_version_counter(tuple(plist), vers)
        for p, d in zip(plist, dests):
            if d:
                Fn.bf16_restamp(p, d)
        return loss

    def zero_grad(self, set_to_none=True):
        if self.reuse_grad_buffers:
            from . import functional as Fn
            Fn.stash_grad_sinks([(p, p.grad) for g in self.param_groups for p in g["params"]])
            set_to_none = True
        super().zero_grad(set_to_none=set_to_none)

    def grad_norm(self):
        """Gradient norm of the last step (after grad_scale) and whether that step was skipped -- one device->host read."""
        s = self.status.tolist()
        return s[0], bool(s[1])
