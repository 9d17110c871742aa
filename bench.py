#!/usr/bin/env python
"""Headline benchmark: Clover pre-training clips/sec (BASELINE.json metric, config c3).

    python bench.py --gpus N --steps K --warmup W            # clover_b200 arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement of the reference
                                                               (oracle/, the reference is pure Python and cannot
                                                               travel to the GPU box) on the host cores

A step = one pass of the hot path over one synthetic batch: CloverPretrain forward (2x Video Swin-B,
2x BERT-base, 2x 3-layer fusion, heads, MLM focal + tri-modal NCE/ranking losses) + backward +
data-parallel gradient all-reduce (N > 1) + AdamW update of the fp32 master weights.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "clips/s"
SWIN_B = dict(embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32), img_in=1024)
# BASELINE.json configs that are whole training steps (c1 is the CPU case = the reference arm's sample, c2 the window-attention
# microbenchmark reported inside the c3 line).  flop = algorithmic FLOP per clip, forward + backward (SURVEY.md 8(d) / BASELINE.md 4).
WORKLOADS = {
    "c3": dict(metric="clover_pretrain_clips_per_sec", frames=8, L=32, clips=64, flop=885e9, grad_clip=15.0, wd=0.005,
               text="c3: CloverPretrain step, Video Swin-B + BERT-base text + 3-layer fusion, tri-modal NCE/ranking + MLM focal, "
                    "fwd+bwd+grad-allreduce+AdamW"),
    "c4": dict(metric="clover_finetune_retrieval_clips_per_sec", frames=16, L=40, clips=16, flop=3 * (281.3e9 + 6.9e9), grad_clip=5.0,
               wd=0.001, task="retrieval",
               text="c4: CloverFinetune(task='retrieval') step, Video Swin-B on 16x224x224 clips (N = 392 windows) + BERT-base on "
                    "40-token queries, NormSoftmaxLoss over the all-gathered batch, fwd+bwd+grad-allreduce+AdamW"),
    "c5": dict(metric="clover_finetune_videoqa_clips_per_sec", frames=16, L=40, clips=16, flop=3 * (281.3e9 + 6.9e9 + 20.1e9),
               grad_clip=50.0, wd=0.001, task="video_qa",
               text="c5: CloverFinetune(task='video_qa') step, Video Swin-B on 16x224x224 clips + BERT-base + 3-layer fusion over "
                    "432 tokens + QA_OE_Head(1500) + CE, fwd+bwd+grad-allreduce+AdamW"),
}
METRIC = WORKLOADS["c3"]["metric"]
FLOP_PER_CLIP_FWD_BWD = WORKLOADS["c3"]["flop"]


def model_cfg(regularisers=True, workload="c3"):
    from clover_b200.configs import SHIPPED_REGULARISERS, finetune_cfg, pretrain_cfg
    if workload == "c3":
        reg = SHIPPED_REGULARISERS if regularisers else {}
        return pretrain_cfg(SWIN_B["embed"], SWIN_B["depths"], SWIN_B["heads"], SWIN_B["img_in"], 768, 30522, 12, 3, 4, **reg)
    # fine-tune shapes: mm_backbone.num_frames must cover T = 16 / 2
    drop = 0.1 if regularisers else 0.0
    cfg = finetune_cfg(WORKLOADS[workload]["task"], frames_half=8, bert_dropout=drop, qa_dropout=0.5 if regularisers else 0.0)
    cfg["backbone"]["drop_path_rate"] = 0.3 if regularisers else 0.0
    return cfg


def workload_config(clips, n_gpus, regularisers=True, workload="c3"):
    from clover_b200.configs import SHIPPED_REGULARISERS
    w = WORKLOADS[workload]
    reg = SHIPPED_REGULARISERS if regularisers else dict(bert_dropout=0.0, drop_path_rate=0.0, t_head_dropout=0.0)
    mb = clips * 3 * w["frames"] * 224 * 224 * 4 / 1e6
    return {
        "workload": w["text"],
        "clips_per_gpu": clips, "frames": w["frames"], "resolution": 224, "caption_tokens": w["L"], "global_batch": clips * n_gpus,
        "parallelism": f"dp{n_gpus}", "dropout": reg["bert_dropout"], "drop_path": reg["drop_path_rate"],
        "text_head_dropout": reg["t_head_dropout"] if workload == "c3" else None,
        "regularisers": "shipped rates of the reference config (training mode)" if regularisers else
                        "all zero (the parity configuration; --parity-config)",
        "optimizer": f"clover_b200.optim.FusedAdamW (one multi-tensor kernel: AdamW on fp32 masters + grad-norm clip {w['grad_clip']:g} + "
                     "finite check + bf16 weight refresh), paramwise weight decay of the shipped config, inside the timed region",
        "grad_allreduce": "DDP buckets, bf16-compressed (the reference all-reduces half-precision gradients)" if n_gpus > 1 else "none (1 GPU)",
        "l2_policy": f"per-step inputs ({mb:.0f} MB of clips) and activations (tens of GB) far exceed the 126 MB L2",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1381.1), d.get("bf16_tflops", 1639.5), d.get("hbm_gbs", 6545.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, clips=2, seed=1000, regularisers=True):
    """The reference's algorithm (oracle port, fp32, plain PyTorch CPU ops) on the host cores: fwd + bwd + grad-norm clip +
    AdamW (torch.optim, the reference's optimizer) of the same pre-train step on a bounded sample of `clips` clips per step.
    Returns (clips_per_sec, ms_per_step, cores)."""
    import torch
    from clover_b200.synthetic import make_batch, synth_state_dict
    from oracle import clover_oracle as O
    from oracle.state_shapes import pretrain_shapes
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shapes = pretrain_shapes(SWIN_B["embed"], list(SWIN_B["depths"]), list(SWIN_B["heads"]), SWIN_B["img_in"], 768, 3072,
                             30522, 512, 12, 3, 4)
    state = {k: v.requires_grad_(True) for k, v in synth_state_dict(shapes, 7).items()}
    cfg = dict(depths=list(SWIN_B["depths"]), num_heads=list(SWIN_B["heads"]), text_layers=12, fusion_layers=3, bert_heads=12,
               vocab=30522)
    params = list(state.values())
    opt = torch.optim.AdamW(params, lr=1e-7, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.005)
    times = []
    for it in range(warmup + steps):
        batch = make_batch(clips, frames=8, L=32, seed=seed + it)
        t0 = time.perf_counter()
        drop, dps = O.random_regularisers(cfg["depths"], clips, 0.3, 0.1, seed + it) if regularisers else (None, None)
        losses, _ = O.pretrain_forward(state, batch, cfg, drop=drop, drop_paths=dps)
        O.total_loss(losses).backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 15.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return clips / (ms / 1e3), ms, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload != "c3":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm times the c3 (pre-train) step only"}), flush=True)
        return
    clips = 2
    value, ms, cores = cpu_reference_run(args.steps, args.warmup, clips, regularisers=not args.parity_config)
    sample = (f"{clips} clips per step of the c3 step (Swin-B + BERT-base + fusion, 8x224x224, L=32), fp32, "
              "fwd + bwd + grad-norm clip + torch.optim.AdamW; ONE process on all host cores whatever --gpus is")
    # `config` names the WORKLOAD both arms are quoted on (c3, 64 clips per GPU); what this arm actually runs per step is a
    # bounded sample of it -- stated in config.reference_sample and cpu_baseline.sample, never implied to be 64 clips
    config = dict(workload_config(64, args.gpus, not args.parity_config),
                  reference_sample={"clips_per_step": clips, "processes": 1, "optimizer_in_step": True, "dtype": "f32",
                                    "note": "clips/s = clips_per_step / seconds per step of this sample"})
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(dev, regularisers, steps=2, try_clips=(32, 16, 8, 4)):
    """The like-for-like GPU baseline SURVEY.md 0 / 8(d) names: the reference's algorithm in plain PyTorch eager ON THIS B200
    (the oracle restatement on cuda = ATen / cuBLAS / cuDNN library kernels, what the reference itself launches), the same c3
    step (fwd + bwd + clip + torch.optim.AdamW(fused)), in fp32 and under torch.autocast(bfloat16).  Eager materialises every
    attention score tensor, so 64 clips do not fit in 180 GB: the largest batch of `try_clips` that fits is used and stated.
    Test-infrastructure code (oracle/) timed as a BASELINE -- never on the product path."""
    import torch
    from clover_b200.synthetic import make_batch, synth_state_dict
    from oracle import clover_oracle as O
    from oracle.state_shapes import pretrain_shapes
    shapes = pretrain_shapes(SWIN_B["embed"], list(SWIN_B["depths"]), list(SWIN_B["heads"]), SWIN_B["img_in"], 768, 3072,
                             30522, 512, 12, 3, 4)
    cfg = dict(depths=list(SWIN_B["depths"]), num_heads=list(SWIN_B["heads"]), text_layers=12, fusion_layers=3, bert_heads=12,
               vocab=30522)
    out = {}
    for mode in ("bf16_autocast", "fp32"):
        res = None
        for clips in try_clips:
            state = opt = None
            try:
                torch.manual_seed(0)
                state = {k: v.to(dev).requires_grad_(True) for k, v in synth_state_dict(shapes, 7).items()}
                params = list(state.values())
                opt = torch.optim.AdamW(params, lr=1e-7, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.005, fused=True)
                batch = {k: v.to(dev) for k, v in make_batch(clips, frames=8, L=32, seed=1000).items()}

                def one(it):
                    drop = dps = None
                    if regularisers:
                        g = torch.Generator(device=dev).manual_seed(it)
                        rates = torch.linspace(0, 0.3, sum(cfg["depths"])).tolist()
                        drop = lambda kind, x: x * (torch.rand(x.shape, generator=g, device=dev) >= 0.1).to(x.dtype) / 0.9
                        dps = lambda: [None if r == 0 else tuple((torch.rand(clips, generator=g, device=dev) < 1.0 - r).float() / (1.0 - r)
                                                                  for _ in range(2)) for r in rates]
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
                        losses, _ = O.pretrain_forward(state, batch, cfg, drop=drop, drop_paths=dps)
                        loss = O.total_loss(losses)
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 15.0)
                    opt.step()
                    opt.zero_grad(set_to_none=True)
                    return loss
                one(0)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for it in range(steps):
                    loss = one(1 + it)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                res = {"value": clips / ms * 1e3, "unit": UNIT, "clips_per_step": clips, "ms_per_step": ms, "loss": float(loss),
                       "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
            except torch.OutOfMemoryError:
                res = None
            finally:
                state = opt = None
                import gc
                gc.collect()
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
            if res is not None:
                break
        out[mode] = res
    out["what"] = ("reference algorithm in PyTorch eager on the same B200 (oracle restatement on cuda: ATen/cuBLAS library kernels), "
                   "c3 step fwd+bwd+clip+AdamW(fused), largest batch that fits; a baseline, not the product")
    return out


# ------------------------------------------------------------------------------------------------
def window_attention_c2(iters=5):
    """BASELINE config c2: WindowAttention3D core, Swin-B stage-1 width (C = 128, 4 heads x 32), window (8,7,7) -> N = 392,
    64 clip-equivalents of 8x56x56 tokens (4096 windows), shifted and unshifted, bf16 forward + backward (incl. the
    bias-table gradient).  Algorithmic FLOPs of SURVEY.md 8(d): 4 T N C forward, 8 T N C backward."""
    import torch
    from clover_b200 import ops, swin, tables
    dev = torch.device("cuda", torch.cuda.current_device())
    dims, heads, hd, clips = (8, 56, 56), 4, 32, 64
    out_rows = {}
    for shifted in (0, 1):
        win, sh = tables.get_window_size(dims, (8, 7, 7), (4, 3, 3) if shifted else (0, 0, 0))
        N = win[0] * win[1] * win[2]
        nwin = (dims[0] // win[0]) * (dims[1] // win[1]) * (dims[2] // win[2])
        batch = clips * nwin
        g = torch.Generator(device=dev).manual_seed(1)
        qkv = (torch.randn(batch * N, 3 * heads * hd, generator=g, device=dev) * 0.7).bfloat16()
        dout = torch.randn(batch * N, heads * hd, generator=g, device=dev).bfloat16()
        table = torch.randn(2535, heads, generator=g, device=dev) * 0.5
        code, off = tables.rel_code(N, (8, 7, 7))
        code = torch.from_numpy(code).to(dev)
        masked = any(x > 0 for x in sh)
        region = torch.from_numpy(tables.region_ids(*dims, win, sh)).to(dev) if masked else None
        kw = dict(bias_table=table, rel_code=code, code_off=off, region=region, w7=swin._w7_spec(dims, win, sh, (8, 7, 7), dev))
        out = torch.empty(batch * N, heads * hd, dtype=torch.bfloat16, device=dev)
        lse = torch.empty(batch, heads, N, dtype=torch.float32, device=dev)
        dqkv = torch.empty_like(qkv)
        dtab = torch.zeros(2535, heads, device=dev)

        def both():
            ops.attention_fwd(qkv, batch, N, heads, hd, out, lse, **kw)
            ops.attention_bwd(qkv, out, dout, lse, batch, N, heads, hd, dqkv, hd ** -0.5, dbias_table=dtab, **kw)
        for _ in range(3):
            both()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            both()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out_rows["shifted" if shifted else "unshifted"] = {"ms_fwd_bwd": round(ms, 4),
                                                          "tflops": round(12.0 * batch * heads * N * N * hd / ms / 1e9, 1)}
    return out_rows


def run_ours(args):
    import torch
    import torch.distributed as dist
    from clover_b200 import _lib, ops, registry
    from clover_b200.synthetic import make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    registry.register_all()
    torch.manual_seed(0)
    wl = args.workload
    W = WORKLOADS[wl]
    METRIC = W["metric"]
    model = registry.build_model(model_cfg(not args.parity_config, wl)).to(dev)
    model.train()
    # parameters the reference never gives a gradient (text pooler, fusion's unused bert_embedding; the whole fusion encoder
    # in the retrieval fine-tune): keep DDP static
    for n, p in model.named_parameters():
        if ".pooler." in n or ".bert_embedding." in n or (wl == "c4" and n.startswith("multimodal_backbone.")):
            p.requires_grad_(False)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False,
                                                        gradient_as_bucket_view=True, bucket_cap_mb=args.bucket_mb,
                                                        static_graph=args.static_graph)
        if not args.fp32_allreduce:
            # the reference all-reduces HALF-precision gradients (model.half() + allreduce_grads, core/hooks/
            # mmcv_Fp16OptimizerHook.py:120-122): 0.55 GB per step over NVLink instead of 1.1 GB; masters / Adam stay fp32
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            net.register_comm_hook(None, default_hooks.bf16_compress_hook)
    from clover_b200.optim import FusedAdamW, param_groups_from_cfg
    if wl == "c3":       # configs/exp_local/pretrain_webvid_cc3m.py:129-137
        paramwise = dict(norm_decay_mult=0.0, bias_decay_mult=0.0, custom_keys={"absolute_pos_embed": dict(decay_mult=0.0),
                                                                                 "relative_position_bias_table": dict(decay_mult=0.0)})
    else:                # configs/exp_local/finetune_msrvttQA.py:90-97 / finetune_msrvtt_retrieval.py
        paramwise = dict(norm_decay_mult=0.0, bias_decay_mult=0.0, custom_keys={"qa_head": dict(lr_mult=10)})
    opt = FusedAdamW(param_groups_from_cfg(model, 1e-7, W["wd"], paramwise), betas=(0.9, 0.98), eps=1e-8, max_grad_norm=W["grad_clip"],
                     reuse_grad_buffers=not args.no_grad_sinks)
    clips = args.clips if args.clips > 0 else W["clips"]
    if wl == "c3":
        keys = ("imgs", "label", "token_ids", "segment_ids", "input_mask", "mlm_label", "v_token_mask")
        raw = make_batch(clips, frames=8, L=32, seed=1000 + rank)
    else:
        from clover_b200.synthetic import make_finetune_batch
        keys = ("imgs", "label", "token_ids", "segment_ids", "input_mask")
        raw = make_finetune_batch(W["task"], clips, frames=16, size=224, L=40, seed=1000 + rank)
    host = {k: v.pin_memory() for k, v in raw.items()}
    devb = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    h2d = sum(host[k].numel() * host[k].element_size() for k in keys)
    last_losses = {}

    def step(batch):
        kw = {k: batch[k] for k in keys[2:]}
        losses = net(batch["imgs"], batch["label"], return_loss=True, **kw)
        last_losses.clear()
        last_losses.update({k: v.detach() for k, v in losses.items()})
        loss = sum(v for k, v in losses.items() if "loss" in k)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(devb)
    sync()
    # ---- timed region 1: inputs resident in HBM ----------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    prof_raw = []
    for i in range(args.steps):
        if i == 0:
            ops.profile_begin()          # per-launch CUDA events on the FIRST timed step only (they cost ~3 % of a step)
        loss = step(devb)
        if i == 0:
            prof_raw = ops.profile_detach()
    e1.record()
    sync()
    ms_total = e0.elapsed_time(e1)
    prof = ops.profile_resolve(prof_raw)
    prof_steps = 1
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = world * clips / (ms_step / 1e3)
    # ---- multi-rank numerical check: the alignment losses are evaluated on the ALL-GATHERED embeddings, so every rank must
    # hold the identical value (models/utils/gather_loss.py:47-62); the per-rank MLM / QA losses legitimately differ
    consistency = None
    if world > 1:
        names = sorted(k for k in last_losses if k in ("nce_loss", "rank_t_tm_loss", "v_nce_loss", "rank_v_vm_loss", "retrieval_nce_loss"))
        if names:
            mine = torch.stack([last_losses[k].float().reshape(()) for k in names])
            allv = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            allv = torch.stack(allv)
            spread = float(((allv.max(0).values - allv.min(0).values) / allv.abs().max(0).values.clamp_min(1e-12)).max())
            consistency = {"global_loss_entries": names, "max_relative_spread_over_ranks": spread}
            assert spread <= 1e-5, f"alignment losses differ across ranks: {allv.tolist()}"
    # ---- timed region 2: end to end through the public API with host buffers ----------------------
    # Every step copies ITS inputs from pinned host memory (on a copy stream, one step ahead of the compute stream, like any
    # input pipeline) and reads ITS loss back to the host (asynchronous copy into pinned memory, consumed one step later, like
    # a logger): K copies in, K reads out, all inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    loss_pinned = torch.zeros(2, dtype=torch.float32).pin_memory()
    read_events = [torch.cuda.Event(), torch.cuda.Event()]

    def run_e2e(host_batch):
        def stage():
            with torch.cuda.stream(copy_stream):
                b = {k: v.to(dev, non_blocking=True) for k, v in host_batch.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return b, ev

        sync()
        e0.record()
        nxt = stage()
        for i in range(args.steps):
            b, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            if i + 1 < args.steps:
                nxt = stage()
            loss = step(b)
            for t in b.values():
                t.record_stream(torch.cuda.current_stream())
            loss_pinned[i % 2:i % 2 + 1].copy_(loss.detach().float().reshape(1), non_blocking=True)   # device -> host read of the result
            read_events[i % 2].record()
            if i > 0:
                read_events[(i - 1) % 2].synchronize()
        read_events[(args.steps - 1) % 2].synchronize()
        last = loss_pinned[(args.steps - 1) % 2].clone()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * clips / (float(t) / args.steps / 1e3), last

    if args.trace:                   # developer aid: kernel timeline of two more steps (compute + NCCL streams) per rank
        from torch.profiler import profile, ProfilerActivity
        sync()
        with profile(activities=[ProfilerActivity.CUDA]) as tp:
            for _ in range(2):
                step(devb)
            sync()
        os.makedirs(args.trace, exist_ok=True)
        tp.export_chrome_trace(os.path.join(args.trace, f"trace_rank{rank}.json.gz"))
    if args.kernels_only:            # profiler captures (ncu): only the device-timed region above matters
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "ms_per_step": ms_step,
                              "note": "--kernels-only run (for ncu): no e2e / roofline / baseline legs; not a bench line"}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    e2e_value, loss_host = run_e2e(host)
    # ---- timed region 3 (extra): the same loop fed with RAW uint8 frames, normalised inside the patch gather -- the reference's
    # GPUNormalize module hook (utils/module_hooks.py:35-87): a quarter of the bytes cross PCIe
    host8 = dict(host)
    g8 = torch.Generator().manual_seed(2000 + rank)
    host8["imgs"] = torch.randint(0, 256, tuple(host["imgs"].shape), dtype=torch.uint8, generator=g8).pin_memory()
    model.backbone.set_input_normalization([123.675, 116.28, 103.53], [58.395, 57.12, 57.375])
    h2d8 = sum(host8[k].numel() * host8[k].element_size() for k in keys)
    step({k: v.to(dev) for k, v in host8.items()})
    e2e8_value, _ = run_e2e(host8)

    if rank == 0:
        sus, burst, hbm, src = measured_peaks()
        fam = {}
        for f, fl, by, ms in prof:
            a = fam.setdefault(f, [0.0, 0.0, 0.0, 0])
            a[0] += fl; a[1] += by; a[2] += ms; a[3] += 1
        fam_total = sum(v[2] for v in fam.values()) / prof_steps

        def pre(prefix, i):
            return sum(v[i] for k, v in fam.items() if k.startswith(prefix))
        if os.environ.get("CLOVER_B200_PROFILE_SHAPES", "0") == "1":
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            rows = sorted(((k, v[3] // prof_steps, v[2] / prof_steps, v[0] / max(v[2], 1e-9) / 1e9, v[1] / max(v[2], 1e-9) / 1e6)
                           for k, v in fam.items()), key=lambda r: -r[2])
            with open(os.path.join(ROOT, "gpurun_out", "family_times.txt"), "w") as f:
                f.write("family | launches/step | ms/step | TFLOP/s | GB/s\n")
                for r in rows:
                    f.write(f"{r[0]} | {r[1]} | {r[2]:.3f} | {r[3]:.1f} | {r[4]:.0f}\n")
            g = [sum(v[i] for k, v in fam.items() if k.startswith("gemm")) for i in range(4)]
        else:
            g = fam.get("gemm", [0.0, 0.0, 1e-9, 1])
        gemm_tflops = g[0] / (g[2] * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)", "achieved": gemm_tflops, "peak": sus,
                    "unit": "TFLOP/s", "frac": gemm_tflops / sus, "traffic": traffic,
                    "traffic_source": "profiles/gemm_traffic.json: ncu dram__bytes_read + write per GEMM launch of this command (r02z launch list), not re-measured in this run",
                    "peak_source": f"{src} bf16_tflops_sustained",
                    "launches_timed": g[3], "avg_launch_ms": g[2] / max(1, g[3]),
                    "share_of_step": g[2] / (ms_step * prof_steps),
                    "launch_timing": "CUDA events around every launch of the first timed step",
                    "families_ms_per_step": {k: round(pre(k, 2) / prof_steps, 3) for k in sorted({n.split(" ")[0] for n in fam})},
                    "hbm_bound_families": {k: {"ms_per_step": round(pre(k, 2) / prof_steps, 3),
                                                "algorithmic_gbs": round(pre(k, 1) / max(1e-9, pre(k, 2)) / 1e6, 1),
                                                "frac_of_measured_hbm": round(pre(k, 1) / max(1e-9, pre(k, 2)) / 1e6 / hbm, 3)}
                                            for k in sorted({n.split(" ")[0] for n in fam}) if pre(k, 1) > 0 and pre(k, 0) == 0},
                    "untracked_ms_per_step": round(ms_step - fam_total, 3),
                    "window_attn_core_tflops": (pre("attn_fwd_hd32", 0) + pre("attn_bwd_hd32", 0)) /
                                               max(1e-9, (pre("attn_fwd_hd32", 2) + pre("attn_bwd_hd32", 2)) * 1e-3) / 1e12,
                    "step_model_tflops": W["flop"] * clips / (ms_step * 1e-3) / 1e12}
        cpu = eager = None
        max_mem = torch.cuda.max_memory_allocated() / 2 ** 30
        if world == 1 and wl == "c3":
            c2 = window_attention_c2()
            roofline["window_attn_c2_fwd_bwd"] = dict(c2, workload="c2: N=392 window (8,7,7), C=128, 4096 windows, qkv (1.6M x 384 bf16) > L2",
                                                      frac_of_peak={k: round(v["tflops"] / sus, 4) for k, v in c2.items()})
        if world == 1 and wl == "c3" and not args.no_cpu_baseline:
            v, ms, cores = cpu_reference_run(1, 1, 2, regularisers=not args.parity_config)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "2 clips of the same c3 step (fp32 oracle port, fwd+bwd+clip+AdamW), 1 warm-up + 1 timed step"}
        if world == 1 and wl == "c3" and not args.no_eager_baseline:
            del net, opt, devb
            model.cpu()
            from clover_b200 import functional as Fn
            Fn.clear_weight_cache()
            Fn.clear_grad_sinks()
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
            eager = gpu_eager_baseline(dev, not args.parity_config)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(clips, world, not args.parity_config, wl),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "e2e_uint8_staging": {"value": e2e8_value, "unit": UNIT, "h2d_bytes_per_step": h2d8, "d2h_bytes_per_step": 4,
                                  "note": "raw uint8 clips + GPUNormalize folded into the patch gather (utils/module_hooks.py:35-87)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "gpu_eager_baseline": eager, "rank_consistency": consistency,
            "loss": float(loss_host), "max_mem_gb": max_mem,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS),
                    help="BASELINE.json config: c3 pre-train step (the headline, default), c4 retrieval fine-tune, c5 video-QA fine-tune")
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU (default: 64 for c3, 16 for c4 / c5 = the shipped configs)")
    ap.add_argument("--kernels-only", action="store_true", help="run warm-up + timed steps and exit (short command for ncu captures)")
    ap.add_argument("--fp32-allreduce", action="store_true", help="N > 1: all-reduce fp32 gradient buckets instead of bf16")
    ap.add_argument("--bucket-mb", type=int, default=100, help="N > 1: DDP gradient bucket size (developer sweep)")
    ap.add_argument("--static-graph", action="store_true", help="N > 1: DistributedDataParallel(static_graph=True) (developer sweep)")
    ap.add_argument("--no-grad-sinks", action="store_true", help="allocate fresh gradient tensors every step (developer A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager-on-GPU baseline leg (N = 1, c3)")
    ap.add_argument("--parity-config", action="store_true",
                    help="zero dropout / drop-path (the configuration of the parity tests) instead of the shipped training rates")
    ap.add_argument("--trace", default=None, help="developer aid: directory for a torch.profiler kernel timeline of two extra steps")
    ap.add_argument("--lib", default=None, help="developer A/B runs: another build of libclover_b200.so to load")
    args = ap.parse_args()
    if args.lib:
        from clover_b200 import _lib
        _lib.set_library(args.lib)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
