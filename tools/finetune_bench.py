"""Throughput of the fine-tune recognisers at the BASELINE c4 / c5 shapes: Video Swin-B + BERT-base +
3-layer fusion, 16x224x224 clips (T = 8 -> the full (8,7,7) window, N = 392), 40-token queries, bf16.

    python tools/finetune_bench.py [--task retrieval|video_qa] [--clips 16] [--steps 5]

Prints one JSON line per task: clips/s of forward + backward + FusedAdamW on one B200 (CUDA events, inputs resident)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="retrieval,video_qa")
    ap.add_argument("--clips", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    from clover_b200 import registry
    from clover_b200.configs import finetune_cfg
    from clover_b200.optim import FusedAdamW, param_groups_from_cfg
    from clover_b200.synthetic import make_finetune_batch
    registry.register_all()
    dev = torch.device("cuda", 0)
    for task in args.task.split(","):
        torch.manual_seed(0)
        cfg = finetune_cfg(task, frames_half=8, bert_dropout=0.1, qa_dropout=0.1)
        cfg["backbone"]["drop_path_rate"] = 0.3
        model = registry.build_model(cfg).to(dev).train()
        for n, p in model.named_parameters():
            if ".pooler." in n or ".bert_embedding." in n or (task == "retrieval" and n.startswith("multimodal_backbone.")):
                p.requires_grad_(False)
        opt = FusedAdamW(param_groups_from_cfg(model, 1e-7, 0.001, dict(norm_decay_mult=0.0, bias_decay_mult=0.0,
                                                                        custom_keys={"qa_head": dict(lr_mult=10)})),
                         betas=(0.9, 0.98), max_grad_norm=50.0)
        b = {k: v.to(dev) for k, v in make_finetune_batch(task, args.clips, frames=16, size=224, L=40, seed=3).items()}
        kw = {k: b[k] for k in ("token_ids", "segment_ids", "input_mask")}

        def step():
            losses = model(b["imgs"], b["label"], return_loss=True, **kw)
            loss = sum(v for k, v in losses.items() if "loss" in k)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"task": task, "clips_per_gpu": args.clips, "frames": 16, "query_tokens": 40, "ms_per_step": round(ms, 2),
                          "clips_per_sec": round(args.clips / ms * 1e3, 1), "loss": float(loss.detach()),
                          "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                          "regularisers": "shipped (drop_path 0.3, BERT 0.1, QA head 0.1)"}), flush=True)
        del model, opt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
