"""BERT / fusion attention core microbenchmark (head_dim 64, 12 heads, key mask, dropout 0.1 as in training) at the c3 shapes
(128 sample-passes): text S = 32, fusion S = 228; c5 fusion S = 432.  Compares the tcgen05 kernels (attention_h64.cu) with the
mma.sync kernels they replace.      python tools/attn64_microbench.py [--drop 0.1]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--drop", type=float, default=0.1)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    from clover_b200 import ops
    dev = torch.device("cuda", 0)
    heads, hd = 12, 64
    for name, batch, seq in (("text_c3", 128, 32), ("fusion_c3", 128, 228), ("fusion_c5", 32, 432)):
        g = torch.Generator(device=dev).manual_seed(3)
        qkv = (torch.randn(batch * seq, 3 * heads * hd, generator=g, device=dev) * 0.5).bfloat16()
        dout = torch.randn(batch * seq, heads * hd, generator=g, device=dev).bfloat16()
        km = torch.zeros(batch, seq, device=dev)
        km[:, seq - 5:] = -10000.0
        out = torch.empty(batch * seq, heads * hd, dtype=torch.bfloat16, device=dev)
        lse = torch.empty(batch, heads, seq, device=dev)
        dqkv = torch.empty_like(qkv)
        drop = (args.drop, 99, 0) if args.drop > 0 else None
        row = {"shape": name, "batch": batch, "seq": seq, "heads": heads, "drop_p": args.drop}
        for path in ("tcgen05", "mma_sync"):
            ops.USE_TC64_ATTENTION = path == "tcgen05"
            for which in ("fwd", "bwd"):
                fn = (lambda: ops.attention_fwd(qkv, batch, seq, heads, hd, out, lse, key_mask=km, drop=drop)) if which == "fwd" else \
                     (lambda: ops.attention_bwd(qkv, out, dout, lse, batch, seq, heads, hd, dqkv, 0.125, key_mask=km, drop=drop))
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                fl = 4.0 * batch * heads * seq * seq * hd * (1 if which == "fwd" else 2)
                row[f"{path}_{which}_ms"] = round(ms, 4)
                row[f"{path}_{which}_tflops"] = round(fl / ms / 1e9, 1)
        ops.USE_TC64_ATTENTION = True
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
