"""WindowAttention3D core microbenchmark (BASELINE.json config c2 and the c3 production shapes).

    python tools/attn_microbench.py [--shapes s1,s3,c2] [--iters 10] [--which fwd,bwd] [--shifted 0|1]

Times the attention core (QK^T + bias + mask + softmax + PV, and its backward incl. the bias-table gradient)
through the C ABI with CUDA events, one launch per iteration on tensors larger than L2, and prints TFLOP/s on the
algorithmic FLOPs of SURVEY.md 8(d) (4*T*N*C forward, 8*T*N*C backward).  Used under ncu for profiles/.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {                      # name: (clips, (D, H, W) tokens, heads)  -- Swin-B, head_dim 32
    "s1": (64, (4, 56, 56), 4),       # c3 stage 1: 4096 windows x 196 tokens, C = 128
    "s2": (64, (4, 28, 28), 8),
    "s3": (64, (4, 14, 14), 16),      # c3 stage 3: 256 windows, C = 512 (18 of the 24 blocks)
    "s4": (64, (4, 7, 7), 32),
    "c2": (64, (8, 56, 56), 4),       # BASELINE c2: window 8x7x7 -> N = 392, stage-1 width
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="s1,s3")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--which", default="fwd,bwd")
    ap.add_argument("--shifted", type=int, default=1)
    ap.add_argument("--dbias", type=int, default=1, help="0: backward without the bias-table gradient (no dS dump)")
    ap.add_argument("--phases", type=int, default=0, help="1: per-phase cycle sums of the forward's softmax warp (w7_fwd_dbg instantiation)")
    ap.add_argument("--clips", type=int, default=0, help="override the clip count of every shape (c3 batches 128 clip-passes)")
    ap.add_argument("--lib", default=None, help="another build of libclover_b200.so (same-box A/B)")
    args = ap.parse_args()
    if args.lib:
        from clover_b200 import _lib
        _lib.set_library(args.lib)
    from clover_b200 import ops, swin, tables
    dev = torch.device("cuda")
    res = []
    for name in args.shapes.split(","):
        clips, dims, heads = SHAPES[name]
        clips = args.clips or clips
        hd = 32
        win, sh = tables.get_window_size(dims, (8, 7, 7), (4, 3, 3) if args.shifted else (0, 0, 0))
        N = win[0] * win[1] * win[2]
        nwin = (dims[0] // win[0]) * (dims[1] // win[1]) * (dims[2] // win[2])
        batch = clips * nwin
        g = torch.Generator(device="cuda").manual_seed(1)
        qkv = (torch.randn(batch * N, 3 * heads * hd, generator=g, device=dev) * 0.7).bfloat16()
        dout = torch.randn(batch * N, heads * hd, generator=g, device=dev).bfloat16()
        table = torch.randn(2535, heads, generator=g, device=dev) * 0.5
        code, off = tables.rel_code(N, (8, 7, 7))
        code = torch.from_numpy(code).to(dev)
        masked = any(s > 0 for s in sh)
        region = torch.from_numpy(tables.region_ids(*dims, win, sh)).to(dev) if masked else None
        spec = swin._w7_spec(dims, win, sh, (8, 7, 7), dev)
        out = torch.empty(batch * N, heads * hd, dtype=torch.bfloat16, device=dev)
        lse = torch.empty(batch, heads, N, dtype=torch.float32, device=dev)
        dqkv = torch.empty_like(qkv)
        dtab = torch.zeros(2535, heads, device=dev)
        kw = dict(bias_table=table, rel_code=code, code_off=off, region=region, w7=spec)
        flops_f = 4.0 * batch * heads * N * N * hd
        row = {"shape": name, "windows": batch, "N": N, "heads": heads, "shifted": bool(masked)}
        ops.attention_fwd(qkv, batch, N, heads, hd, out, lse, **kw)
        for which in args.which.split(","):
            fn = ((lambda: ops.attention_fwd(qkv, batch, N, heads, hd, out, lse, **kw)) if which == "fwd" else
                  (lambda: ops.attention_bwd(qkv, out, dout, lse, batch, N, heads, hd, dqkv, hd ** -0.5, dbias_table=dtab if args.dbias else None, **kw)))
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
            fl = flops_f if which == "fwd" else 2 * flops_f
            row[which + "_ms"] = round(ms, 4)
            row[which + "_tflops"] = round(fl / ms / 1e9, 1)
        if args.phases:
            dbg = torch.zeros(296 * 8, dtype=torch.int64, device=dev)
            ops.set_tunable("w7_fwd_dbg", dbg.data_ptr())
            try:
                ops.attention_fwd(qkv, batch, N, heads, hd, out, lse, **kw)
                torch.cuda.synchronize()
            finally:
                ops.set_tunable("w7_fwd_dbg", -1)
            d = dbg.view(296, 8).cpu().double()
            d = d[d[:, 5] > 0]
            per_tile = (d[:, :5].sum(0) / d[:, 5].sum()).tolist()
            row["fwd_phase_cycles_per_tile"] = dict(zip(["wait_s", "pass1", "pass2", "wait_o", "epilogue"], [round(x) for x in per_tile]))
            row["fwd_tiles_per_cta"] = round(float(d[:, 5].mean()), 1)
        res.append(row)
        print(json.dumps(row), flush=True)
    return res


if __name__ == "__main__":
    main()
