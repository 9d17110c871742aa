"""LayerNorm kernel microbenchmark at the Swin-B c3 shapes (128 clip-passes): achieved algorithmic GB/s of clv_lnr_fwd /
clv_lnr_bwd in the block configurations (norm1 through the window row map, norm2 with the residual gradient).
    python tools/ln_microbench.py            # prints one JSON line per case"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from clover_b200 import ops, tables
    dev = torch.device("cuda", 0)
    peak = 6547.5
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    clips = 128
    for C, dims in ((128, (4, 56, 56)), (256, (4, 28, 28)), (512, (4, 14, 14)), (1024, (4, 7, 7))):
        T = clips * dims[0] * dims[1] * dims[2]
        win, sh = tables.get_window_size(dims, (8, 7, 7), (4, 3, 3))
        rmap = torch.from_numpy(tables.window_row_map(*dims, win, sh)).to(dev)
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(T, C, device=dev, generator=g)
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        y16 = torch.empty(T, C, dtype=torch.bfloat16, device=dev)
        mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
        dy16 = torch.randn(T, C, device=dev, generator=g).bfloat16()
        dres = torch.randn(T, C, device=dev, generator=g)
        dx = torch.empty(T, C, device=dev)
        dx16 = torch.empty(T, C, dtype=torch.bfloat16, device=dev)
        small = torch.zeros(3 * C, device=dev)
        cases = {
            "fwd_norm2": (lambda: ops.lnr_fwd(x, gamma, beta, 1e-5, y16, mean=mean, rstd=rstd), 6),
            "fwd_norm1_mapped": (lambda: ops.lnr_fwd(x, gamma, beta, 1e-5, y16, mean=mean, rstd=rstd, row_map=rmap, y_mapped=True), 6),
            "bwd_norm2": (lambda: ops.lnr_bwd(x, gamma, beta, 1e-5, mean, rstd, dy16, dx=dx, dres=dres, dx_bf16=dx16, row_map=rmap,
                                              dx_bf16_mapped=True, dgamma=small[:C], dbeta=small[C:2 * C], dxsum=small[2 * C:]), 16),
            "bwd_norm1": (lambda: ops.lnr_bwd(x, gamma, beta, 1e-5, mean, rstd, dy16, dx=dx, dres=dx, dx_bf16=dx16, row_map=rmap,
                                              dy_mapped=True, dgamma=small[:C], dbeta=small[C:2 * C], dxsum=small[2 * C:]), 16),
        }
        for name, (fn, bpe) in cases.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            gbs = T * C * bpe / ms / 1e6
            print(json.dumps({"case": name, "rows": T, "C": C, "ms": round(ms, 4), "algorithmic_GBps": round(gbs, 1),
                              "frac_of_measured_hbm": round(gbs / peak, 3), "bytes_per_element": bpe}), flush=True)
    # PatchMerging LN (2x2 gather + LN(4C)): fp32 tokens in (4 B), bf16 out (2 B) forward; x + dy in, fp32 dx out backward (10 B)
    for C, dims in ((128, (4, 56, 56)), (256, (4, 28, 28)), (512, (4, 14, 14))):
        D, H, W = dims
        T = clips * D * H * W
        rows = T // 4
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(T, C, device=dev, generator=g)
        gamma, beta = torch.ones(4 * C, device=dev), torch.zeros(4 * C, device=dev)
        y16 = torch.empty(rows, 4 * C, dtype=torch.bfloat16, device=dev)
        mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
        dy16 = torch.randn(rows, 4 * C, device=dev, generator=g).bfloat16()
        dx = torch.empty(T, C, device=dev)
        small = torch.zeros(8 * C, device=dev)
        cases = {
            "fwd_merge": (lambda: ops.layernorm_fwd(x, gamma, beta, 1e-5, y16, mean=mean, rstd=rstd, merge=(clips, D, H, W, C)), 6),
            "bwd_merge": (lambda: ops.layernorm_bwd(x, gamma, beta, 1e-5, mean, rstd, dy16, rows=rows, dx=dx, dgamma=small[:4 * C],
                                                     dbeta=small[4 * C:], merge=(clips, D, H, W, C)), 10),
        }
        for name, (fn, bpe) in cases.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            gbs = T * C * bpe / ms / 1e6
            print(json.dumps({"case": name, "rows": rows, "C": 4 * C, "ms": round(ms, 4), "algorithmic_GBps": round(gbs, 1),
                              "frac_of_measured_hbm": round(gbs / peak, 3), "bytes_per_element": bpe}), flush=True)


if __name__ == "__main__":
    main()
