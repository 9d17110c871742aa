"""Micro-benchmark of the tcgen05 GEMM on the shapes that dominate the Clover pre-train step (c3, 64 clips/GPU).
   python tools/gemm_probe.py [reps]      -> TFLOP/s and GB/s per shape (CUDA events, L2 flushed between reps)"""
import sys

import torch

sys.path.insert(0, ".")
from clover_b200 import ops  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
# (M, N, K, a_t, b_t, epilogue)  epilogue: b=bias g=gelu(+pre) r=residual fp32 p=gelu' multiply, o16/o32 output, ks = split-K
SHAPES = [
    # c3 with clean + masked passes batched (128 clip-passes per launch)
    (100352, 2048, 512, 0, 0, "bg", "o16", 1), (100352, 2048, 512, 0, 1, "p", "o16", 1), (100352, 512, 2048, 0, 0, "br", "o32", 1),
    (512, 2048, 100352, 1, 1, "", "o32", 4),
    # tile-shape sweep for the weight gradients (tag: bn256): k_splits >= 10 makes the dispatcher pick 128x256 tiles
    (512, 2048, 100352, 1, 1, "", "o32", 10), (2048, 512, 100352, 1, 1, "", "o32", 10), (2048, 512, 100352, 1, 1, "", "o32", 4),
    (1536, 512, 100352, 1, 1, "", "o32", 6), (1536, 512, 100352, 1, 1, "", "o32", 13), (512, 512, 100352, 1, 1, "", "o32", 18),
    (512, 2048, 100352, 1, 1, "", "o32", 5), (512, 2048, 100352, 1, 1, "", "o32", 9), (2048, 512, 100352, 1, 1, "", "o32", 5),
    (2048, 512, 100352, 1, 1, "", "o32", 9), (1536, 512, 100352, 1, 1, "", "o32", 4), (1536, 512, 100352, 1, 1, "", "o32", 8),
    (50176, 2048, 512, 0, 0, "bg", "o16", 1), (50176, 2048, 512, 0, 1, "p", "o16", 1), (50176, 1536, 512, 0, 0, "b", "o16", 1),
    (50176, 512, 2048, 0, 0, "br", "o32", 1), (50176, 512, 2048, 0, 1, "", "o16", 1), (50176, 512, 512, 0, 0, "br", "o32", 1),
    (802816, 512, 128, 0, 0, "bg", "o16", 1), (802816, 512, 128, 0, 1, "p", "o16", 1), (802816, 384, 128, 0, 0, "b", "o16", 1),
    (802816, 128, 512, 0, 0, "br", "o32", 1), (200704, 1024, 256, 0, 0, "bg", "o16", 1),
    (512, 2048, 50176, 1, 1, "", "o32", 4), (2048, 512, 50176, 1, 1, "", "o32", 4), (512, 512, 50176, 1, 1, "", "o32", 18),
    (128, 512, 802816, 1, 1, "", "o32", 74), (8192, 8192, 8192, 0, 0, "", "o16", 1),
    # split-K sweep for the stage-3 weight gradients (tag: wsweep)
    (512, 2048, 50176, 1, 1, "", "o32", 9), (512, 2048, 50176, 1, 1, "", "o32", 14), (512, 2048, 50176, 1, 1, "", "o32", 18),
    (2048, 512, 50176, 1, 1, "", "o32", 9), (2048, 512, 50176, 1, 1, "", "o32", 18),
    (1536, 512, 50176, 1, 1, "", "o32", 6), (1536, 512, 50176, 1, 1, "", "o32", 12), (1536, 512, 50176, 1, 1, "", "o32", 24),
    (512, 512, 50176, 1, 1, "", "o32", 9), (512, 512, 50176, 1, 1, "", "o32", 37),
    # BERT-base text encoder weight gradients at the c3 batch (4096 caption tokens; tag: K=4096) and the fusion encoder's (K=29184)
    (3072, 768, 4096, 1, 1, "", "o32", 1), (3072, 768, 4096, 1, 1, "", "o32", 2), (3072, 768, 4096, 1, 1, "", "o32", 4),
    (768, 3072, 4096, 1, 1, "", "o32", 1), (768, 3072, 4096, 1, 1, "", "o32", 2), (768, 3072, 4096, 1, 1, "", "o32", 4),
    (2304, 768, 4096, 1, 1, "", "o32", 1), (2304, 768, 4096, 1, 1, "", "o32", 2), (2304, 768, 4096, 1, 1, "", "o32", 4),
    (768, 768, 4096, 1, 1, "", "o32", 2), (768, 768, 4096, 1, 1, "", "o32", 4), (768, 768, 4096, 1, 1, "", "o32", 8),
    (768, 768, 4096, 1, 1, "", "o32", 16),
    # stage-1 weight gradients (K = 1.6 M token rows, tiny outputs; tag: K=1605632)
    (512, 128, 1605632, 1, 1, "", "o32", 74), (512, 128, 1605632, 1, 1, "", "o32", 37), (512, 128, 1605632, 1, 1, "", "o32", 148),
    (128, 512, 1605632, 1, 1, "", "o32", 74), (128, 512, 1605632, 1, 1, "", "o32", 148), (384, 128, 1605632, 1, 1, "", "o32", 98),
    (128, 128, 1605632, 1, 1, "", "o32", 296), (128, 128, 1605632, 1, 1, "", "o32", 148),
    (4096, 3072, 768, 0, 0, "bg", "o16", 1), (4096, 768, 3072, 0, 0, "b", "o16", 1), (4096, 2304, 768, 0, 0, "b", "o16", 1),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    only = sys.argv[2] if len(sys.argv) > 2 else None
    rowsum = len(sys.argv) > 3 and sys.argv[3] == "rowsum"      # weight-gradient shapes also produce the bias gradient
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for (M, N, K, at, bt, ep, od, ks) in SHAPES:
        tag = f"M={M} N={N} K={K} at={at} bt={bt} ep={ep}{od} ks={ks}"
        if only and only not in tag:
            continue
        a = torch.randn((K, M) if at else (M, K), device="cuda").to(BF16)
        b = torch.randn((K, N) if bt else (N, K), device="cuda").to(BF16)
        out = torch.empty(M, N, dtype=BF16 if od == "o16" else F32, device="cuda")
        kw = {}
        if "b" in ep:
            kw["bias"] = torch.randn(N, device="cuda")
        if "g" in ep:
            kw["act"] = "gelu"
            kw["out_pre"] = torch.empty(M, N, dtype=BF16, device="cuda")
        if "p" in ep:
            kw["gelu_pre"] = torch.randn(M, N, device="cuda").to(BF16)
        if "r" in ep:
            kw["residual"] = torch.randn(M, N, device="cuda")
        if rowsum and at and od == "o32":
            kw["rowsum"] = torch.zeros(M, device="cuda")
            tag += " +rowsum"
        ms = []
        for _ in range(reps + 1):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(a, b, out, a_t=bool(at), b_t=bool(bt), k_splits=ks, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = sorted(ms[1:])[len(ms[1:]) // 2]
        nbytes = 2 * (M * K + N * K) + out.element_size() * M * N * (2 if "g" in ep else 1) + (2 * M * N if "p" in ep else 0) + (4 * M * N if "r" in ep else 0)
        print(f"{tag:70s} {t:8.3f} ms  {2 * M * N * K / t / 1e9:8.1f} TFLOP/s  {nbytes / t / 1e6:8.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
