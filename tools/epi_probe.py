"""GEMM epilogue probe: times one launch of the Swin-B hot shapes per epilogue variant (b bias, g GELU + pre-activation copy,
G GELU only, p x GELU'(pre), r residual, 16 / 32 output width).   python tools/epi_probe.py [other libclover_b200.so]"""
import sys, torch
sys.path.insert(0, ".")
from clover_b200 import _lib, ops
if len(sys.argv) > 1:
    _lib.set_library(sys.argv[1])
BF16, F32 = torch.bfloat16, torch.float32
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def run(M,N,K,bt,variant):
    a = torch.randn(M, K, device="cuda").to(BF16)
    b = torch.randn((K, N) if bt else (N, K), device="cuda").to(BF16)
    kw = {}
    od = F32 if "32" in variant else BF16
    out = torch.empty(M, N, dtype=od, device="cuda")
    if "b" in variant: kw["bias"] = torch.randn(N, device="cuda")
    if "g" in variant or "G" in variant: kw["act"] = "gelu"
    if "g" in variant: kw["out_pre"] = torch.empty(M, N, dtype=BF16, device="cuda")
    if "p" in variant: kw["gelu_pre"] = torch.randn(M, N, device="cuda").to(BF16)
    if "r" in variant: kw["residual"] = torch.randn(M, N, device="cuda")
    ms=[]
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, b, out, b_t=bool(bt), **kw); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sorted(ms[1:])[2]
    print(f"M={M} N={N} K={K} bt={bt} {variant:6s} {t:7.3f} ms {2*M*N*K/t/1e9:7.1f} TFLOP/s", flush=True)
for v in ("16", "b16", "bG16", "bg16", "p16", "32", "br32"):
    run(100352, 2048, 512, 0, v)
for box in (1, 0):
    ops.set_tunable("gemm_box", box)
    print(f"# gemm_box = {box}")
    run(100352, 2048, 512, 1, "p16")
    run(1605632, 512, 128, 1, "p16")
    run(401408, 1024, 256, 1, "p16")
for box in (2, 1):
    ops.set_tunable("gemm_box", box)
    print(f"# gemm_box = {box} (2: bias-only / plain bf16 epilogues through TMA boxes too)")
    run(100352, 1536, 512, 0, "b16")
    run(100352, 2048, 512, 0, "b16")
    run(100352, 512, 512, 1, "16")
    run(100352, 512, 2048, 1, "16")
    run(401408, 768, 256, 0, "b16")
    run(401408, 256, 1024, 1, "16")
ops.set_tunable("gemm_box", -1)
run(100352, 1536, 512, 0, "b16")
run(100352, 2048, 512, 1, "p16")
for v in ("16", "bg16", "32"):
    run(1605632, 512, 128, 0, v)
for v in ("16","32","br32"):
    run(100352, 512, 2048, 0, v)
run(8192,8192,8192,0,"16")
