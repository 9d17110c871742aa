"""One GEMM shape / epilogue variant, three launches (for ncu captures):  python tools/gemm_one.py M N K bt variant
variant letters: b bias, g GELU + pre-activation copy, G GELU only, p x GELU'(pre), r residual, 16 / 32 output width."""
import sys

import torch

sys.path.insert(0, ".")
from clover_b200 import ops

BF16, F32 = torch.bfloat16, torch.float32
M, N, K, bt = (int(v) for v in sys.argv[1:5])
variant = sys.argv[5]
a = torch.randn(M, K, device="cuda").to(BF16)
b = torch.randn((K, N) if bt else (N, K), device="cuda").to(BF16)
kw = {}
out = torch.empty(M, N, dtype=F32 if "32" in variant else BF16, device="cuda")
if "b" in variant: kw["bias"] = torch.randn(N, device="cuda")
if "g" in variant or "G" in variant: kw["act"] = "gelu"
if "g" in variant: kw["out_pre"] = torch.empty(M, N, dtype=BF16, device="cuda")
if "p" in variant: kw["gelu_pre"] = torch.randn(M, N, device="cuda").to(BF16)
if "r" in variant: kw["residual"] = torch.randn(M, N, device="cuda")
for _ in range(3):
    ops.gemm(a, b, out, b_t=bool(bt), **kw)
torch.cuda.synchronize()
