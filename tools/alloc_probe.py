import sys, os, json, time
sys.path.insert(0, ".")
sys.argv = ["bench.py", "--steps", "1", "--warmup", "0"]
import torch
import bench
from clover_b200 import registry
from clover_b200.synthetic import make_batch
from clover_b200.optim import FusedAdamW, param_groups_from_cfg
registry.register_all()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = registry.build_model(bench.model_cfg(True, "c3")).to(dev); model.train()
for n, p in model.named_parameters():
    if ".pooler." in n or ".bert_embedding." in n: p.requires_grad_(False)
opt = FusedAdamW(param_groups_from_cfg(model, 1e-7, 0.02, dict(norm_decay_mult=0.0, bias_decay_mult=0.0)), betas=(0.9, 0.98), eps=1e-8, max_grad_norm=15, reuse_grad_buffers=True)
raw = make_batch(64, frames=8, L=32, seed=1000)
b = {k: v.to(dev) for k, v in raw.items()}
keys = ("token_ids", "segment_ids", "input_mask", "mlm_label", "v_token_mask")
def step():
    losses = model(b["imgs"], b["label"], return_loss=True, **{k: b[k] for k in keys})
    loss = sum(v for k, v in losses.items() if "loss" in k)
    loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
import gc
for i in range(12):
    s0 = torch.cuda.memory_stats()
    g0 = gc.get_count()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    s1 = torch.cuda.memory_stats()
    print(i, "cpu enqueue ms", round((t1 - t0) * 1e3, 1), "total ms", round((t2 - t0) * 1e3, 1), "cudaMalloc segments +", s1["num_device_alloc"] - s0["num_device_alloc"], "free +", s1["num_device_free"] - s0["num_device_free"], "retries", s1["num_alloc_retries"], "reserved GB", round(s1["reserved_bytes.all.current"] / 2**30, 1), "gc", g0, gc.get_count(), flush=True)
