"""Two-GPU NCCL parity of the data-parallel pre-train step: DistributedDataParallel over the clover_b200
recogniser on two ranks against SINGLE-PROCESS runs of the same model on the concatenated batch.

Reference semantics under test (models/utils/gather_loss.py:47-72 + DDP): every rank evaluates the alignment losses on the
all-gathered embeddings (identical values on all ranks, equal to the single-process full-batch values); the backward keeps the
local slice, DDP averages, so the parameter gradient is (1/W) * grad(L_align(full batch)) + (1/W) * sum_r grad(L_mlm(rank r)).

Needs >= 2 CUDA devices (`gpurun --gpus 2`); skipped on one GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
ALIGN = ("nce_loss", "rank_t_tm_loss", "v_nce_loss", "rank_v_vm_loss")
WATCH = ("backbone.patch_embed.proj.weight", "backbone.layers.1.blocks.1.attn.relative_position_bias_table",
         "backbone.layers.0.blocks.0.mlp.fc1.weight", "text_backbone.bert.encoder.layer.1.attention.self.query.weight",
         "multimodal_backbone.fc_in.weight", "ssl_head.img_projector.3.weight", "mlm_head.predictions.decoder.bias",
         "mlm_ssl_T_head.fc2.weight", "backbone.norm.bias")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build():
    from clover_b200 import registry
    from clover_b200.configs import pretrain_cfg
    from clover_b200.synthetic import synth_state_dict
    registry.register_all()
    bert = dict(num_attention_heads=2, intermediate_size=256, max_position_embeddings=64, vocab_size=1000)
    m = registry.build_model(pretrain_cfg(32, (2, 2), (1, 2), 64, 128, 1000, 2, 2, 2, **bert)).cuda()
    m.load_state_dict(synth_state_dict(m.state_dict(), 50), strict=False)
    for n, p in m.named_parameters():
        if ".pooler." in n or ".bert_embedding." in n:
            p.requires_grad_(False)
    return m.train()


def _run(model, batch, rows, which):
    kw = {k: batch[k][rows].cuda() for k in ("token_ids", "segment_ids", "input_mask", "mlm_label", "v_token_mask")}
    losses = model(batch["imgs"][rows].cuda(), batch["label"][rows].cuda(), return_loss=True, **kw)
    sum(losses[k] for k in which).backward()
    return {k: float(v) for k, v in losses.items()}


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from clover_b200.synthetic import make_batch
        per = 2
        batch = make_batch(world * per, frames=4, L=16, seed=51, size=56, vocab=1000)
        model = _build()
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank], broadcast_buffers=False)
        mine = slice(rank * per, (rank + 1) * per)
        losses = _run(ddp, batch, mine, ALIGN + ("mlm_loss",))
        params = dict(model.named_parameters())
        res = {"losses": losses, "grads": {n: params[n].grad.detach().float().cpu().numpy().copy() for n in WATCH}}
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _reference(world, per):
    """Single process, no process group: full-batch alignment losses / gradients and per-rank MLM gradients."""
    from clover_b200.synthetic import make_batch
    batch = make_batch(world * per, frames=4, L=16, seed=51, size=56, vocab=1000)
    model = _build()
    params = dict(model.named_parameters())
    full = _run(model, batch, slice(0, world * per), ALIGN)
    want = {n: params[n].grad.detach().float().cpu().clone() if params[n].grad is not None else torch.zeros_like(params[n]).cpu()
            for n in WATCH}
    mlm = []
    for r in range(world):
        model.zero_grad(set_to_none=True)
        lr_ = _run(model, batch, slice(r * per, (r + 1) * per), ("mlm_loss",))
        mlm.append(lr_["mlm_loss"])
        for n in WATCH:
            if params[n].grad is not None:
                want[n] += params[n].grad.detach().float().cpu()
    return full, mlm, {n: g / world for n, g in want.items()}


@pytest.mark.timeout(600)
def test_ddp_two_ranks_equals_single_process_on_concatenated_batch():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from conftest import record_parity
    world, per = 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full, mlm, want = _reference(world, per)
    rec = {"align": {}, "grad": {}}
    for k in ALIGN:
        a, b = out[0]["losses"][k], out[1]["losses"][k]
        rec["align"][k] = {"rank0": a, "rank1": b, "single_process_full_batch": full[k]}
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (k, a, b)                       # identical on every rank
        assert abs(a - full[k]) <= 1e-4 * max(1.0, abs(full[k])), (k, a, full[k])      # == single process on the concatenated batch
    for r in range(world):                                                             # the per-rank loss stays per-rank
        assert abs(out[r]["losses"]["mlm_loss"] - mlm[r]) <= 1e-4 * max(1.0, abs(mlm[r]))
    for n in WATCH:
        g0, g1 = torch.from_numpy(out[0]["grads"][n]), torch.from_numpy(out[1]["grads"][n])
        assert torch.equal(g0, g1), n                                                  # DDP: same averaged gradient everywhere
        w = want[n]
        relerr = float((g0.double() - w.double()).norm() / w.double().norm().clamp_min(1e-30))
        cosv = float((g0.double().flatten() @ w.double().flatten()) / (g0.double().norm() * w.double().norm()).clamp_min(1e-30))
        rec["grad"][n] = {"rel": relerr, "cos": cosv}
    record_parity("ddp_2rank_nccl_vs_single_process", rec)
    bad = {n: v for n, v in rec["grad"].items() if v["rel"] > 1e-2 or v["cos"] < 0.9999}
    assert not bad, bad
