"""world_size-2 gloo tests (CPU) of the only model-side collective: the differentiable all-gather of the
alignment embeddings.  Semantics under test are the reference's (models/utils/gather_loss.py:5-72):
forward = concatenation over ranks, backward = LOCAL slice only, so DDP-averaged parameter gradients equal
(1/W) * grad of the global loss."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clover_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clover_b200.gather import GatherLoss, VariedShapeGatherLoss, gather_stacked, gather_stacked_varied
        res = {}
        g = torch.Generator().manual_seed(100 + rank)
        # fixed-size gather
        x = torch.randn(3, 8, generator=g, requires_grad=True)
        y = GatherLoss.apply(x, rank, world)
        w = torch.arange(y.numel(), dtype=torch.float32).view_as(y)
        (y * w).sum().backward()
        res["fixed_out"], res["fixed_grad"], res["fixed_w"] = y.detach(), x.grad.clone(), w
        # ragged gather (rank r holds 2 + r rows)
        xr = torch.randn(2 + rank, 4, generator=g, requires_grad=True)
        yr = VariedShapeGatherLoss.apply(xr, rank, world)
        wr = torch.arange(yr.numel(), dtype=torch.float32).view_as(yr) * 0.5
        (yr * wr).sum().backward()
        res["ragged_out"], res["ragged_grad"], res["ragged_in"], res["ragged_w"] = yr.detach(), xr.grad.clone(), xr.detach(), wr
        # ragged STACKED gather (one size exchange + one collective for several tensors)
        xs = [torch.randn(2 + rank, 4, generator=g, requires_grad=True) for _ in range(3)]
        ys = gather_stacked_varied(xs)
        ws_ = [torch.arange(y_.numel(), dtype=torch.float32).view_as(y_) * (i + 1) for i, y_ in enumerate(ys)]
        sum((y_ * w_).sum() for y_, w_ in zip(ys, ws_)).backward()
        res["sv_out"], res["sv_in"] = [y_.detach() for y_ in ys], [x_.detach() for x_ in xs]
        res["sv_grad"], res["sv_w"] = [x_.grad.clone() for x_ in xs], ws_
        # stacked gather of several embeddings + the loss evaluated on the global batch
        embs = [torch.randn(4, 16, generator=g, requires_grad=True) for _ in range(4)]
        gathered = gather_stacked(embs)
        loss = O.exclusive_nce_ranking(*gathered, t=0.05, margin=5.0)
        (loss["nce_loss"] + loss["rank_t_tm_loss"]).backward()
        res["embs"] = [e.detach() for e in embs]
        res["emb_grads"] = [e.grad.clone() for e in embs]
        res["loss"] = float(loss["nce_loss"] + loss["rank_t_tm_loss"])
        res["x"] = x.detach()
        # logged scalars: one packed all-reduce, every entry averaged over the ranks (recognizers/base.py:254-288); the
        # returned loss tensor stays local (DDP averages its gradient)
        from clover_b200.recognizers import BaseRecognizer
        lt = torch.tensor(float(rank + 1), requires_grad=True)
        total, log_vars = BaseRecognizer._parse_losses({"a_loss": lt * 2.0, "acc": torch.tensor(10.0 * rank), "b_loss": [lt, lt]})
        res["parse_total"], res["parse_log"] = float(total.detach()), dict(log_vars)
        # evaluation: ragged forward_test outputs gathered as tensors (no pickling)
        from clover_b200.evaluation import gather_embeddings
        ev_v, ev_t = torch.randn(3 + rank, 6, generator=g), torch.randn(3 + rank, 5, generator=g)
        gv, gt_ = gather_embeddings(ev_v, ev_t)
        res["ev_in"], res["ev_out"] = [ev_v, ev_t], [gv, gt_]
        # tensors travel through the queue as shared-memory handles that die with this process: send them by value
        res = {k: ([t.numpy().copy() for t in v] if isinstance(v, list) else (v.numpy().copy() if torch.is_tensor(v) else v))
               for k, v in res.items()}
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gather_semantics_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=150) for _ in range(world))
    out = {r: {k: ([torch.from_numpy(a) for a in v] if isinstance(v, list) else (torch.from_numpy(v) if hasattr(v, "dtype") else v))
               for k, v in res.items()} for r, res in out.items()}
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    r0, r1 = out[0], out[1]
    # forward: rank-major concatenation, identical on both ranks
    full = torch.cat([r0["x"], r1["x"]])
    assert torch.equal(r0["fixed_out"], full) and torch.equal(r1["fixed_out"], full)
    # backward: only the local slice of the upstream gradient
    assert torch.equal(r0["fixed_grad"], r0["fixed_w"][:3]) and torch.equal(r1["fixed_grad"], r1["fixed_w"][3:])
    ragged = torch.cat([r0["ragged_in"], r1["ragged_in"]])
    assert torch.equal(r0["ragged_out"], ragged) and torch.equal(r1["ragged_out"], ragged)
    assert torch.equal(r0["ragged_grad"], r0["ragged_w"][:2]) and torch.equal(r1["ragged_grad"], r1["ragged_w"][2:])
    for i in range(3):
        cat = torch.cat([r0["sv_in"][i], r1["sv_in"][i]])
        assert torch.equal(r0["sv_out"][i], cat) and torch.equal(r1["sv_out"][i], cat)
        assert torch.equal(r0["sv_grad"][i], r0["sv_w"][i][:2]) and torch.equal(r1["sv_grad"][i], r1["sv_w"][i][2:])
    # global loss identical on every rank; local grads == the matching rows of the single-process full-batch gradient
    assert abs(r0["loss"] - r1["loss"]) < 1e-6
    glob = [torch.cat([r0["embs"][i], r1["embs"][i]]).requires_grad_(True) for i in range(4)]
    ref = O.exclusive_nce_ranking(*glob, t=0.05, margin=5.0)
    (ref["nce_loss"] + ref["rank_t_tm_loss"]).backward()
    assert abs(float(ref["nce_loss"] + ref["rank_t_tm_loss"]) - r0["loss"]) < 1e-5
    for i in range(4):
        assert torch.allclose(r0["emb_grads"][i], glob[i].grad[:4], atol=1e-6)
        assert torch.allclose(r1["emb_grads"][i], glob[i].grad[4:], atol=1e-6)
    for i in range(2):
        cat = torch.cat([r0["ev_in"][i], r1["ev_in"][i]])
        assert torch.equal(r0["ev_out"][i], cat) and torch.equal(r1["ev_out"][i], cat)
    # _parse_losses: local total = sum of the '*loss*' entries; logged values are rank means
    assert r0["parse_total"] == 4.0 and r1["parse_total"] == 8.0
    for r in (r0, r1):
        assert r["parse_log"] == {"a_loss": 3.0, "acc": 5.0, "b_loss": 3.0, "loss": 6.0}
