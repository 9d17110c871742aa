"""Differentiable all-gather of the alignment embeddings (the only model-side collective).

Semantics are the reference's (mmaction/models/utils/gather_loss.py:5-72): forward concatenates every
rank's tensor along dim 0; backward returns ONLY the local slice of the incoming gradient (no
reduce-scatter), so after DDP's gradient averaging the effective parameter gradient is
(1/W) * grad(L_global) -- reproduced on purpose.

B200-native difference: the six embeddings of a pre-train step are stacked and exchanged with ONE
NCCL all-gather (``gather_stacked``) instead of the reference's 8 x (size exchange + padded gather),
with no host synchronisation; the varied-shape variant keeps the size exchange for ragged batches.
"""
import torch
import torch.distributed as dist


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class GatherLoss(torch.autograd.Function):
    """all-gather of equally sized tensors (reference gather_loss.py:5-22)."""

    @staticmethod
    def forward(ctx, tensor, rank, world_size):
        ctx.rank, ctx.batch = rank, tensor.shape[0]
        if world_size == 1:
            return tensor.clone()
        tensor = tensor.contiguous()
        out = torch.empty((world_size * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
        dist.all_gather_into_tensor(out, tensor) if tensor.is_cuda else dist.all_gather(list(out.chunk(world_size)), tensor)
        return out

    @staticmethod
    def backward(ctx, grad):
        return grad[ctx.batch * ctx.rank: ctx.batch * (ctx.rank + 1)], None, None


class VariedShapeGatherLoss(torch.autograd.Function):
    """all-gather of tensors whose dim-0 differs per rank (reference gather_loss.py:24-72)."""

    @staticmethod
    def forward(ctx, q, rank, ws):
        ctx.rank = rank
        if ws == 1:
            ctx.bounds = (0, q.shape[0])
            return q.clone()
        local = torch.tensor([q.shape[0]], device=q.device, dtype=torch.int64)
        sizes = [torch.zeros_like(local) for _ in range(ws)]
        dist.all_gather(sizes, local)
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes)
        if mx != q.shape[0]:
            q = torch.cat([q, q.new_zeros((mx - q.shape[0],) + tuple(q.shape[1:]))])
        bufs = [torch.zeros_like(q) for _ in range(ws)]
        dist.all_gather(bufs, q.contiguous())
        start = sum(sizes[:rank])
        ctx.bounds = (start, start + sizes[rank])
        return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)

    @staticmethod
    def backward(ctx, grad):
        return grad[ctx.bounds[0]: ctx.bounds[1]], None, None


class _StackedGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rank, ws, *tensors):
        ctx.rank, ctx.ws, ctx.n = rank, ws, len(tensors)
        B = tensors[0].shape[0]
        ctx.B = B
        stack = torch.stack([t.float() for t in tensors], 0).contiguous()          # (n, B, D)
        out = torch.empty((ws,) + tuple(stack.shape), dtype=stack.dtype, device=stack.device)
        if stack.is_cuda:
            dist.all_gather_into_tensor(out.view(ws * stack.shape[0], B, -1), stack)
        else:
            dist.all_gather(list(out.unbind(0)), stack)
        # (ws, n, B, D) -> n tensors of (ws*B, D), rank-major like torch.cat(all_gather(...))
        return tuple(out[:, i].reshape(ws * B, -1) for i in range(len(tensors)))

    @staticmethod
    def backward(ctx, *grads):
        r, B = ctx.rank, ctx.B
        return (None, None) + tuple(g[r * B:(r + 1) * B] for g in grads)


def gather_stacked(tensors):
    """One collective for several (B, D) fp32 embeddings with identical shapes on every rank.
    Returns the gathered (W*B, D) tensors; backward keeps the local slice (GatherLoss semantics)."""
    rank, ws = _world()
    if ws == 1:
        return [t.float() for t in tensors]
    return list(_StackedGather.apply(rank, ws, *tensors))


class _StackedVariedGather(torch.autograd.Function):
    """Several (B_r, D) embeddings whose batch B_r may differ per rank (the last, ragged batch of an epoch): ONE size
    exchange (a single host read of W integers) and ONE padded stacked all-gather instead of the reference's size exchange
    + padded gather per tensor (gather_loss.py:47-58, called 4x per loss).  Backward keeps the local slice."""

    @staticmethod
    def forward(ctx, rank, ws, *tensors):
        B = tensors[0].shape[0]
        dev = tensors[0].device
        local = torch.tensor([B], device=dev, dtype=torch.int64)
        sizes = torch.empty(ws, device=dev, dtype=torch.int64)
        if local.is_cuda:
            dist.all_gather_into_tensor(sizes, local)
        else:
            dist.all_gather(list(sizes.split(1)), local)
        sizes = sizes.tolist()
        mx = max(sizes)
        stack = torch.stack([t.float() for t in tensors], 0)                       # (n, B, D)
        if mx != B:
            stack = torch.cat([stack, stack.new_zeros(stack.shape[0], mx - B, stack.shape[2])], 1)
        stack = stack.contiguous()
        out = torch.empty((ws,) + tuple(stack.shape), dtype=stack.dtype, device=dev)
        if stack.is_cuda:
            dist.all_gather_into_tensor(out.view(ws * stack.shape[0], mx, -1), stack)
        else:
            dist.all_gather(list(out.unbind(0)), stack)
        start = sum(sizes[:rank])
        ctx.bounds = (start, start + sizes[rank])
        return tuple(torch.cat([out[r, i, :n] for r, n in enumerate(sizes)], 0) for i in range(len(tensors)))

    @staticmethod
    def backward(ctx, *grads):
        a, b = ctx.bounds
        return (None, None) + tuple(g[a:b] for g in grads)


def gather_stacked_varied(tensors):
    """gather_stacked for batches that may be ragged across ranks (VariedShapeGatherLoss semantics)."""
    rank, ws = _world()
    if ws == 1:
        return [t.float() for t in tensors]
    return list(_StackedVariedGather.apply(rank, ws, *tensors))
