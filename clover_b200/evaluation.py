"""Retrieval / multiple-choice evaluation on the device.

Same function names, arguments and result dicts as the reference's numpy implementations
(mmaction/core/evaluation/accuracy.py:398-456), which the evaluation hook calls on rank 0 after gathering the
``forward_test`` embeddings (core/hooks/my_eval_hook.py:317-401).  Embeddings may be numpy arrays (the reference's
contract) or CUDA tensors (no host hop); the N x N score matrix, the ground-truth ranks and the R@k / MedR reductions
all stay on the GPU, one scalar read per metric at the end.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _dev(x):
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("clover_b200.evaluation needs a CUDA device (no CPU fallback exists)")
        t = t.cuda()
    return t.float().contiguous()


def gather_embeddings(*tensors, group=None):
    """Cross-rank gather of per-rank `forward_test` outputs (rows may differ per rank: the last batches of a sharded test
    set) as TENSORS: one size exchange + one padded all-gather per tensor, no pickling and no host hop (the reference
    pickles every result list through a byte tensor, core/hooks/my_eval_hook.py:317-401 -> mmcv collect_results_gpu).
    Returns the row-wise concatenation over ranks, in rank order, on every rank (a tuple when several tensors are given)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors[0] if len(tensors) == 1 else tuple(tensors)
    world = dist.get_world_size(group)
    sizes = torch.tensor([t.shape[0] for t in tensors], dtype=torch.int64, device=tensors[0].device)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu()                      # [world, len(tensors)]: the one host read
    out = []
    for k, t in enumerate(tensors):
        n_max = int(all_sizes[:, k].max())
        pad = t.new_zeros((n_max,) + tuple(t.shape[1:]))
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad.contiguous(), group=group)
        out.append(torch.cat([p[:int(all_sizes[r, k])] for r, p in enumerate(parts)], 0))
    return out[0] if len(out) == 1 else tuple(out)


def cosine_scores(a, b):
    """scores[i, j] = <a_i / |a_i|, b_j / |b_j|> fp32 (rows of zero norm are left unscaled, numpy_norm.py:5-8)."""
    a, b = _dev(a), _dev(b)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError(f"cosine_scores: expected (N, D) and (M, D), got {tuple(a.shape)} and {tuple(b.shape)}")
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    ws = torch.empty((a.shape[0] + b.shape[0]) * a.shape[1], dtype=torch.float32, device=a.device)
    rc = _lib.load().clv_cosine_scores(C.c_void_p(a.data_ptr()), a.shape[0], C.c_void_p(b.data_ptr()), b.shape[0], a.shape[1],
                                       C.c_void_p(out.data_ptr()), out.stride(0), C.c_void_p(ws.data_ptr()),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "clv_cosine_scores")
    return out


def retrieval_ranks(scores, gt_col=None):
    """int32 [rows]: position of the ground-truth column (default: column == row) in the descending order of each row."""
    s = _dev(scores)
    gt = None
    if gt_col is not None:
        gt = torch.as_tensor(gt_col).to(device=s.device, dtype=torch.int32).contiguous()
        if gt.numel() != s.shape[0] or int(gt.min()) < 0 or int(gt.max()) >= s.shape[1]:
            raise ValueError("retrieval_ranks: gt_col must hold one valid column per row")
    out = torch.empty(s.shape[0], dtype=torch.int32, device=s.device)
    rc = _lib.load().clv_retrieval_ranks(C.c_void_p(s.data_ptr()), s.stride(0), s.shape[0], s.shape[1],
                                         C.c_void_p(gt.data_ptr()) if gt is not None else None, C.c_void_p(out.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "clv_retrieval_ranks")
    return out


def recall_for_video_text_retrieval(video_embd=None, text_embd=None, input_scores=None, use_sim=False, texts=None):
    """reference accuracy.py:427-456: text -> video retrieval R@1 / R@5 / R@10 (percent), median rank, 'Recall@all'."""
    scores = _dev(input_scores) if input_scores is not None else cosine_scores(text_embd, video_embd)
    ind = retrieval_ranks(scores)
    n = ind.numel()
    metrics = {"Recall@1": float((ind == 0).sum()) / n * 100, "Recall@5": float((ind < 5).sum()) / n * 100,
               "Recall@10": float((ind < 10).sum()) / n * 100}
    srt = torch.sort(ind).values.double()                         # np.median: mean of the two middle values for even n
    metrics["MR"] = float((srt[(n - 1) // 2] + srt[n // 2]) / 2) + 1
    metrics["Recall@all"] = metrics["Recall@1"] + metrics["Recall@5"] + metrics["Recall@10"] - metrics["MR"]
    return metrics


def acc_for_msrvtt_mc(video_embd=None, text_embd=None, label=None, use_sim=False, texts=None):
    """reference accuracy.py:398-424 (MSRVTT / LSMDC multiple choice): video i is scored against its own ans_num candidate
    sentences text[i * ans_num : (i + 1) * ans_num]; accuracy of the arg-max.  use_sim selects cosine scores, else raw dots."""
    v, t = _dev(video_embd), _dev(text_embd)
    ans = t.shape[0] // v.shape[0]
    if use_sim:
        scores = cosine_scores(v, t)
    else:
        scores = torch.empty(v.shape[0], t.shape[0], dtype=torch.float32, device=v.device)
        torch.mm(v, t.t(), out=scores)                           # library GEMM on un-normalised embeddings (not a hot path)
    own = scores.view(v.shape[0], v.shape[0], ans)[torch.arange(v.shape[0]), torch.arange(v.shape[0])]
    pred = own.argmax(dim=-1)
    lab = torch.as_tensor(np.asarray(label)).to(pred.device).view(-1)
    return {"acc": float((pred == lab).double().mean())}
