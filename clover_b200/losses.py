"""Losses on the clover_b200 kernels, same names / constructors / return values as
mmaction/models/losses/{contrastive_loss,focal_loss,cross_entropy_loss}.py.  All run in fp32
(the reference decorates them with @force_fp32)."""
import torch
import torch.nn as nn

from . import functional as Fn
from .gather import GatherLoss, VariedShapeGatherLoss, _world, gather_stacked, gather_stacked_varied


class NormSoftmaxLoss(nn.Module):
    """reference contrastive_loss.py:28-68."""

    def __init__(self, temperature=0.07, cos_sim=False):
        super().__init__()
        self.t, self.use_cos_similarity = temperature, cos_sim
        self.allgather = GatherLoss.apply
        self.rank, self.world_size = _world()
        self.fp16_enabled = False

    def forward(self, video_embd=None, text_embd=None, sim_mat=None):
        if sim_mat is not None:
            raise NotImplementedError("clover_b200: NormSoftmaxLoss(sim_mat=...) is not supported")
        self.rank, self.world_size = _world()
        v, t = gather_stacked([video_embd, text_embd])          # one collective (GatherLoss semantics: equal batch per rank)
        return self.forward_gathered(v, t)

    def forward_gathered(self, v, t):
        # cos_sim=True: x / max(||x||, 1e-8) (sim_matrix :10-18); else F.normalize (eps 1e-12)
        eps = 1e-8 if self.use_cos_similarity else 1e-12
        nce, _ = Fn.NceRankFn.apply(self.t, 0.0, False, eps, v, t)
        return nce


class ExclusiveNCEwithRankingLoss(nn.Module):
    """reference contrastive_loss.py:71-161."""

    def __init__(self, temperature=0.05, use_rank=False, use_rank_ttm=True, use_rank_trtm=True, margin_ttm=5.0,
                 margin_trtm=10.0):
        super().__init__()
        self.t = temperature
        self.allgather = VariedShapeGatherLoss.apply
        self.rank, self.world_size = _world()
        self.margin_ttm, self.margin_trtm = margin_ttm, margin_trtm
        self.use_rank, self.use_rank_ttm, self.use_rank_trtm = use_rank, use_rank_ttm, use_rank_trtm
        self.fp16_enabled = False

    def forward(self, video_embd=None, text_embd=None, text_mask_embd=None, text_recon_embd=None, **kwargs):
        if any(e is None for e in (video_embd, text_embd, text_mask_embd, text_recon_embd)):
            raise NotImplementedError("clover_b200: ExclusiveNCEwithRankingLoss needs all four embeddings "
                                      "(use_Cmask=True, as in the shipped pre-train config)")
        self.rank, self.world_size = _world()
        # one size exchange + one stacked collective for the four embeddings (ragged batches allowed, :111-114)
        g = gather_stacked_varied([video_embd, text_embd, text_mask_embd, text_recon_embd])
        return self.forward_gathered(*g)

    def forward_gathered(self, v, t, tm, tr):
        rank = self.use_rank and self.use_rank_ttm
        nce, rk = Fn.NceRankFn.apply(self.t, self.margin_ttm, rank, 1e-8, v, t, tm, tr)
        losses = {"nce_loss": nce}
        if rank:
            losses["rank_t_tm_loss"] = rk
        return losses


class SoftmaxFocalLossMultiClass(nn.Module):
    """reference focal_loss.py:49-72."""

    def __init__(self, gamma=2.0, reduction="mean"):
        super().__init__()
        if reduction != "mean":
            raise NotImplementedError("clover_b200: only reduction='mean' is supported")
        self.gamma, self.reduction = gamma, reduction
        self.fp16_enabled = False

    def forward(self, input, target):
        return Fn.LogitsFocalFn.apply(input.reshape(-1, input.shape[-1]), target, float(self.gamma), -100)


class CrossEntropyLoss(nn.Module):
    """reference cross_entropy_loss.py:9-83, hard-label branch (the only one Clover's configs reach)."""

    def __init__(self, loss_weight=1.0, class_weight=None):
        super().__init__()
        if class_weight is not None:
            raise NotImplementedError("clover_b200: class_weight is not supported")
        self.loss_weight, self.class_weight = loss_weight, None
        self.fp16_enabled = False

    def forward(self, cls_score, label, **kwargs):
        if cls_score.shape == label.shape:
            raise NotImplementedError("clover_b200: soft-label cross entropy is not supported")
        loss = Fn.LogitsFocalFn.apply(cls_score.reshape(-1, cls_score.shape[-1]), label.reshape(-1).long(), 0.0, -100)
        return loss * self.loss_weight
