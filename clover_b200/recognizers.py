"""Recognisers: the orchestration of the Clover hot path on the clover_b200 modules.

``CloverPretrain`` / ``CloverFinetune`` keep the reference's constructors, loss-dict keys and
``forward`` / ``train_step`` / ``_parse_losses`` contracts
(mmaction/models/recognizers/{base,multimodal_transformer_pretrain,multimodal_transformer_finetune}.py).
The order of the encoder passes follows forward_train line by line; what changes is how data moves:
tokens stay channels-last between modules, the six alignment embeddings cross ranks in ONE
all-gather, and the MLM decoder is fused with its focal loss.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as Fn
from .gather import gather_stacked
from .registry import build_backbone, build_head, build_loss


class BaseRecognizer(nn.Module):
    """The subset of reference recognizers/base.py that Clover's recognisers use."""

    def __init__(self, backbone, cls_head=None, neck=None, freeze_stage=None, freeze_except=[], train_cfg=None,
                 test_cfg=None):
        super().__init__()
        self.backbone_from = "mmaction2"
        self.backbone = build_backbone(backbone)
        if neck is not None or cls_head is not None:
            raise NotImplementedError("clover_b200: neck / cls_head are not part of the Clover hot path")
        self.cls_head = None
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.aux_info = list(train_cfg["aux_info"]) if train_cfg is not None and "aux_info" in train_cfg else []
        self.blending = None
        self.fp16_enabled = False
        self.init_weights()
        if freeze_stage is not None:
            self._freeze(freeze_stage, freeze_except)

    def init_weights(self):
        self.backbone.init_weights()

    def _freeze(self, freeze_stage, freeze_except):
        """reference base.py:138-163: freeze every module whose name contains an entry of freeze_stage."""
        freeze_norm = "norm_layer" not in freeze_except
        for n, m in self.named_modules():
            if any(en in n for en in freeze_except):
                continue
            for fn in freeze_stage:
                if fn in n:
                    if isinstance(m, (nn.LayerNorm, nn.GroupNorm, nn.modules.batchnorm._BatchNorm)):
                        if not freeze_norm:
                            break
                        m.eval()
                    for p in m.parameters():
                        p.requires_grad = False
                    break

    @staticmethod
    def _parse_losses(losses):
        """reference base.py:254-288: loss = unweighted sum of the '*loss*' entries; every logged scalar is
        averaged over ranks.  The per-scalar all-reduce + .item() of the reference is one packed all-reduce."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f"{name} is not a tensor or list of tensors")
        loss = sum(v for k, v in log_vars.items() if "loss" in k)
        log_vars["loss"] = loss
        packed = torch.stack([v.detach().float().reshape(()) for v in log_vars.values()])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            packed = packed / dist.get_world_size()
            dist.all_reduce(packed)
        vals = packed.tolist()
        for k, v in zip(list(log_vars.keys()), vals):
            log_vars[k] = v
        return loss, log_vars

    def forward(self, imgs=None, label=None, return_loss=True, **kwargs):
        if return_loss:
            if label is None:
                raise ValueError("Label should not be None.")
            return self.forward_train(imgs, label, **kwargs)
        return self.forward_test(imgs, **kwargs)

    def train_step(self, data_batch, optimizer=None, **kwargs):
        imgs, label = data_batch["imgs"], data_batch["label"]
        aux = {k: data_batch[k] for k in self.aux_info}
        losses = self(imgs, label, return_loss=True, **aux)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(next(iter(data_batch.values()))))

    val_step = train_step


def _flat(x):
    return x.reshape((-1,) + tuple(x.shape[2:]))


class CloverPretrain(BaseRecognizer):
    """reference multimodal_transformer_pretrain.py:11-230."""

    def __init__(self, mm_backbone, text_backbone=None, freeze_text_backbone=None, freeze_dvae_backbone=None,
                 loss_type=None, ssl_loss=None, ssl_head=None, mlm_head=None, mlm_loss=None, mlm_ssl_head=None,
                 symmetry_rank=False, separate_test=False, from_scratch=False, use_Cmask=True, text_vocab_size=30522,
                 **kwargs):
        super().__init__(**kwargs)
        self.multimodal_backbone = build_backbone(mm_backbone)
        self.text_backbone = build_backbone(text_backbone)
        self.text_vocab_size = text_vocab_size
        self.loss_func = build_loss(loss_type) if loss_type is not None else None
        self.use_Cmask = use_Cmask
        self.mlm_head = build_head(mlm_head) if mlm_head is not None else None
        self.mlm_ssl_V_head = self.mlm_ssl_T_head = None
        if mlm_ssl_head is not None:
            self.mlm_ssl_V_head = build_head(mlm_ssl_head["V"]) if mlm_ssl_head.get("V") else None
            self.mlm_ssl_T_head = build_head(mlm_ssl_head["T"]) if mlm_ssl_head.get("T") else None
        self.mlm_loss_func = build_loss(mlm_loss) if mlm_loss is not None else None
        self.symmetry_rank, self.from_scratch, self.separate_test = symmetry_rank, from_scratch, separate_test
        if ssl_head is not None:
            self.ssl_head_name = ssl_head["type"]
            self.ssl_head = build_head(ssl_head)
            self.ssl_loss = build_loss(ssl_loss)
        if freeze_dvae_backbone is not None:
            self._freeze(freeze_dvae_backbone, [])
        if freeze_text_backbone is not None:
            self._freeze(freeze_text_backbone, [])

    def extract_visual_feat(self, imgs, mask=None):
        return self.backbone(imgs, mask)

    def forward_train(self, imgs, label, token_ids=None, segment_ids=None, input_mask=None, mlm_label=None,
                      dvae_imgs=None, v_token_mask=None, hog_features=None, img_metas=None, **kwargs):
        imgs = _flat(imgs)                                                         # :81
        # rows that carry an MLM label (:137-139): their indices are needed to shrink the MLM head to those rows.  The one
        # device -> host size read happens HERE, at the very start of the step (the labels are an input), not in the middle
        # of the forward where it would drain the launch queue.
        mlm_rows = None
        if mlm_label is not None and self.select_mlm_rows:
            mlm_rows = torch.nonzero(mlm_label.reshape(-1) != -100).squeeze(1)
        Fn.zero_arena_begin(imgs.device)
        if self.from_scratch:
            imgs = imgs / 255.0
        token_ids, text_mask = _flat(token_ids), _flat(input_mask)                 # :85-86
        if mlm_label is not None:
            mlm_label = _flat(mlm_label)
        if not (hasattr(self, "ssl_head") and self.mlm_ssl_V_head is not None and self.symmetry_rank and self.use_Cmask
                and mlm_label is not None and self.mlm_head is not None):
            raise NotImplementedError("clover_b200.CloverPretrain implements the shipped tri-modal configuration "
                                      "(ssl_head + mlm_head + mlm_ssl_head V/T, symmetry_rank, use_Cmask)")
        Bt, L = token_ids.shape
        H = self.multimodal_backbone.hidden_size
        if self.batch_passes and v_token_mask is not None and imgs.shape[0] == Bt:
            return self._forward_train_batched(imgs, token_ids, text_mask, mlm_label, v_token_mask, Bt, L, H, mlm_rows)
        # ---- clean video + clean text ---------------------------------------------------- :91-102
        v_tok, (B, T, h, w) = self.backbone.forward_tokens(imgs)                    # fp32 [B*T*hw, C]
        S = h * w
        ids_clean = torch.where(mlm_label == -100, token_ids, mlm_label)
        T_e = self.text_backbone(ids_clean, text_mask)["last_hidden_state"]        # bf16 (B, L, H)
        v_emb = self.ssl_head.forward_vision_tokens(v_tok, B, T * S)
        t_emb = self.ssl_head.forward_text(T_e)
        # ---- masked text, masked video ---------------------------------------------------- :110-115
        T_m = self.text_backbone(token_ids, text_mask)["last_hidden_state"]
        vm_tok, _ = self.backbone.forward_tokens(imgs, v_token_mask)
        # ---- fusion passes ----------------------------------------------------------------- :117-121
        v_f, _ = self.multimodal_backbone.forward_tokens(vm_tok, B, T, S, T_e, text_mask)
        t_f, _ = self.multimodal_backbone.forward_tokens(v_tok, B, T, S, T_m, text_mask)
        vs = self.multimodal_backbone.v_seq_len(T, S)                               # text tokens start here (cross_transformer.py:111-117)
        t_last = t_f[:, vs:]                                                        # (B, L, H)
        losses = dict()
        # ---- MLM: decoder + row selection + focal loss fused ------------------------------- :129-143
        gamma = getattr(self.mlm_loss_func, "gamma", 0.0) if self.mlm_loss_func is not None else 0.0
        losses["mlm_loss"] = self.mlm_head.focal_loss(t_last.reshape(B * L, H), mlm_label.reshape(-1), gamma=gamma, rows=mlm_rows)
        # ---- tri-modal alignment ------------------------------------------------------------ :147-169
        m_vmf = self.mlm_ssl_V_head(v_f[:, vs])                                     # fused text-CLS slot, (B, H)
        tm_emb = self.ssl_head.forward_text(T_m)
        m_tmf = self.mlm_ssl_T_head(t_last[:, 0])
        vm_emb = self.ssl_head.forward_vision_tokens(vm_tok, B, T * S)
        g_v, g_t, g_tm, g_vmf, g_vm, g_tmf = gather_stacked([v_emb, t_emb, tm_emb, m_vmf, vm_emb, m_tmf])
        losses.update(self.ssl_loss.forward_gathered(g_v, g_t, g_tm, g_vmf))
        l2 = self.ssl_loss.forward_gathered(g_t, g_v, g_vm, g_tmf)
        losses["v_nce_loss"] = l2.pop("nce_loss")
        if self.ssl_loss.use_rank:
            losses["rank_v_vm_loss"] = l2.pop("rank_t_tm_loss")
        return losses

    # The reference runs every encoder twice per step (clean / masked video :91,:113; clean / masked text :96,:110;
    # two fusion passes :117-121).  No layer on the path couples samples (LayerNorm only; ln=True / text_bn=False
    # heads), so the two passes are one pass over a doubled batch: [masked ; clean] clips (an all-zero token mask
    # leaves the patch embedding untouched, swin_transformer_3d.py:222-230), [clean ; masked] captions, and the
    # fusion pairs (masked video, clean text) ; (clean video, masked text) line up without any reshuffle.  Same
    # arithmetic per sample, half the launches, and every parameter gets ONE gradient instead of two accumulated.
    batch_passes = True
    # run the MLM head only on the rows that carry a label (False: on all B*L rows like the reference, no host read)
    select_mlm_rows = True

    def _forward_train_batched(self, imgs, token_ids, text_mask, mlm_label, v_token_mask, B, L, H, mlm_rows=None):
        pe = getattr(self.backbone, "patch_embed", None)
        if pe is not None and hasattr(pe, "pair_supported") and pe.pair_supported(imgs, v_token_mask):
            # both passes see the same clips: one patch gather + projection, two LayerNorm / blend evaluations
            tok2, (B2, T, h, w) = self.backbone.forward_tokens(imgs, v_token_mask, pair=True)
        else:
            imgs2 = torch.cat([imgs, imgs], 0)
            vmask2 = torch.cat([v_token_mask, torch.zeros_like(v_token_mask)], 0)
            tok2, (B2, T, h, w) = self.backbone.forward_tokens(imgs2, vmask2)       # [masked ; clean] fp32 [2B*T*hw, C]
        S = h * w
        ids_clean = torch.where(mlm_label == -100, token_ids, mlm_label)
        ids2 = torch.cat([ids_clean, token_ids], 0)                                 # [clean ; masked]
        tmask2 = torch.cat([text_mask, text_mask], 0)
        T2 = self.text_backbone(ids2, tmask2)["last_hidden_state"]                  # bf16 (2B, L, H)
        vemb2 = self.ssl_head.forward_vision_tokens(tok2, B2, T * S)                # (2B, E): [vm ; v]
        temb2 = self.ssl_head.forward_text(T2)                                      # (2B, E): [t ; tm]
        vm_emb, v_emb = vemb2.split(B)
        t_emb, tm_emb = temb2.split(B)
        f2, _ = self.multimodal_backbone.forward_tokens(tok2, B2, T, S, T2, tmask2)  # (vm, T_e) ; (v, T_m)
        vs = self.multimodal_backbone.v_seq_len(T, S)                               # text tokens start here (cross_transformer.py:111-117)
        t_last = f2[B:, vs:]                                                        # (B, L, H) of the masked-text pass
        losses = dict()
        gamma = getattr(self.mlm_loss_func, "gamma", 0.0) if self.mlm_loss_func is not None else 0.0
        losses["mlm_loss"] = self.mlm_head.focal_loss(t_last.reshape(B * L, H), mlm_label.reshape(-1), gamma=gamma, rows=mlm_rows)
        m_vmf = self.mlm_ssl_V_head(f2[:B, vs])                                     # fused text-CLS slot of the masked-video pass
        m_tmf = self.mlm_ssl_T_head(t_last[:, 0])
        g_v, g_t, g_tm, g_vmf, g_vm, g_tmf = gather_stacked([v_emb, t_emb, tm_emb, m_vmf, vm_emb, m_tmf])
        losses.update(self.ssl_loss.forward_gathered(g_v, g_t, g_tm, g_vmf))
        l2 = self.ssl_loss.forward_gathered(g_t, g_v, g_vm, g_tmf)
        losses["v_nce_loss"] = l2.pop("nce_loss")
        if self.ssl_loss.use_rank:
            losses["rank_v_vm_loss"] = l2.pop("rank_t_tm_loss")
        return losses

    def forward_test(self, imgs, token_ids=None, segment_ids=None, input_mask=None, **kwargs):
        """reference :197-223 (separate_test branch: retrieval embeddings)."""
        imgs = _flat(imgs)
        if self.from_scratch:
            imgs = imgs / 255.0
        if not self.separate_test:
            raise NotImplementedError("clover_b200: only separate_test=True (retrieval embeddings) is supported")
        v_tok, (B, T, h, w) = self.backbone.forward_tokens(imgs)
        C = v_tok.shape[1]
        B_text = token_ids.shape[0]
        if B_text != B:                                                             # average clips of a video
            v_tok = v_tok.view(B_text, -1, T * h * w, C).mean(dim=1).reshape(-1, C)
            B = B_text
        token_ids, input_mask = _flat(token_ids), _flat(input_mask)
        T_e = self.text_backbone(token_ids, input_mask)["last_hidden_state"]
        return self.ssl_head.forward_vision_tokens(v_tok.contiguous(), B, T * h * w), self.ssl_head.forward_text(T_e)

    def forward_gradcam(self, imgs, token_ids=None, input_mask=None):
        return self.forward_test(imgs, token_ids, input_mask)


class CloverFinetune(BaseRecognizer):
    """reference multimodal_transformer_finetune.py:9-203: tasks 'retrieval', 'video_qa' and 'FIB', every answer-feature
    branch of :98-118 (answer_mask / answer_cls with or without itm_head / all-cls + itm_head, qa_head or raw ITM logit)."""

    def __init__(self, mm_backbone, text_backbone=None, freeze_text_backbone=None, loss_type=None, task=None,
                 ssl_head=None, itm_head=None, answer_mask=False, answer_cls=False, qa_head=None, from_scratch=False,
                 text_vocab_size=30522, separate_test=False, **kwargs):
        super().__init__(**kwargs)
        self.multimodal_backbone = build_backbone(mm_backbone)
        self.text_backbone = build_backbone(text_backbone)
        self.text_vocab_size, self.from_scratch, self.separate_test, self.task = text_vocab_size, from_scratch, separate_test, task
        if task == "retrieval":
            self.ssl_head = build_head(ssl_head)
            self.loss_func = build_loss(loss_type)
        elif task in ("video_qa", "FIB"):
            self.answer_mask, self.answer_cls = answer_mask, answer_cls
            self.itm_head = build_head(itm_head) if itm_head is not None else None
            self.qa_head = build_head(qa_head) if qa_head is not None else None
            if self.itm_head is None and self.qa_head is None:
                raise ValueError("CloverFinetune(task=%r) needs a qa_head or an itm_head" % (task,))
            if not answer_mask and not answer_cls and self.itm_head is None:
                raise ValueError("answer_mask=False, answer_cls=False reads the all-cls slot through itm_head (:110-112)")
            self.loss_func = build_loss(loss_type)
            self.loss_type = loss_type["type"]
        else:
            raise NotImplementedError("must have head to do downstream finetuning")

    def extract_visual_feat(self, imgs):
        return self.backbone(imgs)

    def _encode(self, imgs, token_ids, input_mask):
        imgs = _flat(imgs)
        if self.training:
            Fn.zero_arena_begin(imgs.device)
        if self.from_scratch:
            imgs = imgs / 255.0
        B_text = token_ids.shape[0]
        token_ids, input_mask = _flat(token_ids), _flat(input_mask)
        v_tok, (B, T, h, w) = self.backbone.forward_tokens(imgs)
        C = v_tok.shape[1]
        if B_text != B:
            v_tok = v_tok.view(B_text, -1, T * h * w, C).mean(dim=1).reshape(-1, C).contiguous()
            B = B_text
        T_e = self.text_backbone(token_ids, input_mask)["last_hidden_state"]
        return v_tok, (B, T, h * w), T_e, token_ids, input_mask

    def _qa_logits(self, v_tok, B, T, S, T_e, token_ids, input_mask, test=False):
        """:88-118 (train) / :158-188 (test).  Returns (final_output, head-mean attention or None)."""
        if hasattr(self.qa_head, "num_labels"):                                     # :90-92
            n, Bq = self.qa_head.num_labels, B
        else:                                                                       # :93-95 every candidate sees the clip
            n = T_e.shape[0] // B
            v_tok = v_tok.view(B, 1, T * S, -1).expand(-1, n, -1, -1).reshape(B * n * T * S, -1).contiguous()
            Bq = B * n
        mm = self.multimodal_backbone
        res = mm.forward_tokens(v_tok, Bq, T, S, T_e, input_mask, want_last_probs=test)
        out = res[0]
        vs = mm.v_seq_len(T, S)
        if self.answer_mask:                                                        # :98-100 the [MASK] (id 103) positions
            bi, li = torch.where(token_ids.reshape(Bq, -1) == 103)
            feat = out[bi, vs + li]
        elif self.answer_cls:                                                       # :101-108
            # cls_last_hidden_state.squeeze() when the encoder has an all-cls token, else t_last_hidden_state[:, 0]
            feat = out[:, vs - 1] if mm.all_cls_token is not None else out[:, vs]
            if self.itm_head is not None:
                feat = self.itm_head(feat)
        else:                                                                       # :110-112
            feat = self.itm_head(out[:, 0])
        if self.qa_head is not None:
            final = self.qa_head(feat).reshape(-1, n)                               # :115
        elif test:
            final = torch.softmax(feat.float(), dim=-1)[:, 1].reshape(-1, n)        # :187-188
        else:
            final = feat[:, 1]                                                      # :118
        return final, (res[2] if test else None)

    def forward_train(self, imgs, label, token_ids=None, segment_ids=None, input_mask=None, ans_ids=None, ans_mask=None,
                      **kwargs):
        v_tok, (B, T, S), T_e, token_ids, input_mask = self._encode(imgs, token_ids, input_mask)
        losses = dict()
        if self.task == "retrieval":
            v_emb = self.ssl_head.forward_vision_tokens(v_tok, B, T * S)
            t_emb = self.ssl_head.forward_text(T_e)
            losses["retrieval_nce_loss"] = self.loss_func(v_emb, t_emb)
        else:
            logits, _ = self._qa_logits(v_tok, B, T, S, T_e, token_ids, input_mask)
            losses["qa_loss"] = self.loss_func(logits, label.view(-1))
        return losses

    def forward_test(self, imgs, token_ids=None, segment_ids=None, input_mask=None, ans_ids=None, ans_mask=None, **kwargs):
        v_tok, (B, T, S), T_e, token_ids, input_mask = self._encode(imgs, token_ids, input_mask)
        if self.separate_test or self.task == "retrieval":                          # :152-154
            return self.ssl_head.forward_vision_tokens(v_tok, B, T * S), self.ssl_head.forward_text(T_e)
        # reference :188-192: {'result': fp32 logits, 'attention': last fusion layer's head-mean attention probabilities}
        logits, attn = self._qa_logits(v_tok, B, T, S, T_e, token_ids, input_mask, test=True)
        return {"result": logits.to(torch.float32), "attention": attn}

    def forward_gradcam(self, imgs, token_ids=None, input_mask=None):
        return self.forward_test(imgs, token_ids, input_mask)
