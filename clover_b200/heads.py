"""Projection / MLM / QA heads on the clover_b200 kernels (same names, constructors and parameter
trees as mmaction/models/heads/{ssl_head,mlm_itm_head,qa_head}.py).  Inputs may be fp32 or bf16;
embeddings and logits come out in fp32 (they feed the @force_fp32 losses)."""
import torch
import torch.nn as nn

from . import functional as Fn
from .text import bert_config

BF16 = torch.bfloat16


def _xavier_init(module):
    """reference init_weights (ssl_head.py:79-88)."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)


def _head_dropout(module, x, p, kind):
    """nn.Dropout(p) of the reference heads in training mode (counter-based mask, clover_b200.rng); identity otherwise."""
    if p and module.training:
        return Fn.dropout(x, float(p), True, out_fp32=x.dtype == torch.float32, kind=kind)
    return x


def _tokens_channels_last(img):
    """(B, C, T, h, w) -> ([B*S, C] fp32 channels-last tokens, B, S); free for clover_b200 Swin outputs."""
    B, C = img.shape[:2]
    S = img[0, 0].numel()
    t = img.reshape(B, C, S).permute(0, 2, 1).reshape(B * S, C)
    return Fn.to_dtype(t.contiguous(), torch.float32), B, S


class NCEHeadForMM(nn.Module):
    """reference ssl_head.py:8-139 (ln=True, text_bn=False, spatial_type='avg', text_agg_type='cls')."""

    def __init__(self, visual_in_channels, text_in_channels, img_hidden_dim, vts_embed_dim, spatial_type="avg",
                 text_agg_type="avg", ln=False, text_bn=False, dropout_ratio=0.1, init_std=0.01, **kwargs):
        super().__init__()
        if not ln or text_bn:
            raise NotImplementedError("clover_b200: BatchNorm1d projection heads are not supported (ln=True, text_bn=False "
                                      "in every shipped Clover config)")
        if spatial_type != "avg" or text_agg_type != "cls":
            raise NotImplementedError("clover_b200: only spatial_type='avg' and text_agg_type='cls' are supported")
        self.vis_in_channels, self.text_in_channels = visual_in_channels, text_in_channels
        self.spatial_type, self.text_agg_type = spatial_type, text_agg_type
        self.dropout_ratio, self.init_std, self.ln = dropout_ratio, init_std, ln
        self.img_hidden_dim, self.vts_embed_dim = img_hidden_dim, vts_embed_dim
        self.fp16_enabled = False
        self.dropout = nn.Dropout(p=dropout_ratio) if dropout_ratio != 0 else None
        self.img_projector = nn.Sequential(nn.Linear(visual_in_channels, img_hidden_dim), nn.LayerNorm(img_hidden_dim),
                                           nn.GELU(), nn.Linear(img_hidden_dim, vts_embed_dim), nn.LayerNorm(vts_embed_dim))
        self.text_projector = nn.Sequential(nn.Linear(text_in_channels, text_in_channels), nn.GELU(),
                                            nn.Linear(text_in_channels, vts_embed_dim))
        self.avg_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
        _xavier_init(self)

    def init_weights(self):
        _xavier_init(self)

    def forward(self, img, text, text_mask=None, token_ids=None):
        return self.forward_vision(img), self.forward_text(text, text_mask, token_ids)

    def forward_vision_tokens(self, tokens, B, S):
        """tokens fp32 [B*S, C] -> fp32 [B, vts]"""
        p = self.img_projector
        x = Fn.MeanTokensFn.apply(tokens, B, S)
        x = _head_dropout(self, x, self.dropout_ratio, "head_mm_vision")           # ssl_head.py:108-109
        x = Fn.linear(Fn.to_dtype(x, BF16), p[0].weight, p[0].bias, out_fp32=True)
        x = Fn.gelu(Fn.layer_norm(x, p[1].weight, p[1].bias, p[1].eps))
        x = Fn.linear(x, p[3].weight, p[3].bias, out_fp32=True)
        return Fn.layer_norm(x, p[4].weight, p[4].bias, p[4].eps, out_fp32=True)

    def forward_vision(self, img):
        tokens, B, S = _tokens_channels_last(img)
        return self.forward_vision_tokens(tokens, B, S)

    def forward_text(self, text, text_mask=None, token_ids=None):
        """text (B, L, H): CLS token -> Linear -> GELU -> Linear (ssl_head.py:131-137)."""
        p = self.text_projector
        cls = Fn.to_dtype(text[:, 0].contiguous(), BF16)
        return Fn.mlp(cls, p[0].weight, p[0].bias, p[2].weight, p[2].bias, out_fp32=True)


class NCEHeadForVision(nn.Module):
    """reference ssl_head.py:142-221 (ln=True).  Accepts (B, C) as well as (B, S, C): the released
    pre-train path feeds a (B, 768) CLS state, for which mean(dim=1) is undefined."""

    def __init__(self, cross_in_channels=768, visual_in_channels=1024, hidden_dim=768, vts_embed_dim=768,
                 dropout_ratio=0.1, ln=False, init_std=0.01, **kwargs):
        super().__init__()
        if not ln:
            raise NotImplementedError("clover_b200: BatchNorm1d projection heads are not supported (ln=True in shipped configs)")
        self.cross_in_channels, self.visual_in_channels = cross_in_channels, visual_in_channels
        self.vts_embed_dim, self.hidden_dim, self.dropout_ratio, self.init_std, self.ln = vts_embed_dim, hidden_dim, dropout_ratio, init_std, ln
        self.dropout = nn.Dropout(p=dropout_ratio) if dropout_ratio != 0 else None
        self.img_fc1 = nn.Linear(visual_in_channels, hidden_dim * 2)
        self.img_bn1 = nn.LayerNorm(hidden_dim * 2)
        self.img_act = nn.GELU()
        self.img_fc2 = nn.Linear(hidden_dim * 2, vts_embed_dim)
        self.img_bn2 = nn.LayerNorm(vts_embed_dim)
        _xavier_init(self)

    def init_weights(self):
        _xavier_init(self)

    def forward(self, img):
        if img.dim() == 3:
            B, S, C = img.shape
            x = Fn.to_dtype(Fn.MeanTokensFn.apply(Fn.to_dtype(img.reshape(B * S, C).contiguous(), torch.float32), B, S), BF16)
        else:
            x = Fn.to_dtype(img.contiguous(), BF16)
        x = _head_dropout(self, x, self.dropout_ratio, "head_vision")              # ssl_head.py:211-212
        x = Fn.linear(x, self.img_fc1.weight, self.img_fc1.bias, out_fp32=True)
        x = Fn.gelu(Fn.layer_norm(x, self.img_bn1.weight, self.img_bn1.bias, self.img_bn1.eps))
        x = Fn.linear(x, self.img_fc2.weight, self.img_fc2.bias, out_fp32=True)
        return Fn.layer_norm(x, self.img_bn2.weight, self.img_bn2.bias, self.img_bn2.eps, out_fp32=True)


class NCEHeadForText(nn.Module):
    """reference ssl_head.py:224-297 (text_bn=False)."""

    def __init__(self, cross_in_channels=768, vts_embed_dim=768, dropout_ratio=0.1, text_bn=False, **kwargs):
        super().__init__()
        if text_bn:
            raise NotImplementedError("clover_b200: text_bn=True is not supported")
        self.cross_in_channels, self.vts_embed_dim, self.dropout_ratio, self.text_bn = cross_in_channels, vts_embed_dim, dropout_ratio, text_bn
        self.fp16_enabled = False
        self.dropout = nn.Dropout(p=dropout_ratio) if dropout_ratio != 0 else None
        self.fc1 = nn.Linear(cross_in_channels, cross_in_channels)
        self.bn = None
        self.act = nn.GELU()
        self.fc2 = nn.Linear(cross_in_channels, vts_embed_dim)
        _xavier_init(self)

    def init_weights(self):
        _xavier_init(self)

    def forward(self, mask_word_feat):
        x = Fn.to_dtype(mask_word_feat.contiguous(), BF16)
        if self.dropout_ratio and self.training:                                   # dropout sits between GELU and fc2 (:292-293)
            h = Fn.linear(x, self.fc1.weight, self.fc1.bias, act="gelu")
            h = _head_dropout(self, h, self.dropout_ratio, "head_text")
            return Fn.linear(h, self.fc2.weight, self.fc2.bias, out_fp32=True)
        return Fn.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, out_fp32=True)


class BertPredictionHeadTransform(nn.Module):
    def __init__(self, hidden_size, eps=1e-12):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=eps)


class BertLMPredictionHead(nn.Module):
    """HF BertLMPredictionHead (reference mlm_itm_head.py:25-41).  The decoder is an independent
    (vocab, hidden) matrix, not tied to the text encoder's embeddings."""

    def __init__(self, hidden_size, vocab_size):
        super().__init__()
        self.transform = BertPredictionHeadTransform(hidden_size)
        self.decoder = nn.Linear(hidden_size, vocab_size, bias=True)
        for m in (self.transform.dense, self.decoder):
            m.weight.data.normal_(0.0, 0.02)
            m.bias.data.zero_()


class MLMHead(nn.Module):
    """reference mlm_itm_head.py:43-52.  ``forward`` returns fp32 logits for every row (reference
    contract); the recogniser uses :meth:`focal_loss`, which never materialises unselected rows' use."""

    def __init__(self, hidden_size, vocab_size, **kwargs):
        super().__init__()
        self.predictions = BertLMPredictionHead(hidden_size, vocab_size)
        self.fp16_enabled = False
        self.vocab_size = vocab_size

    def transform(self, h2d):
        t = self.predictions.transform
        x = Fn.linear(Fn.to_dtype(h2d.contiguous(), BF16), t.dense.weight, t.dense.bias, act="gelu", out_fp32=True)
        return Fn.layer_norm(x, t.LayerNorm.weight, t.LayerNorm.bias, t.LayerNorm.eps)

    def forward(self, sequence_output):
        shp = sequence_output.shape
        x = self.transform(sequence_output.reshape(-1, shp[-1]))
        d = self.predictions.decoder
        V = d.weight.shape[0]
        y = _linear_any_n(x, d)                      # the shipped vocabulary (30522) is not a multiple of 8: padded weight cache
        return y.reshape(*shp[:-1], V)

    def focal_loss(self, sequence_output, mlm_label, gamma=2.0, ignore_index=-100, rows=None):
        """Row selection (pretrain.py:137-139) + transform + decoder + SoftmaxFocalLossMultiClass.  The reference runs the
        head on all B*L rows and indexes the (B*L, 30522) fp32 logits afterwards; here the masked rows are selected FIRST
        (`rows`: their int64 indices, computed once per step by the recogniser), so the transform, the vocabulary GEMM, its two
        backward GEMMs and the logits tensor shrink to the ~20-30 % of the tokens that carry a label -- the same arithmetic on
        the same rows (LayerNorm / GELU / dense are row-wise).  rows=None keeps every row (labels == ignore_index are skipped
        inside the loss kernel)."""
        shp = sequence_output.shape
        h = sequence_output.reshape(-1, shp[-1])
        lab = mlm_label.reshape(-1)
        if rows is not None and rows.numel() > 0:
            h = Fn.GatherRowsFn.apply(h.contiguous(), rows)
            lab = lab.index_select(0, rows)
            pad = (-h.shape[0]) % 8                      # the GEMM tiles want a multiple of 8 rows: pad with ignored rows
            if pad:
                h = torch.cat([h, h.new_zeros(pad, h.shape[1])], 0)
                lab = torch.cat([lab, lab.new_full((pad,), ignore_index)], 0)
        x = self.transform(h)
        d = self.predictions.decoder
        return Fn.VocabFocalFn.apply(x, d.weight, d.bias, lab, gamma, ignore_index)


class ITMHead(nn.Module):
    """reference mlm_itm_head.py:55-97: Dropout(0.1) -> Linear(hidden, hidden) -> Tanh -> Linear(hidden, 2)
    (the dropout rate is hard-coded; ``dropout_ratio`` / ``finetune`` of the shipped config are swallowed by **kwargs)."""

    def __init__(self, hidden_dim=768, **kwargs):
        super().__init__()
        self.itm_projector = nn.Sequential(nn.Dropout(p=0.1), nn.Linear(hidden_dim, hidden_dim), nn.Tanh(),
                                           nn.Linear(hidden_dim, 2))
        self.fp16_enabled = False
        _xavier_init(self)

    def init_weights(self):
        _xavier_init(self)

    def forward(self, cls_feature):
        c = self.itm_projector
        x = _head_dropout(self, Fn.to_dtype(cls_feature.contiguous(), BF16), c[0].p, "head_itm")
        x = Fn.tanh(Fn.linear(x, c[1].weight, c[1].bias))
        return _linear_any_n(x, c[3])


class QA_OE_Head(nn.Module):
    """reference qa_head.py:42-88."""

    def __init__(self, hidden_dim=768, dropout_ratio=0.5, num_labels=None, **kwargs):
        super().__init__()
        self.num_labels, self.dropout_ratio = num_labels, dropout_ratio
        self.vqa_classifier = nn.Sequential(nn.Dropout(dropout_ratio), nn.Linear(hidden_dim, hidden_dim // 2),
                                            nn.LayerNorm(hidden_dim // 2), nn.GELU(), nn.Linear(hidden_dim // 2, num_labels))
        _xavier_init(self)

    def init_weights(self):
        _xavier_init(self)

    def forward(self, cls_feature):
        c = self.vqa_classifier
        x = _head_dropout(self, Fn.to_dtype(cls_feature.contiguous(), BF16), self.dropout_ratio, "head_qa_oe")   # qa_head.py:60
        x = Fn.linear(x, c[1].weight, c[1].bias, out_fp32=True)
        x = Fn.gelu(Fn.layer_norm(x, c[2].weight, c[2].bias, c[2].eps))
        return _linear_any_n(x, c[4])


class QA_MC_head(nn.Module):
    """reference qa_head.py:7-39."""

    def __init__(self, hidden_dim, dropout_ratio=0.1):
        super().__init__()
        self.dropout_ratio = dropout_ratio
        self.mc_vqa_classifier = nn.Sequential(nn.Dropout(dropout_ratio), nn.Linear(hidden_dim, 256), nn.LayerNorm(256),
                                               nn.GELU(), nn.Linear(256, 1))
        _xavier_init(self)

    def init_weights(self):
        _xavier_init(self)

    def forward(self, x):
        c = self.mc_vqa_classifier
        x = _head_dropout(self, Fn.to_dtype(x.contiguous(), BF16), self.dropout_ratio, "head_qa_mc")             # qa_head.py:12
        x = Fn.linear(x, c[1].weight, c[1].bias, out_fp32=True)
        x = Fn.gelu(Fn.layer_norm(x, c[2].weight, c[2].bias, c[2].eps))
        return _linear_any_n(x, c[4])


def _linear_any_n(x, lin):
    """Linear whose out_features is not a multiple of 8 (QA logits): pad the weight rows in the bf16 cache."""
    N = lin.weight.shape[0]
    if N % 8 == 0:
        return Fn.linear(x, lin.weight, lin.bias, out_fp32=True)
    return Fn.PaddedLinearFn.apply(x, lin.weight, lin.bias)
