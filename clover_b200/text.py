"""BERT text encoder on the clover_b200 kernels.

``BertFromPretrained`` keeps the reference's name / constructor / return structure
(mmaction/models/backbones/bert_from_hugface.py) and the HuggingFace ``BertModel`` parameter tree
(``bert.embeddings.*``, ``bert.encoder.layer.{i}.*``, ``bert.pooler.dense.*``) so checkpoints are
interchangeable.  The arithmetic follows transformers 4.6.1 (pinned by the reference's install.sh:27;
SURVEY.md App. E1): post-LN encoder, additive (1-m)*-10000 key mask, erf-GELU, LN eps 1e-12.
No ``from_pretrained`` download happens (no network): weights are HF-style random init unless a
state dict is loaded.
"""
import torch
import torch.nn as nn

from . import functional as Fn
from . import ops

BERT_BASE = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, initializer_range=0.02)


def bert_config(**overrides):
    cfg = dict(BERT_BASE)
    cfg.update({k: v for k, v in overrides.items() if k in cfg})
    return cfg


class ModelOutput(dict):
    """dict with attribute access, like HF's ModelOutput (the reference indexes by key)."""
    __getattr__ = dict.get


class BertEmbeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        H = cfg["hidden_size"]
        self.word_embeddings = nn.Embedding(cfg["vocab_size"], H, padding_idx=0)
        self.position_embeddings = nn.Embedding(cfg["max_position_embeddings"], H)
        self.token_type_embeddings = nn.Embedding(cfg["type_vocab_size"], H)
        self.LayerNorm = nn.LayerNorm(H, eps=cfg["layer_norm_eps"])
        self.dropout = nn.Dropout(cfg["hidden_dropout_prob"])
        self.eps = cfg["layer_norm_eps"]
        self.register_buffer("position_ids", torch.arange(cfg["max_position_embeddings"]).expand((1, -1)), persistent=False)

    def forward(self, input_ids):
        """(B, L) int64 -> bf16 [B*L, H]"""
        if input_ids.shape[1] > self.position_embeddings.weight.shape[0]:
            raise ValueError("sequence longer than max_position_embeddings")
        h = Fn.BertEmbedFn.apply(input_ids, self.word_embeddings.weight, self.position_embeddings.weight,
                                 self.token_type_embeddings.weight, self.LayerNorm.weight, self.LayerNorm.bias, self.eps)
        if self.training and self.dropout.p > 0:            # HF BertEmbeddings: dropout(LayerNorm(sum of embeddings))
            h = Fn.dropout(h, self.dropout.p, True, kind="bert_embeddings")
        return h

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        state_dict.pop(prefix + "position_ids", None)      # persistent buffer in transformers 4.6.1 checkpoints
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class _SelfAttention(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        H = cfg["hidden_size"]
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)
        self.dropout = nn.Dropout(cfg["attention_probs_dropout_prob"])


class _SelfOutput(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        H = cfg["hidden_size"]
        self.dense = nn.Linear(H, H)
        self.LayerNorm = nn.LayerNorm(H, eps=cfg["layer_norm_eps"])
        self.dropout = nn.Dropout(cfg["hidden_dropout_prob"])


class _Attention(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.self = _SelfAttention(cfg)
        self.output = _SelfOutput(cfg)


class _Intermediate(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg["hidden_size"], cfg["intermediate_size"])


class _Output(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg["intermediate_size"], cfg["hidden_size"])
        self.LayerNorm = nn.LayerNorm(cfg["hidden_size"], eps=cfg["layer_norm_eps"])
        self.dropout = nn.Dropout(cfg["hidden_dropout_prob"])


class BertLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.attention = _Attention(cfg)
        self.intermediate = _Intermediate(cfg)
        self.output = _Output(cfg)
        self.heads = cfg["num_attention_heads"]
        self.head_dim = cfg["hidden_size"] // self.heads
        self.eps = cfg["layer_norm_eps"]
        if self.head_dim not in (32, 64):
            raise ValueError(f"clover_b200 attention kernels support head_dim 32 or 64, got {self.head_dim}")

    def forward_tokens(self, h, key_mask, B, S):
        """h bf16 [B*S, H]; key_mask fp32 additive [B, S]."""
        s = self.attention.self
        scale = self.head_dim ** -0.5
        qkv = Fn.QkvLinearFn.apply(h, s.query.weight, s.query.bias, s.key.weight, s.key.bias, s.value.weight, s.value.bias,
                                   scale, self)
        o = self.attention.output
        p_attn = s.dropout.p if self.training else 0.0       # HF 4.6.1: dropout on the attention probabilities,
        p_ao = o.dropout.p if self.training else 0.0         # on the two dense outputs before their residual adds
        p_out = self.output.dropout.p if self.training else 0.0
        ctx = Fn.AttentionFn.apply(qkv, key_mask, B, S, self.heads, self.head_dim, scale, float(p_attn))
        if p_ao > 0:
            a = Fn.dropout(Fn.linear(ctx, o.dense.weight, o.dense.bias), p_ao, True, residual=h, out_fp32=True,
                           kind="bert_self_output")
        else:
            a = Fn.linear(ctx, o.dense.weight, o.dense.bias, residual=h, out_fp32=True)
        a = Fn.layer_norm(a, o.LayerNorm.weight, o.LayerNorm.bias, self.eps)
        m = (a, self.intermediate.dense.weight, self.intermediate.dense.bias, self.output.dense.weight, self.output.dense.bias)
        if p_out > 0:
            f = Fn.dropout(Fn.mlp(*m), p_out, True, residual=a, out_fp32=True, kind="bert_output")
        else:
            f = Fn.mlp(*m, residual=a, out_fp32=True)
        return Fn.layer_norm(f, self.output.LayerNorm.weight, self.output.LayerNorm.bias, self.eps)


def _layer_probs_mean(layer, h, key_mask, B, S):
    """Head-mean attention probabilities (B, S, S) fp32 of one layer on input h: what HF's output_attentions=True gives
    the reference at finetune.py:192 (`attentions[-1].mean(dim=1)`).  Evaluation only."""
    s = layer.attention.self
    scale = layer.head_dim ** -0.5
    with torch.no_grad():
        qkv = Fn.QkvLinearFn.apply(h, s.query.weight, s.query.bias, s.key.weight, s.key.bias, s.value.weight, s.value.bias,
                                   scale, layer)
        return ops.attention_probs_mean(qkv, B, S, layer.heads, layer.head_dim, key_mask=key_mask)


class BertEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(cfg) for _ in range(cfg["num_hidden_layers"])])

    def forward_tokens(self, h, attention_mask, B, S, want_last_probs=False):
        """attention_mask (B, S) 1 = attend -> additive (1-m)*-10000 (transformers 4.6.1).
        want_last_probs: also return the last layer's head-mean attention probabilities (B, S, S)."""
        km = ((1.0 - attention_mask.to(torch.float32)) * -10000.0).contiguous()
        probs = None
        for i, layer in enumerate(self.layer):
            if want_last_probs and i == len(self.layer) - 1:
                probs = _layer_probs_mean(layer, h, km, B, S)
            h = layer.forward_tokens(h, km, B, S)
        return (h, probs) if want_last_probs else h


class _Pooler(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg["hidden_size"], cfg["hidden_size"])
        self.activation = nn.Tanh()


class BertModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.embeddings = BertEmbeddings(cfg)
        self.encoder = BertEncoder(cfg)
        self.pooler = _Pooler(cfg)       # parameters kept for checkpoint parity; never used by Clover (no grad)
        self.apply(self._init)

    def _init(self, m):
        std = self.cfg["initializer_range"]
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(0.0, std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(0.0, std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)

    def forward(self, input_ids=None, attention_mask=None, **kwargs):
        B, L = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        h = self.embeddings(input_ids)
        h = self.encoder.forward_tokens(h, attention_mask, B, L)
        return ModelOutput(last_hidden_state=h.view(B, L, -1), pooler_output=None)


class BertFromPretrained(nn.Module):
    """reference bert_from_hugface.py:7-32.  Extra keyword arguments matching BertConfig fields
    (hidden_size, num_attention_heads, ...) override bert-base-uncased for tests."""

    def __init__(self, pretrained_model="bert-base-uncased", layer_norm_eps=1e-12, num_hidden_layers=12, **kwargs):
        super().__init__()
        cfg = bert_config(layer_norm_eps=layer_norm_eps, num_hidden_layers=num_hidden_layers, **kwargs)
        self.bert = BertModel(cfg)

    def forward(self, token_ids=None, input_mask=None, **kwargs):
        return self.bert(input_ids=token_ids, attention_mask=input_mask)
