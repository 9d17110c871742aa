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
        return Fn.BertEmbedFn.apply(input_ids, self.word_embeddings.weight, self.position_embeddings.weight,
                                    self.token_type_embeddings.weight, self.LayerNorm.weight, self.LayerNorm.bias, self.eps)

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
        ctx = Fn.AttentionFn.apply(qkv, key_mask, B, S, self.heads, self.head_dim, scale)
        o = self.attention.output
        a = Fn.linear(ctx, o.dense.weight, o.dense.bias, residual=h, out_fp32=True)
        a = Fn.layer_norm(a, o.LayerNorm.weight, o.LayerNorm.bias, self.eps)
        f = Fn.mlp(a, self.intermediate.dense.weight, self.intermediate.dense.bias, self.output.dense.weight,
                   self.output.dense.bias, residual=a, out_fp32=True)
        return Fn.layer_norm(f, self.output.LayerNorm.weight, self.output.LayerNorm.bias, self.eps)


class BertEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(cfg) for _ in range(cfg["num_hidden_layers"])])

    def forward_tokens(self, h, attention_mask, B, S):
        """attention_mask (B, S) 1 = attend -> additive (1-m)*-10000 (transformers 4.6.1)."""
        km = ((1.0 - attention_mask.to(torch.float32)) * -10000.0).contiguous()
        if self.training and any(l.attention.self.dropout.p > 0 or l.output.dropout.p > 0 for l in self.layer):
            raise NotImplementedError("clover_b200: BERT dropout > 0 in training mode is not implemented yet; build with "
                                      "hidden_dropout_prob=0 / attention_probs_dropout_prob=0 or call .eval()")
        for layer in self.layer:
            h = layer.forward_tokens(h, km, B, S)
        return h


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
