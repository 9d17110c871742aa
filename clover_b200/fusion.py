"""Cross-modal fusion encoder on the clover_b200 kernels.

Same name / constructor / outputs as the reference's CrossModalTransformerFromPretrained
(mmaction/models/backbones/cross_transformer.py:11-141): fc_in, spatial + temporal + token-type
embeddings and LayerNorm on the video tokens, concatenation with the text states, an additive
-10000 padding mask, N BERT layers, sliced outputs.  Parameter tree: fc_in, vis_space_pos,
vis_tempor_pos, token_type_embeddings, norm, bert_embedding.*, bert_encoder.layer.{i}.*.
"""
import torch
import torch.nn as nn

from . import functional as Fn
from .text import BertEmbeddings, BertEncoder, ModelOutput, bert_config


class CrossModalTransformerFromPretrained(nn.Module):
    def __init__(self, pretrained_model="bert-base-uncased", img_in_size=768, hidden_size=768, num_frames=4,
                 spacial_tokens=7 * 7, token_types=2, num_hidden_layers=12, layer_norm_eps=1e-12, word_pos_start=False,
                 use_prompt=False, use_text_cls=False, return_mask=False, **kwargs):
        super().__init__()
        if use_prompt and use_text_cls:
            raise ValueError("use_prompt=True needs use_text_cls=False (the reference reads all_cls_token, cross_transformer.py:100)")
        cfg = bert_config(hidden_size=hidden_size, num_hidden_layers=num_hidden_layers, layer_norm_eps=layer_norm_eps, **kwargs)
        self.bert_embedding = BertEmbeddings(cfg)   # only used when text_input_embeds is None (never by the recognisers)
        self.bert_encoder = BertEncoder(cfg)
        self.use_prompt = use_prompt
        if not use_text_cls:                                  # :31-36 (the class default; configs/exp_local/finetune_lsmdc_FIB.py)
            self.all_cls_token = nn.Parameter(torch.zeros(1, 1, hidden_size))
            nn.init.trunc_normal_(self.all_cls_token, mean=0.0, std=0.02)
            if use_prompt:
                self.prompt_token = nn.Parameter(torch.zeros(1, 4, hidden_size))
                nn.init.trunc_normal_(self.prompt_token, mean=0.0, std=0.02)
        else:
            self.all_cls_token = None
        self.vis_space_pos = nn.Parameter(0.02 * torch.randn(1, 1, spacial_tokens, hidden_size))
        self.vis_tempor_pos = nn.Parameter(0.02 * torch.randn(1, num_frames, 1, hidden_size))
        self.token_type_embeddings = nn.Embedding(token_types, hidden_size)
        self.norm = nn.LayerNorm(hidden_size)
        self.word_pos_start, self.num_frames, self.spacial_tokens = word_pos_start, num_frames, spacial_tokens
        self.img_in_size, self.hidden_size, self.return_mask = img_in_size, hidden_size, return_mask
        if img_in_size != hidden_size:
            self.fc_in = nn.Linear(img_in_size, hidden_size)
        self.fp16_enabled = False
        std = cfg["initializer_range"]
        for m in list(self.bert_embedding.modules()) + list(self.bert_encoder.modules()):
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(0.0, std)
                m.bias.data.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.data.normal_(0.0, std)

    @property
    def extra_tokens(self):
        """Learned tokens between the video and the text tokens: 0 (use_text_cls=True), 1 (all_cls) or 5 (prompt + all_cls)."""
        return 0 if self.all_cls_token is None else (5 if self.use_prompt else 1)

    def v_seq_len(self, T, S):
        """Where the text tokens start in the fused sequence (:111-116)."""
        return T * S + self.extra_tokens

    def forward_tokens(self, v_tokens, B, T, S, text_states, text_input_mask, want_last_probs=False):
        """v_tokens: [B*T*S, img_in] (fp32 or bf16); text_states: (Bt, L, H) bf16/fp32.  Returns the
        encoder output bf16 [B, T*S + E + L', H] (E = extra_tokens; L' = L * Bt/B for multiple-choice folding, :79-82)."""
        H = self.hidden_size
        v16 = Fn.to_dtype(v_tokens.contiguous(), torch.bfloat16)
        if self.img_in_size != self.hidden_size:
            v16 = Fn.linear(v16, self.fc_in.weight, self.fc_in.bias)
        if T > self.vis_tempor_pos.shape[1]:
            raise ValueError(f"fusion encoder built with num_frames={self.vis_tempor_pos.shape[1]} but got T={T} "
                             "(cross_transformer.py:41,89)")
        t16 = Fn.to_dtype(text_states.reshape(-1, H).contiguous(), torch.bfloat16)
        L = t16.shape[0] // B
        mask = text_input_mask.reshape(B, L)
        extra, E = None, self.extra_tokens
        if E:                                                                       # :99-104
            extra = self.all_cls_token.reshape(1, H)
            if self.use_prompt:
                extra = torch.cat([self.prompt_token.reshape(4, H), extra], 0)
        z = Fn.FusionInputFn.apply(v16, t16, self.vis_space_pos, self.vis_tempor_pos, self.token_type_embeddings.weight,
                                   self.norm.weight, self.norm.bias, extra, B, T, S, L)
        tot = T * S + E + L
        full_mask = torch.cat([torch.ones(B, T * S + E, dtype=mask.dtype, device=mask.device), mask], dim=1)
        if want_last_probs:      # (out, mask, head-mean attention probabilities of the last layer (B, S', S'))
            out, probs = self.bert_encoder.forward_tokens(z, full_mask, B, tot, want_last_probs=True)
            return out.view(B, tot, H), full_mask, probs
        out = self.bert_encoder.forward_tokens(z, full_mask, B, tot)
        return out.view(B, tot, H), full_mask

    def forward(self, visual_token=None, text_input_ids=None, text_input_mask=None, text_input_embeds=None, **kwargs):
        """reference :64-124.  visual_token (B, T, S, D_in)."""
        B, T, S, Din = visual_token.shape
        if text_input_embeds is None:
            if self.word_pos_start:
                raise NotImplementedError("clover_b200: word_pos_start=True with text_input_ids is not supported")
            Bt, L = text_input_ids.shape
            text_input_embeds = self.bert_embedding(text_input_ids).view(Bt, L, self.hidden_size)
        v = visual_token.reshape(B * T * S, Din)
        out, full_mask = self.forward_tokens(v, B, T, S, text_input_embeds, text_input_mask)
        res = ModelOutput(last_hidden_state=out, attentions=None)
        vs = self.v_seq_len(T, S)
        res["t_last_hidden_state"] = out[:, vs:]
        res["v_last_hidden_state"] = out[:, :T * S]
        if self.all_cls_token is not None:
            res["cls_last_hidden_state"] = out[:, vs - 1:vs]
        if self.return_mask:
            return res, full_mask
        return res
