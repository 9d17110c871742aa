"""TEST INFRASTRUCTURE ONLY -- parameter name -> shape tables of the reference's modules,
used to synthesise oracle weights without instantiating any model."""


def swin_shapes(embed, depths, heads, mask_token=True):
    sh = {"patch_embed.proj.weight": (embed, 3, 2, 4, 4), "patch_embed.proj.bias": (embed,),
          "patch_embed.norm.weight": (embed,), "patch_embed.norm.bias": (embed,)}
    if mask_token:
        sh["mask_token"] = (1, embed, 1, 1, 1)
    for s, (d, nh) in enumerate(zip(depths, heads)):
        C = embed * 2 ** s
        for j in range(d):
            p = f"layers.{s}.blocks.{j}."
            sh.update({p + "norm1.weight": (C,), p + "norm1.bias": (C,),
                       p + "attn.relative_position_bias_table": (2535, nh),
                       p + "attn.qkv.weight": (3 * C, C), p + "attn.qkv.bias": (3 * C,),
                       p + "attn.proj.weight": (C, C), p + "attn.proj.bias": (C,),
                       p + "norm2.weight": (C,), p + "norm2.bias": (C,),
                       p + "mlp.fc1.weight": (4 * C, C), p + "mlp.fc1.bias": (4 * C,),
                       p + "mlp.fc2.weight": (C, 4 * C), p + "mlp.fc2.bias": (C,)})
        if s < len(depths) - 1:
            p = f"layers.{s}.downsample."
            sh.update({p + "reduction.weight": (2 * C, 4 * C), p + "norm.weight": (4 * C,), p + "norm.bias": (4 * C,)})
    Cf = embed * 2 ** (len(depths) - 1)
    sh.update({"norm.weight": (Cf,), "norm.bias": (Cf,)})
    return sh


def bert_shapes(prefix, layers, H=128, I=256, vocab=1000, maxpos=64, embeddings=True, pooler=False):
    sh = {}
    if embeddings:
        e = prefix + "embeddings."
        sh.update({e + "word_embeddings.weight": (vocab, H), e + "position_embeddings.weight": (maxpos, H),
                   e + "token_type_embeddings.weight": (2, H), e + "LayerNorm.weight": (H,), e + "LayerNorm.bias": (H,)})
    for i in range(layers):
        p = f"{prefix}encoder.layer.{i}." if embeddings else f"{prefix}layer.{i}."
        for n in ("query", "key", "value"):
            sh[p + f"attention.self.{n}.weight"] = (H, H)
            sh[p + f"attention.self.{n}.bias"] = (H,)
        sh.update({p + "attention.output.dense.weight": (H, H), p + "attention.output.dense.bias": (H,),
                   p + "attention.output.LayerNorm.weight": (H,), p + "attention.output.LayerNorm.bias": (H,),
                   p + "intermediate.dense.weight": (I, H), p + "intermediate.dense.bias": (I,),
                   p + "output.dense.weight": (H, I), p + "output.dense.bias": (H,),
                   p + "output.LayerNorm.weight": (H,), p + "output.LayerNorm.bias": (H,)})
    return sh


def pretrain_shapes(embed, depths, heads, img_in, hidden, inter, vocab, maxpos, text_layers, fusion_layers, frames_half):
    sh = {"backbone." + k: v for k, v in swin_shapes(embed, depths, heads).items()}
    sh.update(bert_shapes("text_backbone.bert.", text_layers, hidden, inter, vocab, maxpos))
    sh.update({"text_backbone.bert.pooler.dense.weight": (hidden, hidden), "text_backbone.bert.pooler.dense.bias": (hidden,)})
    m = "multimodal_backbone."
    sh.update(bert_shapes(m + "bert_encoder.", fusion_layers, hidden, inter, embeddings=False))
    sh.update({m + "vis_space_pos": (1, 1, 49, hidden), m + "vis_tempor_pos": (1, frames_half, 1, hidden),
               m + "token_type_embeddings.weight": (2, hidden), m + "norm.weight": (hidden,), m + "norm.bias": (hidden,)})
    if img_in != hidden:
        sh.update({m + "fc_in.weight": (hidden, img_in), m + "fc_in.bias": (hidden,)})
    s = "ssl_head."
    sh.update({s + "img_projector.0.weight": (2 * hidden, img_in), s + "img_projector.0.bias": (2 * hidden,),
               s + "img_projector.1.weight": (2 * hidden,), s + "img_projector.1.bias": (2 * hidden,),
               s + "img_projector.3.weight": (hidden, 2 * hidden), s + "img_projector.3.bias": (hidden,),
               s + "img_projector.4.weight": (hidden,), s + "img_projector.4.bias": (hidden,),
               s + "text_projector.0.weight": (hidden, hidden), s + "text_projector.0.bias": (hidden,),
               s + "text_projector.2.weight": (hidden, hidden), s + "text_projector.2.bias": (hidden,)})
    p = "mlm_head.predictions."
    sh.update({p + "transform.dense.weight": (hidden, hidden), p + "transform.dense.bias": (hidden,),
               p + "transform.LayerNorm.weight": (hidden,), p + "transform.LayerNorm.bias": (hidden,),
               p + "decoder.weight": (vocab, hidden), p + "decoder.bias": (vocab,), p + "bias": (vocab,)})
    v = "mlm_ssl_V_head."
    sh.update({v + "img_fc1.weight": (2 * hidden, hidden), v + "img_fc1.bias": (2 * hidden,), v + "img_bn1.weight": (2 * hidden,),
               v + "img_bn1.bias": (2 * hidden,), v + "img_fc2.weight": (hidden, 2 * hidden), v + "img_fc2.bias": (hidden,),
               v + "img_bn2.weight": (hidden,), v + "img_bn2.bias": (hidden,)})
    t = "mlm_ssl_T_head."
    sh.update({t + "fc1.weight": (hidden, hidden), t + "fc1.bias": (hidden,), t + "fc2.weight": (hidden, hidden), t + "fc2.bias": (hidden,)})
    return sh


def finetune_shapes(task, embed, depths, heads, img_in, hidden, inter, vocab, maxpos, text_layers, fusion_layers, frames_half,
                    num_labels=1500):
    """CloverFinetune state (multimodal_transformer_finetune.py:20-57): the pre-train encoders without the mask token,
    plus the retrieval projection head or a QA head."""
    full = pretrain_shapes(embed, depths, heads, img_in, hidden, inter, vocab, maxpos, text_layers, fusion_layers, frames_half)
    keep = ("backbone.", "text_backbone.", "multimodal_backbone.") + (("ssl_head.",) if task == "retrieval" else ())
    sh = {k: v for k, v in full.items() if k.startswith(keep) and k != "backbone.mask_token"}
    if task == "FIB":                                    # use_text_cls=False encoder + ITM head (finetune_lsmdc_FIB.py)
        sh["multimodal_backbone.all_cls_token"] = (1, 1, hidden)
        i = "itm_head.itm_projector."
        sh.update({i + "1.weight": (hidden, hidden), i + "1.bias": (hidden,), i + "3.weight": (2, hidden), i + "3.bias": (2,)})
    if task in ("video_qa", "FIB"):
        q = "qa_head.vqa_classifier."
        sh.update({q + "1.weight": (hidden // 2, hidden), q + "1.bias": (hidden // 2,), q + "2.weight": (hidden // 2,),
                   q + "2.bias": (hidden // 2,), q + "4.weight": (num_labels, hidden // 2), q + "4.bias": (num_labels,)})
    elif task == "video_qa_mc":
        q = "qa_head.mc_vqa_classifier."
        sh.update({q + "1.weight": (256, hidden), q + "1.bias": (256,), q + "2.weight": (256,), q + "2.bias": (256,),
                   q + "4.weight": (1, 256), q + "4.bias": (1,)})
    return sh
