"""Model dicts of the reference's shipped configs, as Python (configs/exp_local/*.py are mmcv config files;
these helpers reproduce their ``model = dict(...)`` so tests, smoke() and bench.py build exactly that model)."""


def pretrain_cfg(embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32), img_in=1024, hidden=768, vocab=30522,
                 text_layers=12, fusion_layers=3, frames_half=4, bert_dropout=0.0, drop_path_rate=0.0, t_head_dropout=0.0, **bert):
    """The model dict of configs/exp_local/pretrain_webvid_cc3m.py:22-104 (+ swin3d_base_stride.py).  The stochastic
    regularisers default to 0 (parity runs); the shipped values are bert_dropout=0.1 (HF BertConfig defaults),
    drop_path_rate=0.3 (swin3d_base_stride.py:7) and t_head_dropout=0.1 (pretrain_webvid_cc3m.py:86): SHIPPED_REGULARISERS."""
    aux = ["token_ids", "segment_ids", "input_mask", "mlm_label", "v_token_mask"]
    # HF BERT's dropout rates; the reference classes swallow unknown kwargs (bert_from_hugface.py:9, cross_transformer.py:15)
    drop = dict(hidden_dropout_prob=bert_dropout, attention_probs_dropout_prob=bert_dropout)
    return dict(
        type="CloverPretrain", freeze_stage=None, separate_test=True, use_Cmask=True,
        backbone=dict(type="SwinTransformer3D", stride=(2, 4, 4), mask_token=True, pretrained2d=False, pretrained=None,
                      embed_dim=embed, depths=list(depths), num_heads=list(heads), patch_size=(2, 4, 4),
                      window_size=(8, 7, 7), drop_path_rate=drop_path_rate, patch_norm=True),
        freeze_text_backbone=None, text_vocab_size=vocab,
        mm_backbone=dict(type="CrossModalTransformerFromPretrained", use_text_cls=True, use_prompt=False,
                         pretrained_model="bert-base-uncased", num_hidden_layers=fusion_layers, img_in_size=img_in,
                         hidden_size=hidden, num_frames=frames_half, spacial_tokens=49, token_types=2,
                         layer_norm_eps=1e-12, word_pos_start=False, **drop, **bert),
        text_backbone=dict(type="BertFromPretrained", num_hidden_layers=text_layers,
                           **drop, **(dict(bert, hidden_size=hidden) if bert else {})),
        cls_head=None,
        ssl_head=dict(type="NCEHeadForMM", visual_in_channels=img_in, text_in_channels=hidden, img_hidden_dim=hidden * 2,
                      vts_embed_dim=hidden, ln=True, spatial_type="avg", text_agg_type="cls", dropout_ratio=0),
        mlm_head=dict(type="MLMHead", hidden_size=hidden, vocab_size=vocab),
        mlm_ssl_head=dict(V=dict(type="NCEHeadForVision", visual_in_channels=hidden, cross_in_channels=hidden,
                                 hidden_dim=hidden, ln=True, vts_embed_dim=hidden, dropout_ratio=0),
                          T=dict(type="NCEHeadForText", cross_in_channels=hidden, vts_embed_dim=hidden, text_bn=False,
                                 dropout_ratio=t_head_dropout)),
        mlm_loss=dict(type="SoftmaxFocalLossMultiClass", gamma=2.0), loss_type=dict(type="CrossEntropyLoss"),
        ssl_loss=dict(type="ExclusiveNCEwithRankingLoss", temperature=0.05, use_rank=True, use_rank_ttm=True,
                      use_rank_trtm=False, margin_ttm=5.0, margin_trtm=10.0),
        symmetry_rank=True, train_cfg=dict(aux_info=aux))


SHIPPED_REGULARISERS = dict(bert_dropout=0.1, drop_path_rate=0.3, t_head_dropout=0.1)


def finetune_cfg(task="retrieval", embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32), img_in=1024, hidden=768, vocab=30522,
                 text_layers=12, fusion_layers=3, frames_half=8, bert_dropout=0.0, num_labels=1500, qa_dropout=0.0, **bert):
    """The model dicts of configs/exp_local/finetune_msrvtt_retrieval.py:22-70 (task='retrieval'),
    finetune_msrvttQA.py:23-66 (task='video_qa', open-ended head) and finetune_msrvtt_mc.py (task='video_qa_mc',
    multiple-choice head).  frames_half is mm_backbone.num_frames (must cover T = frames/2)."""
    base = pretrain_cfg(embed, depths, heads, img_in, hidden, vocab, text_layers, fusion_layers, frames_half, bert_dropout,
                        **bert)
    cfg = dict(type="CloverFinetune", freeze_stage=None, text_vocab_size=vocab, cls_head=None, itm_head=None,
               backbone=dict(base["backbone"], mask_token=False), mm_backbone=base["mm_backbone"],
               text_backbone=base["text_backbone"], train_cfg=dict(aux_info=["token_ids", "segment_ids", "input_mask"]))
    if task == "retrieval":
        cfg.update(task="retrieval", separate_test=True, ssl_head=base["ssl_head"],
                   loss_type=dict(type="NormSoftmaxLoss", cos_sim=True, temperature=0.05),
                   test_cfg=dict(feature_extraction=False))
    elif task == "video_qa":
        cfg.update(task="video_qa", separate_test=False, ssl_head=None, answer_cls=True,
                   qa_head=dict(type="QA_OE_Head", hidden_dim=hidden, dropout_ratio=qa_dropout, num_labels=num_labels),
                   loss_type=dict(type="CrossEntropyLoss"))
    elif task == "video_qa_mc":
        cfg.update(task="video_qa", separate_test=False, ssl_head=None, answer_cls=True,
                   qa_head=dict(type="QA_MC_head", hidden_dim=hidden, dropout_ratio=qa_dropout),
                   loss_type=dict(type="CrossEntropyLoss"))
    elif task == "FIB":
        # configs/exp_local/finetune_lsmdc_FIB.py:22-60: the fusion encoder keeps the class default use_text_cls=False (an
        # all-cls token is appended to the video tokens), the answer is read at the [MASK] position, the ITM head is built
        # but not on the path
        mm = {k: v for k, v in base["mm_backbone"].items() if k not in ("use_text_cls", "use_prompt")}
        cfg.update(task="FIB", separate_test=False, ssl_head=None, answer_mask=True, mm_backbone=mm,
                   itm_head=dict(type="ITMHead", hidden_dim=hidden, dropout_ratio=0.5, finetune=True),
                   qa_head=dict(type="QA_OE_Head", hidden_dim=hidden, dropout_ratio=qa_dropout, num_labels=num_labels),
                   loss_type=dict(type="CrossEntropyLoss"))
    elif task == "video_qa_itm":
        # the answer_cls + itm_head + no qa_head branch of finetune.py:101-108,116-118 with use_text_cls=False: the all-cls
        # state goes through the ITM head and logit 1 is the score (not used by a shipped config; covered for completeness)
        mm = {k: v for k, v in base["mm_backbone"].items() if k not in ("use_text_cls", "use_prompt")}
        cfg.update(task="video_qa", separate_test=False, ssl_head=None, answer_cls=True, mm_backbone=mm, qa_head=None,
                   itm_head=dict(type="ITMHead", hidden_dim=hidden), loss_type=dict(type="CrossEntropyLoss"))
    else:
        raise ValueError(task)
    return cfg
