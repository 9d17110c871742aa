"""CPU-only: drop-in boundary -- registry selection by the reference's class names, constructor
signatures of the shipped configs, and state-dict key/shape/dtype parity with the reference."""
import json
import os

import pytest
import torch

from clover_b200 import registry
from clover_b200.configs import pretrain_cfg


def test_registry_has_reference_names():
    reg = registry.register_all()
    for name in ("SwinTransformer3D", "BertFromPretrained", "CrossModalTransformerFromPretrained", "NCEHeadForMM",
                 "NCEHeadForVision", "NCEHeadForText", "MLMHead", "QA_OE_Head", "QA_MC_head",
                 "ExclusiveNCEwithRankingLoss", "NormSoftmaxLoss", "SoftmaxFocalLossMultiClass", "CrossEntropyLoss",
                 "CloverPretrain", "CloverFinetune"):
        assert name in reg
    # same-name re-registration must need force=True, like mmcv's Registry
    with pytest.raises(KeyError):
        registry.MODELS.register_module(name="SwinTransformer3D", module=registry.MODELS.get("SwinTransformer3D"))


def test_state_dict_matches_reference_swinb(golden_dir):
    registry.register_all()
    ref = json.load(open(os.path.join(golden_dir, "state_keys_pretrain_swinb.json")))
    model = registry.build_model(pretrain_cfg())
    sd = model.state_dict()
    ours = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()}
    assert set(ours) == set(ref), (sorted(set(ours) ^ set(ref))[:20])
    for k in ref:
        assert ours[k] == ref[k], (k, ours[k], ref[k])
    assert sum(p.numel() for p in model.parameters()) == 274604674       # 274.6 M
    # reference-format checkpoints round-trip with strict=True; 4.6.1's persistent position_ids is tolerated
    extra = dict(sd)
    extra["text_backbone.bert.embeddings.position_ids"] = torch.arange(512)[None]
    model.load_state_dict(extra, strict=True)


def test_finetune_configs_build():
    registry.register_all()
    base = pretrain_cfg(embed=32, depths=(2, 2), heads=(1, 2), img_in=64, hidden=128, vocab=1000, text_layers=1,
                        fusion_layers=1, frames_half=2, num_attention_heads=2, intermediate_size=256,
                        max_position_embeddings=64, vocab_size=1000)
    common = {k: base[k] for k in ("backbone", "mm_backbone", "text_backbone")}
    r = registry.build_model(dict(type="CloverFinetune", task="retrieval", separate_test=True, ssl_head=base["ssl_head"],
                                  loss_type=dict(type="NormSoftmaxLoss", temperature=0.05, cos_sim=True), **common))
    assert hasattr(r, "ssl_head") and r.loss_func.use_cos_similarity
    q = registry.build_model(dict(type="CloverFinetune", task="video_qa", answer_cls=True,
                                  qa_head=dict(type="QA_OE_Head", hidden_dim=128, dropout_ratio=0.0, num_labels=1500),
                                  loss_type=dict(type="CrossEntropyLoss"), **common))
    assert q.qa_head.num_labels == 1500
    with pytest.raises(NotImplementedError):
        registry.build_model(dict(type="CloverFinetune", task="nope", **common))
