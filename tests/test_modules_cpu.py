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


def test_inflate_2d_checkpoint_matches_reference(golden_dir, tmp_path):
    """SwinTransformer3D.init_weights(pretrained2d=True) (swin_transformer_3d.py:130-181): 2-D patch kernel repeated over
    time / pd, 11x11 bias tables resized bicubically to 13x13 and tiled 2*wd-1 times -- against the executed reference."""
    import numpy as np
    import torch
    from clover_b200 import swin
    from clover_b200.synthetic import synth_swin2d_checkpoint
    path = str(tmp_path / "swin2d.pth")
    synth_swin2d_checkpoint(path)
    torch.manual_seed(0)
    m = swin.SwinTransformer3D(pretrained=path, pretrained2d=True, patch_size=(2, 4, 4), embed_dim=32, depths=[2, 2],
                               num_heads=[1, 2], window_size=(8, 7, 7), patch_norm=True)
    m.init_weights()
    g = np.load(os.path.join(golden_dir, "inflate_2d.npz"))
    sd = m.state_dict()
    for k in g.files:
        assert np.allclose(sd[k].numpy(), g[k], rtol=1e-6, atol=1e-7), k
    assert sd["layers.0.blocks.0.attn.relative_position_index"].shape == (392, 392)      # re-initialised, not loaded


def test_checkpoint_round_trip_reference_format(tmp_path):
    """save_checkpoint / load_checkpoint keep the reference's file layout: {'meta', 'state_dict'}, no 'module.' prefix,
    non-strict loading that reports instead of raising, stale Swin index buffers ignored."""
    import torch
    from clover_b200 import checkpoint, swin
    torch.manual_seed(1)
    kw = dict(pretrained=None, pretrained2d=False, patch_size=(2, 4, 4), embed_dim=32, depths=[2], num_heads=[1], window_size=(8, 7, 7))
    a, b = swin.SwinTransformer3D(**kw), swin.SwinTransformer3D(**kw)
    path = str(tmp_path / "epoch_1.pth")
    wrapped = torch.nn.Sequential()
    wrapped.module = a                                             # DataParallel-style wrapper
    ck = checkpoint.save_checkpoint(wrapped, path, meta=dict(epoch=1, iter=10))
    assert set(ck) == {"meta", "state_dict"} and ck["meta"]["epoch"] == 1
    assert all(not k.startswith("module.") for k in ck["state_dict"])
    disk = torch.load(path, map_location="cpu")
    disk["state_dict"] = {"module." + k: v for k, v in disk["state_dict"].items()}
    disk["state_dict"]["module.layers.0.blocks.0.attn_mask"] = torch.zeros(3)
    disk["state_dict"]["module.extra.weight"] = torch.zeros(3)
    del disk["state_dict"]["module.norm.bias"]
    torch.save(disk, path)
    res = checkpoint.load_checkpoint(b, path)
    assert res["missing_keys"] == ["norm.bias"] and res["unexpected_keys"] == ["extra.weight"]
    sa, sb = a.state_dict(), b.state_dict()
    assert all(torch.equal(sa[k], sb[k]) for k in sa if k != "norm.bias")
