"""Pins oracle/clover_oracle.py against vectors produced by executing the unmodified reference
(oracle/make_golden.py).  CPU only.  Integer tables: bit-exact.  fp32: 1e-4 relative (north-star)."""
import json
import os

import numpy as np
import pytest
import torch

from clover_b200.synthetic import make_batch, make_finetune_batch, named_tensor, synth_state_dict
from oracle import clover_oracle as O
from oracle.state_shapes import bert_shapes, finetune_shapes, pretrain_shapes, swin_shapes

torch.set_num_threads(max(1, (os.cpu_count() or 2)))


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def relerr(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_tables_bit_exact(golden_dir):
    g = _load(golden_dir, "tables.npz")
    assert np.array_equal(O.relative_position_index((8, 7, 7)), g["rel_index_877"].astype(np.int64))
    assert np.array_equal(O.relative_position_index((2, 7, 7)), g["rel_index_277"].astype(np.int64))
    meta = json.loads(str(g["meta"]))
    for i, m in enumerate(meta):
        win, sh = O.get_window_size(m["dims"], m["window_cfg"], m["shift_cfg"])
        assert list(win) == m["window"] and list(sh) == m["shift"]
        D, H, W = m["dims"]
        mask = O.compute_mask(D, H, W, win, sh)
        shape = tuple(g[f"mask_shape_{i}"])
        ref_bits = np.unpackbits(g[f"mask_bits_{i}"])[: int(np.prod(shape))].reshape(shape).astype(bool)
        assert mask.shape == shape
        assert np.array_equal(mask != 0, ref_bits)
        assert set(np.unique(mask)).issubset({0.0, -100.0})
        gi = O.window_gather_index(2, D, H, W, win, sh)
        assert np.array_equal(gi, g[f"gather_{i}"].astype(np.int64))


def _wa_state(seed=3, C=64, heads=2):
    shapes = {"relative_position_bias_table": (2535, heads), "qkv.weight": (3 * C, C), "qkv.bias": (3 * C,),
              "proj.weight": (C, C), "proj.bias": (C,)}
    return synth_state_dict(shapes, seed)


@pytest.mark.parametrize("tag,dims", [("a", (4, 14, 7)), ("b", (8, 7, 14))])
def test_window_attention(golden_dir, tag, dims):
    g = _load(golden_dir, "window_attention.npz")
    C, heads = 64, 2
    win, sh = O.get_window_size(dims, (8, 7, 7), (4, 3, 3))
    N = win[0] * win[1] * win[2]
    nW = (dims[0] // win[0]) * (dims[1] // win[1]) * (dims[2] // win[2])
    for mtag in ("nomask", "mask"):
        st = {k: v.clone().requires_grad_(True) for k, v in _wa_state().items()}
        x = (named_tensor(f"wa_x_{tag}", (nW, N, C), 5) * 20).requires_grad_(True)
        gy = named_tensor(f"wa_g_{tag}", (nW, N, C), 6) * 20
        mask = torch.from_numpy(O.compute_mask(*dims, win, sh)) if mtag == "mask" else None
        y = O.window_attention(st, "", x, heads, (8, 7, 7), mask)
        (y * gy).sum().backward()
        assert relerr(y.detach(), g[f"{tag}_{mtag}_y"]) < 1e-5
        assert relerr(x.grad, g[f"{tag}_{mtag}_dx"]) < 1e-5
        assert relerr(st["relative_position_bias_table"].grad, g[f"{tag}_{mtag}_dtable"]) < 1e-5
        assert relerr(st["qkv.weight"].grad, g[f"{tag}_{mtag}_dqkvw"]) < 1e-5


@pytest.mark.parametrize("tag,embed,depths,heads,shape", [
    ("s3", 32, [2, 2, 2], [1, 2, 4], (2, 3, 4, 112, 112)),
    ("tshift", 32, [2], [1], (1, 3, 32, 56, 56)),
    ("pad", 32, [2, 2], [1, 2], (1, 3, 6, 60, 52)),
])
def test_swin_small(golden_dir, tag, embed, depths, heads, shape):
    g = _load(golden_dir, "swin_small.npz")
    st = synth_state_dict(swin_shapes(embed, depths, heads), 11)
    x = named_tensor(f"swin_x_{tag}", shape, 12) * 20
    y = O.swin_forward(st, x, depths, heads)
    assert relerr(y, g[f"{tag}_y"]) < 1e-5
    if tag == "s3":
        st = {k: v.clone().requires_grad_(True) for k, v in st.items()}
        vm = make_batch(shape[0], seed=4)["v_token_mask"]
        ym, w = O.swin_forward(st, x, depths, heads, mask=vm)
        assert relerr(ym.detach(), g[f"{tag}_ym"]) < 1e-5
        assert np.array_equal(w.numpy(), g[f"{tag}_w"])
        gy = named_tensor("swin_g", tuple(ym.shape), 13) * 20
        (ym * gy).sum().backward()
        for k in g.files:
            if k.startswith(f"{tag}_grad::"):
                assert relerr(st[k.split("::")[1]].grad, g[k]) < 1e-4, k


def test_bert_fusion_heads(golden_dir):
    g = _load(golden_dir, "bert_fusion_heads.npz")
    batch = make_batch(3, L=16, seed=2, vocab=1000)
    ids, msk = batch["token_ids"][:, 0], batch["input_mask"][:, 0]
    st = synth_state_dict(bert_shapes("bert.", 2), 21)
    assert relerr(O.text_encoder(st, ids, msk, 2, heads=2), g["bert_last"]) < 1e-5

    fs = bert_shapes("bert_encoder.", 2, embeddings=False)
    fs.update({"vis_space_pos": (1, 1, 49, 128), "vis_tempor_pos": (1, 2, 1, 128), "token_type_embeddings.weight": (2, 128),
               "norm.weight": (128,), "norm.bias": (128,), "fc_in.weight": (128, 96), "fc_in.bias": (128,)})
    fst = synth_state_dict(fs, 22)
    vt = named_tensor("fusion_v", (3, 2, 49, 96), 23) * 20
    ts = named_tensor("fusion_t", (3, 16, 128), 24) * 20
    o = O.fusion_encoder(fst, vt, ts, msk, 2, heads=2)
    assert relerr(o["last_hidden_state"], g["fusion_last"]) < 1e-5
    assert relerr(o["t_last_hidden_state"], g["fusion_t_last"]) < 1e-5

    ms = {"predictions.transform.dense.weight": (128, 128), "predictions.transform.dense.bias": (128,),
          "predictions.transform.LayerNorm.weight": (128,), "predictions.transform.LayerNorm.bias": (128,),
          "predictions.decoder.weight": (1000, 128), "predictions.decoder.bias": (1000,),
          "predictions.bias": (1000,)}
    mst = synth_state_dict(ms, 25)
    # HF ties predictions.bias and predictions.decoder.bias; load_state_dict(strict=False) applied both in
    # key order, so the module ended up with whichever came last in its state_dict.
    cand = [relerr(O.mlm_head({**mst, "predictions.decoder.bias": mst[k]}, ts), g["mlm_logits"])
            for k in ("predictions.decoder.bias", "predictions.bias")]
    assert min(cand) < 1e-5

    h1 = synth_state_dict({"img_projector.0.weight": (96, 64), "img_projector.0.bias": (96,), "img_projector.1.weight": (96,),
                           "img_projector.1.bias": (96,), "img_projector.3.weight": (40, 96), "img_projector.3.bias": (40,),
                           "img_projector.4.weight": (40,), "img_projector.4.bias": (40,), "text_projector.0.weight": (48, 48),
                           "text_projector.0.bias": (48,), "text_projector.2.weight": (40, 48), "text_projector.2.bias": (40,)}, 26)
    feat = named_tensor("head_feat", (3, 64, 2, 7, 7), 27) * 20
    txt = named_tensor("head_txt", (3, 16, 48), 28) * 20
    assert relerr(O.nce_head_mm_vision(h1, feat), g["mm_v"]) < 1e-5
    assert relerr(O.nce_head_mm_text(h1, txt), g["mm_t"]) < 1e-5
    h2 = synth_state_dict({"img_fc1.weight": (64, 48), "img_fc1.bias": (64,), "img_bn1.weight": (64,), "img_bn1.bias": (64,),
                           "img_fc2.weight": (40, 64), "img_fc2.bias": (40,), "img_bn2.weight": (40,), "img_bn2.bias": (40,)}, 29)
    assert relerr(O.nce_head_vision(h2, txt[:, 0]), g["v_head"]) < 1e-5
    assert relerr(O.nce_head_vision(h2, txt), g["v_head_seq"]) < 1e-5
    h3 = synth_state_dict({"fc1.weight": (48, 48), "fc1.bias": (48,), "fc2.weight": (40, 48), "fc2.bias": (40,)}, 30)
    assert relerr(O.nce_head_text(h3, txt[:, 0]), g["t_head"]) < 1e-5
    q1 = synth_state_dict({"vqa_classifier.1.weight": (24, 48), "vqa_classifier.1.bias": (24,), "vqa_classifier.2.weight": (24,),
                           "vqa_classifier.2.bias": (24,), "vqa_classifier.4.weight": (30, 24), "vqa_classifier.4.bias": (30,)}, 31)
    assert relerr(O.qa_oe_head(q1, txt[:, 0]), g["qa_oe"]) < 1e-5
    q2 = synth_state_dict({"mc_vqa_classifier.1.weight": (256, 48), "mc_vqa_classifier.1.bias": (256,),
                           "mc_vqa_classifier.2.weight": (256,), "mc_vqa_classifier.2.bias": (256,),
                           "mc_vqa_classifier.4.weight": (1, 256), "mc_vqa_classifier.4.bias": (1,)}, 32)
    assert relerr(O.qa_mc_head(q2, txt[:, 0]), g["qa_mc"]) < 1e-5


def test_losses(golden_dir):
    g = _load(golden_dir, "losses.npz")
    for Bg in (6, 33):
        embs = [(named_tensor(f"loss_e{i}_{Bg}", (Bg, 24), 40) * 20).requires_grad_(True) for i in range(4)]
        d = O.exclusive_nce_ranking(*embs, t=0.05, margin=5.0)
        (d["nce_loss"] + d["rank_t_tm_loss"]).backward()
        assert abs(float(d["nce_loss"]) - float(g[f"excl_nce_{Bg}"])) < 1e-4 * abs(float(g[f"excl_nce_{Bg}"]))
        assert abs(float(d["rank_t_tm_loss"]) - float(g[f"excl_rank_{Bg}"])) < 1e-4 * max(1.0, abs(float(g[f"excl_rank_{Bg}"])))
        for i, e in enumerate(embs):
            assert relerr(e.grad, g[f"excl_grad{i}_{Bg}"]) < 1e-4
        a = (named_tensor(f"ns_a_{Bg}", (Bg, 24), 41) * 20).requires_grad_(True)
        b = (named_tensor(f"ns_b_{Bg}", (Bg, 24), 42) * 20).requires_grad_(True)
        v = O.norm_softmax_loss(a, b, 0.05, True)
        v.backward()
        assert abs(float(v) - float(g[f"normsoftmax_{Bg}"])) < 1e-4 * abs(float(g[f"normsoftmax_{Bg}"]))
        assert relerr(a.grad, g[f"normsoftmax_ga_{Bg}"]) < 1e-4
    logits = (named_tensor("focal_logits", (11, 500), 43) * 60).requires_grad_(True)
    v = O.softmax_focal_multiclass(logits, torch.from_numpy(g["focal_tgt"]), 2.0)
    v.backward()
    assert abs(float(v) - float(g["focal"])) < 1e-5 * abs(float(g["focal"]))
    assert relerr(logits.grad, g["focal_grad"]) < 1e-5
    logits2 = (named_tensor("ce_logits", (7, 30), 45) * 60).requires_grad_(True)
    v = O.cross_entropy(logits2, torch.from_numpy(g["ce_tgt"]))
    v.backward()
    assert abs(float(v) - float(g["ce"])) < 1e-5 * abs(float(g["ce"]))
    assert relerr(logits2.grad, g["ce_grad"]) < 1e-5


def _run_pretrain_case(golden_dir, fname, shapes, cfg, batch, seed, tol):
    g = _load(golden_dir, fname)
    st = synth_state_dict(shapes, seed)
    # HF ties predictions.bias <-> predictions.decoder.bias (one Parameter, two keys); the reference module
    # therefore holds the value loaded last (state_dict order: predictions.bias first, decoder.bias last).
    st = {k: v.clone().requires_grad_(True) for k, v in st.items()}
    losses, _ = O.pretrain_forward(st, batch, cfg)
    total = O.total_loss(losses)
    total.backward()
    for k, v in losses.items():
        ref = float(g[f"loss::{k}"])
        assert abs(float(v.detach()) - ref) <= tol * max(1.0, abs(ref)), (k, float(v.detach()), ref)
    assert abs(float(total.detach()) - float(g["loss::loss"])) <= tol * abs(float(g["loss::loss"]))
    for k in g.files:
        if k.startswith("gradnorm::"):
            name = k.split("::")[1]
            gn = float(st[name].grad.double().norm())
            assert abs(gn - float(g[k])) <= 10 * tol * max(1e-6, float(g[k])), (name, gn, float(g[k]))
        if k.startswith("grad::"):
            name = k.split("::")[1]
            assert relerr(st[name].grad, g[k]) < 10 * tol, name
        if k.startswith("gradsample::"):
            name = k.split("::")[1]
            idx = g["gradidx::" + name]
            assert relerr(st[name].grad.reshape(-1)[idx], g[k]) < 10 * tol, name
    nograd = set(json.loads(str(g["nograd_keys"])))
    assert "text_backbone.bert.pooler.dense.weight" in nograd
    assert st["text_backbone.bert.pooler.dense.weight"].grad is None


def test_pretrain_tiny(golden_dir):
    shapes = pretrain_shapes(32, [2, 2], [1, 2], 64, 128, 256, 1000, 64, 2, 2, 2)
    cfg = dict(depths=[2, 2], num_heads=[1, 2], text_layers=2, fusion_layers=2, bert_heads=2, vocab=1000)
    batch = make_batch(3, frames=4, L=16, seed=51, size=56, vocab=1000)
    _run_pretrain_case(golden_dir, "pretrain_tiny.npz", shapes, cfg, batch, 50, 1e-4)


@pytest.mark.slow
def test_pretrain_c1_full_size(golden_dir):
    """BASELINE config 1: Swin-T + BERT-base + 3-layer fusion, B=2, 8x224x224, L=32 (fp32 CPU)."""
    shapes = pretrain_shapes(96, [2, 2, 6, 2], [3, 6, 12, 24], 768, 768, 3072, 30522, 512, 12, 3, 4)
    cfg = dict(depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], text_layers=12, fusion_layers=3, bert_heads=12, vocab=30522)
    batch = make_batch(2, frames=8, L=32, seed=61, size=224, vocab=30522)
    _run_pretrain_case(golden_dir, "pretrain_c1.npz", shapes, cfg, batch, 60, 2e-4)


FT_SMALL = dict(depths=[2, 2], num_heads=[1, 2], text_layers=2, fusion_layers=2, bert_heads=2, vocab=1000)


@pytest.mark.slow
def test_pretrain_swin_b_headline_model(golden_dir):
    """The headline model of BASELINE c3 (Video Swin-B + BERT-base + 3-layer fusion) at B=2, 8x224x224, L=32 (fp32 CPU)."""
    shapes = pretrain_shapes(128, [2, 2, 18, 2], [4, 8, 16, 32], 1024, 768, 3072, 30522, 512, 12, 3, 4)
    cfg = dict(depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], text_layers=12, fusion_layers=3, bert_heads=12, vocab=30522)
    batch = make_batch(2, frames=8, L=32, seed=63, size=224, vocab=30522)
    _run_pretrain_case(golden_dir, "pretrain_swinb.npz", shapes, cfg, batch, 62, 2e-4)


@pytest.mark.parametrize("tag,task", [("retrieval", "retrieval"), ("qa_oe", "video_qa"), ("qa_mc", "video_qa_mc"), ("fib", "FIB")])
def test_finetune(golden_dir, tag, task):
    """CloverFinetune (BASELINE c4 / c5 shapes scaled down: 16-frame clips -> T = 8 -> the full (8,7,7) window)
    against the executed reference: train loss + gradients, then the forward_test outputs."""
    g = _load(golden_dir, f"finetune_{tag}.npz")
    shapes = finetune_shapes(task, 32, [2, 2], [1, 2], 64, 128, 256, 1000, 64, 2, 2, 8, num_labels=50)
    st = {k: v.clone().requires_grad_(True) for k, v in synth_state_dict(shapes, 70).items()}
    batch = make_finetune_batch(task, 3, frames=16, size=56, L=20, vocab=1000, seed=71, num_labels=50, choices=3)
    cfg = dict(FT_SMALL, answer_mask=True, answer_cls=False) if task == "FIB" else FT_SMALL
    losses = O.finetune_forward(st, batch, cfg, task, train=True)
    total = O.total_loss(losses)
    total.backward()
    tol = 1e-4
    for k, v in losses.items():
        ref = float(g[f"loss::{k}"])
        assert abs(float(v.detach()) - ref) <= tol * max(1.0, abs(ref)), (k, float(v.detach()), ref)
    seen = 0
    for k in g.files:
        name = k.split("::")[-1]
        if k.startswith("gradnorm::"):
            gn = float(st[name].grad.double().norm())
            assert abs(gn - float(g[k])) <= 10 * tol * max(1e-6, float(g[k])), (name, gn, float(g[k]))
            seen += 1
        elif k.startswith("grad::"):
            assert relerr(st[name].grad, g[k]) < 10 * tol, name
        elif k.startswith("gradsample::"):
            assert relerr(st[name].grad.reshape(-1)[g["gradidx::" + name]], g[k]) < 10 * tol, name
    assert seen >= 8
    with torch.no_grad():
        res = O.finetune_forward(st, batch, cfg, task, train=False)
    if task == "retrieval":
        assert relerr(res[0], g["test::visual_emb"]) < tol and relerr(res[1], g["test::text_emb"]) < tol
    else:
        assert relerr(res["result"], g["test::result"]) < tol
        rows = g["test::attention_rows"]
        assert relerr(res["attention"][:, rows], g["test::attention_sample"]) < tol


def _eval_inputs(n, noise):
    v = named_tensor(f"eval_v_{n}", (n, 96), 7).numpy() * 20
    t = v + noise * named_tensor(f"eval_t_{n}", (n, 96), 8).numpy() * 20
    t[3] = 0.0
    return v, t


def test_retrieval_metrics(golden_dir):
    """Oracle restatement of recall_for_video_text_retrieval against the executed reference function."""
    g = _load(golden_dir, "eval_retrieval.npz")
    for n, noise in ((64, 1.0), (501, 7.0)):
        m, _ = O.retrieval_metrics(*_eval_inputs(n, noise))
        for k, v in m.items():
            assert abs(v - float(g[f"{n}::{k}"])) < 1e-9, (n, k, v, float(g[f"{n}::{k}"]))
