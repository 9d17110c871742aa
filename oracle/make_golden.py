"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz|json by EXECUTING THE UNMODIFIED
REFERENCE (through oracle/ref_shim.py) in the build container.

    python -m oracle.make_golden            # needs /root/reference; writes tests/golden/

The fixtures hold only integer tables and reference OUTPUTS (plus state-dict key lists); weights
and inputs are regenerated bit-identically from clover_b200.synthetic (numpy PCG64 keyed by
tensor name), which keeps the committed files small.  Dropout / drop-path are zeroed
(SURVEY.md section 7 hard part 7).
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402
from clover_b200.synthetic import named_tensor, synth_state_dict, make_batch, make_finetune_batch, synth_swin2d_checkpoint  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL_BERT = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, vocab_size=1000,
                  max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


def load_synth(module, seed=0):
    sd = module.state_dict()
    new = synth_state_dict(sd, seed)
    module.load_state_dict(new, strict=False)
    return new


def zero_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0


def gen_tables(ref):
    sw = ref.swin
    out = {}
    att = sw.WindowAttention3D(32, (8, 7, 7), 1)
    out["rel_index_877"] = att.relative_position_index.numpy().astype(np.int16)
    att2 = sw.WindowAttention3D(32, (2, 7, 7), 1)
    out["rel_index_277"] = att2.relative_position_index.numpy().astype(np.int16)
    cases = [  # (D,H,W) configured window, configured shift
        ((4, 56, 56), (8, 7, 7), (4, 3, 3)),
        ((8, 28, 28), (8, 7, 7), (4, 3, 3)),
        ((16, 14, 14), (8, 7, 7), (4, 3, 3)),
        ((4, 7, 7), (8, 7, 7), (4, 3, 3)),
        ((2, 14, 14), (8, 7, 7), (4, 3, 3)),
        ((16, 7, 7), (8, 7, 7), (4, 3, 3)),
    ]
    meta = []
    for i, (dims, wcfg, scfg) in enumerate(cases):
        win, sh = sw.get_window_size(dims, wcfg, scfg)
        D, H, W = dims
        m = sw.compute_mask(D, H, W, win, sh, torch.device("cpu"), torch.float32).numpy()
        assert set(np.unique(m)).issubset({0.0, -100.0})
        out[f"mask_bits_{i}"] = np.packbits(m != 0)
        out[f"mask_shape_{i}"] = np.array(m.shape)
        B = 2
        x = torch.arange(B * D * H * W, dtype=torch.int64).view(B, D, H, W, 1)
        if any(s > 0 for s in sh):
            xs = torch.roll(x, shifts=(-sh[0], -sh[1], -sh[2]), dims=(1, 2, 3))
        else:
            xs = x
        part = sw.window_partition(xs, win)[..., 0]
        out[f"gather_{i}"] = part.numpy().astype(np.int32)
        # reverse + roll back must be the inverse permutation
        back = sw.window_reverse(part.view(-1, *win, 1), win, B, D, H, W)
        if any(s > 0 for s in sh):
            back = torch.roll(back, shifts=sh, dims=(1, 2, 3))
        assert torch.equal(back, x)
        meta.append(dict(dims=dims, window_cfg=wcfg, shift_cfg=scfg, window=win, shift=sh))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **out)


def gen_window_attention(ref):
    sw = ref.swin
    out = {}
    for tag, dims, heads in (("a", (4, 14, 7), 2), ("b", (8, 7, 14), 2)):
        C = 64
        m = sw.WindowAttention3D(C, (8, 7, 7), heads, qkv_bias=True)
        load_synth(m, seed=3)
        win, sh = sw.get_window_size(dims, (8, 7, 7), (4, 3, 3))
        N = win[0] * win[1] * win[2]
        nW = (dims[0] // win[0]) * (dims[1] // win[1]) * (dims[2] // win[2])
        B_ = nW
        x = named_tensor(f"wa_x_{tag}", (B_, N, C), 5) * 20
        x.requires_grad_(True)
        g = named_tensor(f"wa_g_{tag}", (B_, N, C), 6) * 20
        mask = sw.compute_mask(*dims, win, sh, torch.device("cpu"), torch.float32)
        for mtag, mk in (("nomask", None), ("mask", mask)):
            m.zero_grad()
            x.grad = None
            y = m(x, mk)
            (y * g).sum().backward()
            out[f"{tag}_{mtag}_y"] = y.detach().numpy()
            out[f"{tag}_{mtag}_dx"] = x.grad.numpy().copy()
            out[f"{tag}_{mtag}_dtable"] = m.relative_position_bias_table.grad.numpy().copy()
            out[f"{tag}_{mtag}_dqkvw"] = m.qkv.weight.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "window_attention.npz"), **out)


def gen_swin(ref):
    sw = ref.swin
    out = {}
    # (tag, embed, depths, heads, input shape)
    cases = [("s3", 32, [2, 2, 2], [1, 2, 4], (2, 3, 4, 112, 112)),
             ("tshift", 32, [2], [1], (1, 3, 32, 56, 56)),
             ("pad", 32, [2, 2], [1, 2], (1, 3, 6, 60, 52))]
    for tag, embed, depths, heads, shape in cases:
        torch.manual_seed(0)
        m = sw.SwinTransformer3D(pretrained=None, pretrained2d=False, embed_dim=embed, depths=depths,
                                 num_heads=heads, window_size=(8, 7, 7), drop_path_rate=0.0, mask_token=True)
        load_synth(m, seed=11)
        m.eval()
        x = named_tensor(f"swin_x_{tag}", shape, 12) * 20
        y = m(x)
        out[f"{tag}_y"] = y.detach().numpy()
        if tag == "s3":
            vm = make_batch(shape[0], seed=4)["v_token_mask"]
            ym, w = m(x, vm)
            out[f"{tag}_ym"] = ym.detach().numpy()
            out[f"{tag}_w"] = w.detach().numpy()
            g = named_tensor("swin_g", tuple(ym.shape), 13) * 20
            (ym * g).sum().backward()
            for k in ("patch_embed.proj.weight", "mask_token", "layers.0.blocks.1.attn.relative_position_bias_table",
                      "layers.1.downsample.reduction.weight", "layers.2.blocks.0.mlp.fc1.bias", "norm.weight"):
                out[f"{tag}_grad::{k}"] = dict(m.named_parameters())[k].grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "swin_small.npz"), **out)


def gen_bert_fusion_heads(ref):
    out = {}
    ref_shim.BERT_OVERRIDES.clear()
    ref_shim.BERT_OVERRIDES.update(SMALL_BERT)
    torch.manual_seed(0)
    tb = ref.bert.BertFromPretrained(num_hidden_layers=2)
    load_synth(tb, seed=21)
    tb.eval()
    batch = make_batch(3, L=16, seed=2, vocab=1000)
    ids = batch["token_ids"][:, 0]
    msk = batch["input_mask"][:, 0]
    out["bert_last"] = tb(ids, msk)["last_hidden_state"].detach().numpy()

    fm = ref.cross.CrossModalTransformerFromPretrained(
        img_in_size=96, hidden_size=128, num_frames=2, spacial_tokens=49, token_types=2, num_hidden_layers=2,
        layer_norm_eps=1e-12, use_text_cls=True)
    load_synth(fm, seed=22)
    fm.eval()
    vt = named_tensor("fusion_v", (3, 2, 49, 96), 23) * 20
    ts = named_tensor("fusion_t", (3, 16, 128), 24) * 20
    o = fm(visual_token=vt, text_input_mask=msk, text_input_embeds=ts)
    out["fusion_last"] = o["last_hidden_state"].detach().numpy()
    out["fusion_t_last"] = o["t_last_hidden_state"].detach().numpy()

    mh = ref.mlm_head.MLMHead(128, 1000)
    load_synth(mh, seed=25)
    out["mlm_logits"] = mh(ts).detach().numpy()
    ref_shim.BERT_OVERRIDES.clear()

    H = ref.ssl_head
    h1 = H.NCEHeadForMM(visual_in_channels=64, text_in_channels=48, img_hidden_dim=96, vts_embed_dim=40, ln=True,
                        spatial_type="avg", text_agg_type="cls", dropout_ratio=0)
    load_synth(h1, seed=26)
    feat = named_tensor("head_feat", (3, 64, 2, 7, 7), 27) * 20
    txt = named_tensor("head_txt", (3, 16, 48), 28) * 20
    v, t = h1(feat, txt)
    out["mm_v"], out["mm_t"] = v.detach().numpy(), t.detach().numpy()
    h2 = H.NCEHeadForVision(visual_in_channels=48, cross_in_channels=48, hidden_dim=32, ln=True, vts_embed_dim=40,
                            dropout_ratio=0)
    load_synth(h2, seed=29)
    out["v_head"] = h2(txt[:, 0]).detach().numpy()          # (B,C) input through the D1 workaround
    out["v_head_seq"] = h2(txt).detach().numpy()            # (B,S,C) input: mean over S
    h3 = H.NCEHeadForText(cross_in_channels=48, vts_embed_dim=40, text_bn=False, dropout_ratio=0.0)
    load_synth(h3, seed=30)
    out["t_head"] = h3(txt[:, 0]).detach().numpy()
    q1 = ref.qa_head.QA_OE_Head(hidden_dim=48, dropout_ratio=0.0, num_labels=30)
    load_synth(q1, seed=31)
    out["qa_oe"] = q1(txt[:, 0]).detach().numpy()
    q2 = ref.qa_head.QA_MC_head(48, dropout_ratio=0.0)
    load_synth(q2, seed=32)
    out["qa_mc"] = q2(txt[:, 0]).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "bert_fusion_heads.npz"), **out)


def gen_losses(ref):
    out = {}
    ref_shim.ensure_gloo_group()
    L = ref.contrastive
    for Bg in (6, 33):
        embs = [(named_tensor(f"loss_e{i}_{Bg}", (Bg, 24), 40) * 20).requires_grad_(True) for i in range(4)]
        lf = L.ExclusiveNCEwithRankingLoss(temperature=0.05, use_rank=True, use_rank_ttm=True, use_rank_trtm=False,
                                           margin_ttm=5.0)
        d = lf(*embs)
        (d["nce_loss"] + d["rank_t_tm_loss"]).backward()
        out[f"excl_nce_{Bg}"] = d["nce_loss"].detach().numpy()
        out[f"excl_rank_{Bg}"] = d["rank_t_tm_loss"].detach().numpy()
        for i, e in enumerate(embs):
            out[f"excl_grad{i}_{Bg}"] = e.grad.numpy().copy()
        a = (named_tensor(f"ns_a_{Bg}", (Bg, 24), 41) * 20).requires_grad_(True)
        b = (named_tensor(f"ns_b_{Bg}", (Bg, 24), 42) * 20).requires_grad_(True)
        ns = L.NormSoftmaxLoss(temperature=0.05, cos_sim=True)
        v = ns(a, b)
        v.backward()
        out[f"normsoftmax_{Bg}"] = v.detach().numpy()
        out[f"normsoftmax_ga_{Bg}"] = a.grad.numpy().copy()
    logits = (named_tensor("focal_logits", (11, 500), 43) * 60).requires_grad_(True)
    tgt = torch.from_numpy(np.random.default_rng(44).integers(0, 500, size=11))
    fl = ref.focal_loss.SoftmaxFocalLossMultiClass(gamma=2.0)
    v = fl(logits, tgt)
    v.backward()
    out["focal"] = v.detach().numpy()
    out["focal_grad"] = logits.grad.numpy().copy()
    out["focal_tgt"] = tgt.numpy()
    logits2 = (named_tensor("ce_logits", (7, 30), 45) * 60).requires_grad_(True)
    tgt2 = torch.from_numpy(np.random.default_rng(46).integers(0, 30, size=7))
    ce = ref.ce_loss.CrossEntropyLoss()
    v = ce(logits2, tgt2)
    v.backward()
    out["ce"] = v.detach().numpy()
    out["ce_grad"] = logits2.grad.numpy().copy()
    out["ce_tgt"] = tgt2.numpy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)


def pretrain_cfg(embed, depths, heads, img_in, hidden, vocab, text_layers, fusion_layers, frames_half):
    aux = ["token_ids", "segment_ids", "input_mask", "mlm_label", "v_token_mask"]
    return dict(
        type="CloverPretrain", separate_test=True, use_Cmask=True,
        backbone=dict(type="SwinTransformer3D", stride=(2, 4, 4), mask_token=True, pretrained2d=False, pretrained=None,
                      embed_dim=embed, depths=depths, num_heads=heads, patch_size=(2, 4, 4), window_size=(8, 7, 7),
                      drop_path_rate=0.0, patch_norm=True),
        text_vocab_size=vocab,
        mm_backbone=dict(type="CrossModalTransformerFromPretrained", use_text_cls=True, use_prompt=False,
                         num_hidden_layers=fusion_layers, img_in_size=img_in, hidden_size=hidden,
                         num_frames=frames_half, spacial_tokens=49, token_types=2, layer_norm_eps=1e-12,
                         word_pos_start=False),
        text_backbone=dict(type="BertFromPretrained", num_hidden_layers=text_layers),
        cls_head=None,
        ssl_head=dict(type="NCEHeadForMM", visual_in_channels=img_in, text_in_channels=hidden,
                      img_hidden_dim=hidden * 2, vts_embed_dim=hidden, ln=True, spatial_type="avg",
                      text_agg_type="cls", dropout_ratio=0),
        mlm_head=dict(type="MLMHead", hidden_size=hidden, vocab_size=vocab),
        mlm_ssl_head=dict(
            V=dict(type="NCEHeadForVision", visual_in_channels=hidden, cross_in_channels=hidden, hidden_dim=hidden,
                   ln=True, vts_embed_dim=hidden, dropout_ratio=0),
            T=dict(type="NCEHeadForText", cross_in_channels=hidden, vts_embed_dim=hidden, text_bn=False,
                   dropout_ratio=0.0)),
        mlm_loss=dict(type="SoftmaxFocalLossMultiClass", gamma=2.0),
        loss_type=dict(type="CrossEntropyLoss"),
        ssl_loss=dict(type="ExclusiveNCEwithRankingLoss", temperature=0.05, use_rank=True, use_rank_ttm=True,
                      use_rank_trtm=False, margin_ttm=5.0, margin_trtm=10.0),
        symmetry_rank=True, train_cfg=dict(aux_info=aux))


def _sample_idx(flat):
    """2048 sample positions of a large gradient; for sparse gradients (word embeddings: only the rows of the tokens in
    the batch are non-zero) the sample is drawn from the non-zero entries, so it is not a handful of lucky hits."""
    rng = np.random.default_rng(5)
    nz = np.flatnonzero(flat.numpy())
    if nz.size < flat.numel() // 2:
        return np.sort(rng.choice(nz, size=min(2048, nz.size), replace=False))
    return rng.integers(0, flat.numel(), size=2048)


GRAD_KEYS = ["backbone.patch_embed.proj.weight", "backbone.mask_token",
             "backbone.layers.0.blocks.1.attn.relative_position_bias_table",
             "backbone.layers.1.blocks.0.attn.qkv.weight", "backbone.norm.bias",
             "text_backbone.bert.embeddings.word_embeddings.weight",
             "text_backbone.bert.encoder.layer.1.attention.self.query.weight",
             "multimodal_backbone.fc_in.weight", "multimodal_backbone.vis_tempor_pos",
             "multimodal_backbone.bert_encoder.layer.0.output.dense.weight",
             "ssl_head.img_projector.3.weight", "ssl_head.text_projector.0.bias",
             "mlm_head.predictions.decoder.weight", "mlm_ssl_V_head.img_fc1.weight", "mlm_ssl_T_head.fc2.weight"]


SWINB_EXTRA_GRAD_KEYS = ["backbone.layers.0.blocks.0.norm1.weight", "backbone.layers.0.blocks.0.attn.qkv.weight",
                         "backbone.layers.2.blocks.9.mlp.fc1.weight", "backbone.layers.2.blocks.17.attn.proj.weight",
                         "backbone.layers.2.blocks.4.attn.relative_position_bias_table",
                         "backbone.layers.3.blocks.1.attn.relative_position_bias_table", "backbone.layers.3.blocks.1.mlp.fc2.bias",
                         "backbone.layers.1.downsample.reduction.weight",
                         "text_backbone.bert.encoder.layer.11.output.dense.weight",
                         "multimodal_backbone.bert_encoder.layer.2.attention.self.value.weight"]


def gen_pretrain(ref, tag, cfg, bert_over, B, frames, size, L, vocab, seed, full_grads, extra_keys=(), embeddings=False):
    ref_shim.ensure_gloo_group()
    ref_shim.BERT_OVERRIDES.clear()
    ref_shim.BERT_OVERRIDES.update(bert_over)
    torch.manual_seed(0)
    m = ref.builder.build_model(cfg)
    ref_shim.BERT_OVERRIDES.clear()
    zero_dropout(m)
    load_synth(m, seed=seed)
    batch = make_batch(B, frames=frames, L=L, seed=seed + 1, size=size, vocab=vocab)
    kw = {k: batch[k] for k in ("token_ids", "segment_ids", "input_mask", "mlm_label", "v_token_mask")}
    captured = []
    hook = m.ssl_loss.register_forward_pre_hook(lambda mod, args: captured.append([a.detach().clone() for a in args]))
    losses = m(batch["imgs"], batch["label"], return_loss=True, **kw)
    hook.remove()
    total, log_vars = m._parse_losses(losses)
    total.backward()
    out = {f"loss::{k}": np.float64(v) for k, v in log_vars.items()}
    if embeddings:
        # the two ssl_loss calls of forward_train (pretrain.py:151,161): (V_e, T_e, T_m, M_Vmf) then (T_e, V_e, V_m, M_Tmf)
        (v_e, t_e, t_m, m_vmf), (_, _, v_m, m_tmf) = captured
        for n, e in (("v", v_e), ("t", t_e), ("tm", t_m), ("vmf", m_vmf), ("vm", v_m), ("tmf", m_tmf)):
            out[f"emb::{n}"] = e.float().numpy().copy()
    params = dict(m.named_parameters())
    out["nograd_keys"] = np.array(json.dumps(sorted(k for k, p in params.items() if p.grad is None)))
    for k in list(GRAD_KEYS) + list(extra_keys):
        if k not in params:
            continue
        g = params[k].grad
        out[f"gradnorm::{k}"] = np.float64(g.double().norm())
        if full_grads and g.numel() <= 70000:
            out[f"grad::{k}"] = g.numpy().copy()
        else:
            flat = g.reshape(-1)
            idx = _sample_idx(flat)
            out[f"gradidx::{k}"] = idx
            out[f"gradsample::{k}"] = flat.numpy()[idx].copy()
    np.savez_compressed(os.path.join(OUT, f"pretrain_{tag}.npz"), **out)
    return m


FT_GRAD_KEYS = ["backbone.patch_embed.proj.weight", "backbone.layers.0.blocks.1.attn.relative_position_bias_table",
                "backbone.layers.1.blocks.0.attn.qkv.weight", "backbone.norm.bias",
                "text_backbone.bert.embeddings.word_embeddings.weight",
                "text_backbone.bert.encoder.layer.1.attention.self.query.weight",
                "multimodal_backbone.fc_in.weight", "multimodal_backbone.vis_tempor_pos",
                "multimodal_backbone.bert_encoder.layer.0.output.dense.weight",
                "ssl_head.img_projector.3.weight", "ssl_head.text_projector.0.bias",
                "qa_head.vqa_classifier.1.weight", "qa_head.vqa_classifier.4.bias",
                "qa_head.mc_vqa_classifier.1.weight", "qa_head.mc_vqa_classifier.4.weight",
                "multimodal_backbone.all_cls_token"]


def gen_finetune(ref, tag, task, cfg, bert_over, B, frames, size, L, vocab, seed, num_labels=50):
    """Executes the reference CloverFinetune (multimodal_transformer_finetune.py:59-197): forward_train losses +
    gradients, then forward_test outputs in eval mode."""
    ref_shim.ensure_gloo_group()
    ref_shim.BERT_OVERRIDES.clear()
    ref_shim.BERT_OVERRIDES.update(bert_over)
    torch.manual_seed(0)
    m = ref.builder.build_model(cfg)
    ref_shim.BERT_OVERRIDES.clear()
    zero_dropout(m)
    load_synth(m, seed=seed)
    batch = make_finetune_batch(task, B, frames, size, L, vocab, seed + 1, num_labels, choices=3)
    kw = {k: batch[k] for k in ("token_ids", "segment_ids", "input_mask")}
    losses = m(batch["imgs"], batch["label"], return_loss=True, **kw)
    total, log_vars = m._parse_losses(losses)
    total.backward()
    out = {f"loss::{k}": np.float64(v) for k, v in log_vars.items()}
    params = dict(m.named_parameters())
    out["nograd_keys"] = np.array(json.dumps(sorted(k for k, p in params.items() if p.grad is None)))
    for k in FT_GRAD_KEYS:
        if k not in params or params[k].grad is None:
            continue
        g = params[k].grad
        out[f"gradnorm::{k}"] = np.float64(g.double().norm())
        if g.numel() <= 70000:
            out[f"grad::{k}"] = g.numpy().copy()
        else:
            flat = g.reshape(-1)
            idx = _sample_idx(flat)
            out[f"gradidx::{k}"] = idx
            out[f"gradsample::{k}"] = flat.numpy()[idx].copy()
    m.eval()
    with torch.no_grad():
        res = m(batch["imgs"], None, return_loss=False, **kw)
    if task == "retrieval":
        out["test::visual_emb"], out["test::text_emb"] = res[0].numpy(), res[1].numpy()
    else:
        out["test::result"] = res["result"].numpy()
        att = res["attention"].numpy().astype(np.float32)               # (B, S, S): head-mean of the last fusion layer
        rows = np.array([0, att.shape[1] // 2, att.shape[1] - L, att.shape[1] - 1])
        out["test::attention_rows"], out["test::attention_sample"] = rows, att[:, rows]
    np.savez_compressed(os.path.join(OUT, f"finetune_{tag}.npz"), **out)
    print(tag, {k: float(v) for k, v in log_vars.items()})
    return m


def gen_eval(ref):
    """Executes recall_for_video_text_retrieval (core/evaluation/accuracy.py:427-456) on seeded embeddings with
    structure (text i is a noisy copy of video i) so that the ranks are spread over 0..N-1."""
    import importlib.util
    sys.modules["mmaction.utils"].normalize_fn = ref_shim._load("mmaction.utils.numpy_norm", "mmaction/utils/numpy_norm.py").normalize_fn
    acc = ref_shim._load("mmaction.core.evaluation.accuracy", "mmaction/core/evaluation/accuracy.py")
    out = {}
    for n, noise in ((64, 1.0), (501, 7.0)):
        v = named_tensor(f"eval_v_{n}", (n, 96), 7).numpy() * 20
        t = v + noise * named_tensor(f"eval_t_{n}", (n, 96), 8).numpy() * 20
        t[3] = 0.0                                              # a zero row (normalize_fn leaves it)
        m = acc.recall_for_video_text_retrieval(video_embd=v, text_embd=t)
        for k, val in m.items():
            out[f"{n}::{k}"] = np.float64(val)
    np.savez_compressed(os.path.join(OUT, "eval_retrieval.npz"), **out)
    print({k: float(v) for k, v in out.items()})


def gen_inflate(ref):
    """Executes SwinTransformer3D.inflate_weights (swin_transformer_3d.py:130-181) on the synthetic 2-D checkpoint."""
    import logging
    import tempfile
    path = os.path.join(tempfile.mkdtemp(), "swin2d.pth")
    synth_swin2d_checkpoint(path)
    torch.manual_seed(0)
    m = ref.swin.SwinTransformer3D(pretrained=path, pretrained2d=True, patch_size=(2, 4, 4), embed_dim=32, depths=[2, 2],
                                   num_heads=[1, 2], window_size=(8, 7, 7), patch_norm=True)
    m.inflate_weights(logging.getLogger("golden"))
    sd = m.state_dict()
    out = {"patch_embed.proj.weight": sd["patch_embed.proj.weight"].numpy(),
           "layers.0.blocks.1.attn.relative_position_bias_table": sd["layers.0.blocks.1.attn.relative_position_bias_table"].numpy(),
           "layers.1.blocks.0.attn.relative_position_bias_table": sd["layers.1.blocks.0.attn.relative_position_bias_table"].numpy(),
           "layers.1.blocks.0.mlp.fc1.weight": sd["layers.1.blocks.0.mlp.fc1.weight"].numpy()}
    np.savez_compressed(os.path.join(OUT, "inflate_2d.npz"), **out)
    print({k: v.shape for k, v in out.items()})


def gen_state_keys(ref):
    cfg = pretrain_cfg(128, [2, 2, 18, 2], [4, 8, 16, 32], 1024, 768, 30522, 12, 3, 4)
    torch.manual_seed(0)
    m = ref.builder.build_model(cfg)
    keys = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()}
    json.dump(keys, open(os.path.join(OUT, "state_keys_pretrain_swinb.json"), "w"))
    print("swin-b pretrain params (M):", sum(p.numel() for p in m.parameters()) / 1e6)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = ref_shim.load_reference()
    which = sys.argv[1:] or ["tables", "wa", "swin", "bfh", "losses", "tiny", "c1", "swinb", "ft", "eval", "inflate", "keys"]
    if "tables" in which:
        gen_tables(ref)
    if "wa" in which:
        gen_window_attention(ref)
    if "swin" in which:
        gen_swin(ref)
    if "bfh" in which:
        gen_bert_fusion_heads(ref)
    if "losses" in which:
        gen_losses(ref)
    if "tiny" in which:
        cfg = pretrain_cfg(32, [2, 2], [1, 2], 64, 128, 1000, 2, 2, 2)
        gen_pretrain(ref, "tiny", cfg, SMALL_BERT, B=3, frames=4, size=56, L=16, vocab=1000, seed=50, full_grads=True,
                     embeddings=True)
    if "c1" in which:
        cfg = pretrain_cfg(96, [2, 2, 6, 2], [3, 6, 12, 24], 768, 768, 30522, 12, 3, 4)
        gen_pretrain(ref, "c1", cfg, dict(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
                     B=2, frames=8, size=224, L=32, vocab=30522, seed=60, full_grads=False, embeddings=True)
    if "swinb" in which:
        # the headline configuration (BASELINE c3 model: Video Swin-B + BERT-base + 3-layer fusion) at B = 2
        cfg = pretrain_cfg(128, [2, 2, 18, 2], [4, 8, 16, 32], 1024, 768, 30522, 12, 3, 4)
        gen_pretrain(ref, "swinb", cfg, dict(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
                     B=2, frames=8, size=224, L=32, vocab=30522, seed=62, full_grads=False,
                     extra_keys=SWINB_EXTRA_GRAD_KEYS, embeddings=True)
    only = [w[3:] for w in which if w.startswith("ft:")]          # e.g. `ft:fib` regenerates one fine-tune fixture
    if "ft" in which or only:
        from clover_b200.configs import finetune_cfg
        small = dict(embed=32, depths=[2, 2], heads=[1, 2], img_in=64, hidden=128, vocab=1000, text_layers=2, fusion_layers=2,
                     frames_half=8)
        # 16-frame clips -> T = 8 token frames -> the full (8,7,7) window, N = 392 (BASELINE c4 / c5 shapes)
        for tag, task in (("retrieval", "retrieval"), ("qa_oe", "video_qa"), ("qa_mc", "video_qa_mc"), ("fib", "FIB")):
            if only and tag not in only:
                continue
            cfg = finetune_cfg(task, num_labels=50, **small)
            for k in ("hidden_dropout_prob", "attention_probs_dropout_prob"):
                cfg["mm_backbone"].pop(k, None), cfg["text_backbone"].pop(k, None)
            cfg["mm_backbone"].pop("pretrained_model", None)
            gen_finetune(ref, tag, task, cfg, SMALL_BERT, B=3, frames=16, size=56, L=20, vocab=1000, seed=70)
    if "eval" in which:
        gen_eval(ref)
    if "inflate" in which:
        gen_inflate(ref)
    if "keys" in which:
        gen_state_keys(ref)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
