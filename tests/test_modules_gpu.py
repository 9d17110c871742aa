"""Module-level parity on the B200: the drop-in nn.Modules (clover_b200.*) against (a) the golden
vectors produced by executing the unmodified reference and (b) the CPU oracle on the same seeded
weights and inputs.  bf16 path vs fp32 reference: <= 2e-2 relative on outputs / gradients,
embedding cosine >= 0.999, losses within 1e-3 (north-star tolerances; relative to max(1, |loss|))."""
import os

import numpy as np
import pytest
import torch

from clover_b200.synthetic import make_batch, named_tensor, synth_state_dict
from oracle import clover_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import clover_b200.registry as reg
    reg.register_all()
    return reg


def rel(a, b):
    a = torch.as_tensor(np.asarray(a.detach().float().cpu() if torch.is_tensor(a) else a)).double()
    b = torch.as_tensor(np.asarray(b.detach().float().cpu() if torch.is_tensor(b) else b)).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.detach().double().cpu().reshape(-1), torch.as_tensor(np.asarray(b)).double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()))


def load_synth(module, seed):
    sd = synth_state_dict(module.state_dict(), seed)
    module.load_state_dict(sd, strict=False)
    return {k: v.clone() for k, v in sd.items()}


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("tag,dims", [("a", (4, 14, 7)), ("b", (8, 7, 14))])
def test_window_attention_module_vs_reference_golden(cb, golden_dir, tag, dims):
    from clover_b200 import swin
    g = _g(golden_dir, "window_attention.npz")
    C, heads = 64, 2
    m = swin.WindowAttention3D(C, (8, 7, 7), heads, qkv_bias=True).cuda()
    load_synth(m, 3)
    win, sh = swin.get_window_size(dims, (8, 7, 7), (4, 3, 3))
    N = win[0] * win[1] * win[2]
    nW = (dims[0] // win[0]) * (dims[1] // win[1]) * (dims[2] // win[2])
    for mtag in ("nomask", "mask"):
        m.zero_grad()
        x = (named_tensor(f"wa_x_{tag}", (nW, N, C), 5) * 20).cuda().requires_grad_(True)
        gy = (named_tensor(f"wa_g_{tag}", (nW, N, C), 6) * 20).cuda()
        mask = swin.compute_mask(*dims, win, sh, "cuda") if mtag == "mask" else None
        if mask is not None:   # the materialised mask equals the reference's bit for bit
            assert np.array_equal(mask.cpu().numpy(), O.compute_mask(*dims, win, sh))
        y = m(x, mask)
        (y * gy).sum().backward()
        assert rel(y, g[f"{tag}_{mtag}_y"]) < TOL
        assert rel(x.grad, g[f"{tag}_{mtag}_dx"]) < TOL
        assert rel(m.relative_position_bias_table.grad, g[f"{tag}_{mtag}_dtable"]) < TOL
        assert rel(m.qkv.weight.grad, g[f"{tag}_{mtag}_dqkvw"]) < TOL


@pytest.mark.parametrize("tag,embed,depths,heads,shape", [
    ("s3", 32, [2, 2, 2], [1, 2, 4], (2, 3, 4, 112, 112)),
    ("tshift", 32, [2], [1], (1, 3, 32, 56, 56)),
    ("pad", 32, [2, 2], [1, 2], (1, 3, 6, 60, 52)),
])
def test_swin_backbone_vs_reference_golden(cb, golden_dir, tag, embed, depths, heads, shape):
    from clover_b200 import swin
    g = _g(golden_dir, "swin_small.npz")
    m = swin.SwinTransformer3D(pretrained=None, pretrained2d=False, embed_dim=embed, depths=depths, num_heads=heads,
                               window_size=(8, 7, 7), drop_path_rate=0.0, mask_token=True).cuda()
    load_synth(m, 11)
    x = (named_tensor(f"swin_x_{tag}", shape, 12) * 20).cuda()
    y = m(x)
    assert tuple(y.shape) == g[f"{tag}_y"].shape
    assert rel(y, g[f"{tag}_y"]) < TOL
    if tag == "s3":
        vm = make_batch(shape[0], seed=4)["v_token_mask"].cuda()
        ym, w = m(x, vm)
        assert rel(ym, g[f"{tag}_ym"]) < TOL
        assert np.array_equal(w.cpu().numpy(), g[f"{tag}_w"])
        gy = (named_tensor("swin_g", tuple(ym.shape), 13) * 20).cuda()
        (ym * gy).sum().backward()
        params = dict(m.named_parameters())
        for k in g.files:
            if k.startswith(f"{tag}_grad::"):
                name = k.split("::")[1]
                assert rel(params[name].grad, g[k]) < 3 * TOL, name
                assert cos(params[name].grad, g[k]) > 0.999, name


def test_text_fusion_heads_vs_reference_golden(cb, golden_dir):
    from clover_b200 import fusion, heads, text
    g = _g(golden_dir, "bert_fusion_heads.npz")
    small = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, vocab_size=1000, max_position_embeddings=64)
    batch = make_batch(3, L=16, seed=2, vocab=1000)
    ids, msk = batch["token_ids"][:, 0].cuda(), batch["input_mask"][:, 0].cuda()
    tb = text.BertFromPretrained(num_hidden_layers=2, **small).cuda().eval()
    load_synth(tb, 21)
    assert rel(tb(ids, msk)["last_hidden_state"], g["bert_last"]) < TOL
    fm = fusion.CrossModalTransformerFromPretrained(img_in_size=96, hidden_size=128, num_frames=2, spacial_tokens=49,
                                                    token_types=2, num_hidden_layers=2, layer_norm_eps=1e-12,
                                                    use_text_cls=True, **{k: v for k, v in small.items() if k != "hidden_size"}).cuda().eval()
    load_synth(fm, 22)
    vt = (named_tensor("fusion_v", (3, 2, 49, 96), 23) * 20).cuda()
    ts = (named_tensor("fusion_t", (3, 16, 128), 24) * 20).cuda()
    o = fm(visual_token=vt, text_input_mask=msk, text_input_embeds=ts)
    assert rel(o["last_hidden_state"], g["fusion_last"]) < TOL
    assert rel(o["t_last_hidden_state"], g["fusion_t_last"]) < TOL
    mh = heads.MLMHead(128, 1000).cuda()
    load_synth(mh, 25)
    assert rel(mh(ts), g["mlm_logits"]) < TOL
    h1 = heads.NCEHeadForMM(visual_in_channels=64, text_in_channels=48, img_hidden_dim=96, vts_embed_dim=40, ln=True,
                            spatial_type="avg", text_agg_type="cls", dropout_ratio=0).cuda()
    load_synth(h1, 26)
    feat = (named_tensor("head_feat", (3, 64, 2, 7, 7), 27) * 20).cuda()
    txt = (named_tensor("head_txt", (3, 16, 48), 28) * 20).cuda()
    v, t = h1(feat, txt)
    assert rel(v, g["mm_v"]) < TOL and rel(t, g["mm_t"]) < TOL
    h2 = heads.NCEHeadForVision(visual_in_channels=48, cross_in_channels=48, hidden_dim=32, ln=True, vts_embed_dim=40,
                                dropout_ratio=0).cuda()
    load_synth(h2, 29)
    assert rel(h2(txt[:, 0]), g["v_head"]) < TOL
    assert rel(h2(txt), g["v_head_seq"]) < TOL
    h3 = heads.NCEHeadForText(cross_in_channels=48, vts_embed_dim=40, text_bn=False, dropout_ratio=0.0).cuda()
    load_synth(h3, 30)
    assert rel(h3(txt[:, 0]), g["t_head"]) < TOL
    q1 = heads.QA_OE_Head(hidden_dim=48, dropout_ratio=0.0, num_labels=30).cuda()
    load_synth(q1, 31)
    assert rel(q1(txt[:, 0]), g["qa_oe"]) < TOL
    q2 = heads.QA_MC_head(48, dropout_ratio=0.0).cuda()
    load_synth(q2, 32)
    assert rel(q2(txt[:, 0]), g["qa_mc"]) < TOL


def test_loss_modules_vs_reference_golden(cb, golden_dir):
    from clover_b200 import losses
    g = _g(golden_dir, "losses.npz")
    for Bg in (6, 33):
        embs = [(named_tensor(f"loss_e{i}_{Bg}", (Bg, 24), 40) * 20).cuda().requires_grad_(True) for i in range(4)]
        lf = losses.ExclusiveNCEwithRankingLoss(temperature=0.05, use_rank=True, use_rank_ttm=True, use_rank_trtm=False, margin_ttm=5.0)
        d = lf(*embs)
        (d["nce_loss"] + d["rank_t_tm_loss"]).backward()
        assert abs(float(d["nce_loss"]) - float(g[f"excl_nce_{Bg}"])) < 1e-4 * abs(float(g[f"excl_nce_{Bg}"]))
        assert abs(float(d["rank_t_tm_loss"]) - float(g[f"excl_rank_{Bg}"])) < 1e-4 * max(1.0, abs(float(g[f"excl_rank_{Bg}"])))
        for i, e in enumerate(embs):
            assert rel(e.grad, g[f"excl_grad{i}_{Bg}"]) < 1e-3
        a = (named_tensor(f"ns_a_{Bg}", (Bg, 24), 41) * 20).cuda().requires_grad_(True)
        b = (named_tensor(f"ns_b_{Bg}", (Bg, 24), 42) * 20).cuda().requires_grad_(True)
        v = losses.NormSoftmaxLoss(temperature=0.05, cos_sim=True)(a, b)
        v.backward()
        assert abs(float(v) - float(g[f"normsoftmax_{Bg}"])) < 1e-4 * abs(float(g[f"normsoftmax_{Bg}"]))
        assert rel(a.grad, g[f"normsoftmax_ga_{Bg}"]) < 1e-3
    logits = (named_tensor("focal_logits", (11, 500), 43) * 60).cuda().requires_grad_(True)
    v = losses.SoftmaxFocalLossMultiClass(gamma=2.0)(logits, torch.from_numpy(g["focal_tgt"]).cuda())
    v.backward()
    assert abs(float(v) - float(g["focal"])) < 1e-5 * abs(float(g["focal"]))
    assert rel(logits.grad, g["focal_grad"]) < 1e-4
    # 30 classes (not a multiple of 4) exercises the generic-width path of the CE kernel
    logits2 = (named_tensor("ce_logits", (7, 30), 45) * 60).cuda().requires_grad_(True)
    v = losses.CrossEntropyLoss()(logits2, torch.from_numpy(g["ce_tgt"]).cuda())
    v.backward()
    assert abs(float(v) - float(g["ce"])) < 1e-5 * abs(float(g["ce"]))
    assert rel(logits2.grad, g["ce_grad"]) < 1e-4


def _pretrain_model(cb, embed, depths, heads, img_in, hidden, vocab, text_layers, fusion_layers, frames_half, bert):
    from clover_b200.configs import pretrain_cfg
    cfg = pretrain_cfg(embed, depths, heads, img_in, hidden, vocab, text_layers, fusion_layers, frames_half, **bert)
    return cb.build_model(cfg).cuda()


def _check_pretrain(model, g, batch, loss_tol):
    kw = {k: batch[k].cuda() for k in ("token_ids", "segment_ids", "input_mask", "mlm_label", "v_token_mask")}
    losses = model(batch["imgs"].cuda(), batch["label"].cuda(), return_loss=True, **kw)
    total, log_vars = model._parse_losses(losses)
    total.backward()
    for k in ("mlm_loss", "nce_loss", "rank_t_tm_loss", "v_nce_loss", "rank_v_vm_loss", "loss"):
        ref = float(g[f"loss::{k}"])
        assert abs(log_vars[k] - ref) <= loss_tol * max(1.0, abs(ref)), (k, log_vars[k], ref)
    params = dict(model.named_parameters())
    worst = {}
    for k in g.files:
        if k.startswith("grad::"):
            name = k.split("::")[1]
            worst[name] = (rel(params[name].grad, g[k]), cos(params[name].grad, g[k]))
        if k.startswith("gradsample::"):
            name = k.split("::")[1]
            idx = torch.from_numpy(g["gradidx::" + name])
            got = params[name].grad.reshape(-1).cpu()[idx]
            worst[name] = (rel(got, g[k]), cos(got, g[k]))
    bad = {n: v for n, v in worst.items() if v[1] < 0.99}
    assert not bad, bad
    import json
    nograd = set(json.loads(str(g["nograd_keys"])))
    ours = {n for n, p in params.items() if p.grad is None}
    assert ours == nograd, (ours ^ nograd)
    return worst


def test_pretrain_step_tiny_vs_reference_golden(cb, golden_dir):
    g = _g(golden_dir, "pretrain_tiny.npz")
    bert = dict(num_attention_heads=2, intermediate_size=256, max_position_embeddings=64, vocab_size=1000)
    m = _pretrain_model(cb, 32, (2, 2), (1, 2), 64, 128, 1000, 2, 2, 2, bert)
    load_synth(m, 50)
    batch = make_batch(3, frames=4, L=16, seed=51, size=56, vocab=1000)
    _check_pretrain(m, g, batch, 2e-2)


def test_pretrain_step_c1_swin_t_vs_reference_golden(cb, golden_dir):
    """BASELINE config 1 shapes (Swin-T + BERT-base + 3-layer fusion, B=2, 8x224x224, L=32) on the B200."""
    g = _g(golden_dir, "pretrain_c1.npz")
    m = _pretrain_model(cb, 96, (2, 2, 6, 2), (3, 6, 12, 24), 768, 768, 30522, 12, 3, 4, {})
    load_synth(m, 60)
    batch = make_batch(2, frames=8, L=32, seed=61, size=224, vocab=30522)
    worst = _check_pretrain(m, g, batch, 2e-2)
    print({k: (round(v[0], 4), round(v[1], 5)) for k, v in worst.items()})
