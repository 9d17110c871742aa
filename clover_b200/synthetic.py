"""Deterministic synthetic weights and batches (numpy PCG64 streams keyed by tensor name).

The reference has no datasets or checkpoints reachable offline, so every test, golden vector and
bench run uses random-init weights and synthetic clips/captions of the shapes in SURVEY.md 8(d).
numpy's default_rng stream is stable across platforms and versions, so the same (name, seed) gives
bit-identical tensors in the build container and on the GPU box.
"""
import zlib

import numpy as np
import torch

CLS, SEP, MASK, PAD = 101, 102, 103, 0


def _rng(name, seed):
    return np.random.default_rng([zlib.crc32(name.encode()), seed])


def named_tensor(name, shape, seed=0, dtype=torch.float32):
    """LayerNorm-like weights ~ 1 + 0.1 N(0,1); biases / tables ~ 0.02 N(0,1); matrices ~ N(0, 1/fan_in) (capped)."""
    shape = tuple(shape)
    r = _rng(name, seed)
    leaf = name.rsplit(".", 1)[-1]
    x = r.standard_normal(shape, dtype=np.float32)
    is_norm = any(t in name for t in ("norm", "LayerNorm", "_bn", "img_projector.1.", "img_projector.4.",
                                      "vqa_classifier.2.", "mc_vqa_classifier.2."))
    if is_norm and leaf == "weight" and len(shape) == 1:
        x = 1.0 + 0.1 * x
    elif len(shape) >= 2 and leaf == "weight" and "embeddings" not in name:
        fan_in = int(np.prod(shape[1:]))
        x = x * min(0.08, 1.0 / np.sqrt(fan_in)) * 1.5
    elif "relative_position_bias_table" in name:
        x = x * 0.5
    else:
        x = x * 0.05
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def synth_state_dict(shapes, seed=0):
    """shapes: mapping name -> shape (or a module's state_dict()).  Integer buffers are left out."""
    out = {}
    for k, v in shapes.items():
        if hasattr(v, "shape"):
            if not torch.is_floating_point(v):
                continue
            v = v.shape
        out[k] = named_tensor(k, v, seed)
    return out


def make_batch(B, frames=8, L=32, seed=1, size=224, vocab=30522, mask_cells=10, grid=7):
    """Synthetic pre-training batch with the reference's input contract:
    imgs (B,1,3,F,size,size) ~ N(0,1); token_ids/input_mask/segment_ids/mlm_label (B,1,L) int64;
    v_token_mask (B,1,grid,grid) int64 with a block of ~mask_cells ones."""
    r = np.random.default_rng([seed, 77])
    imgs = r.standard_normal((B, 1, 3, frames, size, size), dtype=np.float32)
    tok = np.zeros((B, 1, L), dtype=np.int64)
    lab = np.full((B, 1, L), -100, dtype=np.int64)
    lo = min(1000, vocab // 2)
    for b in range(B):
        n = int(r.integers(min(8, L - 2), L - 1))
        ids = r.integers(lo, vocab - 1, size=n)
        tok[b, 0, 0] = CLS if vocab > CLS else 1
        tok[b, 0, 1:1 + n] = ids
        tok[b, 0, 1 + n] = SEP if vocab > SEP else 2
        k = max(1, int(round(0.3 * n)))
        pos = r.permutation(n)[:k] + 1
        lab[b, 0, pos] = tok[b, 0, pos]
        tok[b, 0, pos] = MASK if vocab > MASK else 3
    vm = np.zeros((B, 1, grid, grid), dtype=np.int64)
    for b in range(B):
        h = 2
        w = max(1, mask_cells // h)
        y0 = int(r.integers(0, grid - h + 1))
        x0 = int(r.integers(0, grid - w + 1))
        vm[b, 0, y0:y0 + h, x0:x0 + w] = 1
    return {
        "imgs": torch.from_numpy(imgs),
        "label": torch.zeros(B),
        "token_ids": torch.from_numpy(tok),
        "input_mask": torch.from_numpy((tok != 0).astype(np.int64)),
        "segment_ids": torch.zeros(B, 1, L, dtype=torch.long),
        "mlm_label": torch.from_numpy(lab),
        "v_token_mask": torch.from_numpy(vm),
    }


def make_finetune_batch(task, B, frames=16, size=224, L=40, vocab=30522, seed=1, num_labels=1500, choices=5):
    """Fine-tune batches.  'retrieval' and 'video_qa' (open-ended) carry one caption / question per
    clip ('FIB': with exactly one [MASK] token); 'video_qa_mc' carries `choices` candidate sentences per clip (token_ids (B, choices, L)) and label = the
    index of the right one.  No MLM masking is used by the fine-tune recognisers, the [MASK] ids are just tokens."""
    if task == "video_qa_mc":
        parts = [make_batch(B, frames=frames, L=L, seed=seed + 10 * c, size=size, vocab=vocab) for c in range(choices)]
        batch = {"imgs": parts[0]["imgs"]}
        for k in ("token_ids", "input_mask", "segment_ids"):
            batch[k] = torch.cat([p[k] for p in parts], 1)
        batch["label"] = torch.from_numpy(np.random.default_rng([seed, 5]).integers(0, choices, size=(B, 1)))
        return batch
    b = make_batch(B, frames=frames, L=L, seed=seed, size=size, vocab=vocab)
    batch = {k: b[k] for k in ("imgs", "token_ids", "input_mask", "segment_ids")}
    if task == "FIB":
        # fill-in-the-blank: exactly ONE [MASK] per sentence (finetune.py:98-100 picks the fused states at token id 103 and
        # pairs them with one label per sample); keep the first masked position of make_batch, restore the others
        lab = b["mlm_label"]
        tok = torch.where(lab == -100, b["token_ids"], lab)
        first = (lab != -100).int().argmax(dim=-1, keepdim=True)
        batch["token_ids"] = tok.scatter(-1, first, b["token_ids"].gather(-1, first))
    hi = num_labels if task in ("video_qa", "FIB") else 1
    batch["label"] = torch.from_numpy(np.random.default_rng([seed, 5]).integers(0, hi, size=(B, 1)))
    return batch


def _swin_shapes(embed, depths, heads):
    sh = {"patch_embed.proj.weight": (embed, 3, 2, 4, 4), "patch_embed.proj.bias": (embed,),
          "patch_embed.norm.weight": (embed,), "patch_embed.norm.bias": (embed,)}
    for s, (d, nh) in enumerate(zip(depths, heads)):
        C = embed * 2 ** s
        for j in range(d):
            p = f"layers.{s}.blocks.{j}."
            sh.update({p + "norm1.weight": (C,), p + "norm1.bias": (C,), p + "attn.relative_position_bias_table": (2535, nh),
                       p + "attn.qkv.weight": (3 * C, C), p + "attn.qkv.bias": (3 * C,), p + "attn.proj.weight": (C, C),
                       p + "attn.proj.bias": (C,), p + "norm2.weight": (C,), p + "norm2.bias": (C,),
                       p + "mlp.fc1.weight": (4 * C, C), p + "mlp.fc1.bias": (4 * C,), p + "mlp.fc2.weight": (C, 4 * C),
                       p + "mlp.fc2.bias": (C,)})
        if s < len(depths) - 1:
            p = f"layers.{s}.downsample."
            sh.update({p + "reduction.weight": (2 * C, 4 * C), p + "norm.weight": (4 * C,), p + "norm.bias": (4 * C,)})
    Cf = embed * 2 ** (len(depths) - 1)
    sh.update({"norm.weight": (Cf,), "norm.bias": (Cf,)})
    return sh


def synth_swin2d_checkpoint(path, embed=32, depths=(2, 2), heads=(1, 2), window2d=6, seed=90):
    """A synthetic 2-D Swin checkpoint (the layout of the ImageNet files inflate_weights expects): 2-D patch-embed kernel,
    (2*window2d-1)^2 bias tables (window 6 -> 11x11, so the 13x13 target needs the bicubic resize), stale index buffers."""
    sh = _swin_shapes(embed, list(depths), list(heads))
    sd = {}
    for k, shape in sh.items():
        if k == "patch_embed.proj.weight":
            shape = (embed, 3, 4, 4)
        elif k.endswith("relative_position_bias_table"):
            shape = ((2 * window2d - 1) ** 2, shape[1])
        sd[k] = named_tensor("ck2d." + k, shape, seed)
    sd["layers.0.blocks.0.attn.relative_position_index"] = torch.zeros(36, 36, dtype=torch.long)
    sd["layers.0.blocks.1.attn_mask"] = torch.zeros(4, 36, 36)
    torch.save({"state_dict": sd}, path)
    return sd


