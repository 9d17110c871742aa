"""CPU-only host-logic tests: product-side integer tables are bit-exact against the reference's
golden tables; the C-ABI library loads and exports every symbol include/clover_b200.h declares."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from clover_b200 import tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_tables_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    assert np.array_equal(tables.relative_position_index((8, 7, 7)), g["rel_index_877"].astype(np.int64))
    assert np.array_equal(tables.relative_position_index((2, 7, 7)), g["rel_index_277"].astype(np.int64))
    for N in (196, 392, 98):
        code, off = tables.rel_code(N, (8, 7, 7))
        idx = code.astype(np.int64)[:, None] - code.astype(np.int64)[None, :] + off
        assert np.array_equal(idx, g["rel_index_877"].astype(np.int64)[:N, :N])
    meta = json.loads(str(g["meta"]))
    for i, m in enumerate(meta):
        win, sh = tables.get_window_size(tuple(m["dims"]), tuple(m["window_cfg"]), tuple(m["shift_cfg"]))
        assert list(win) == m["window"] and list(sh) == m["shift"]
        D, H, W = m["dims"]
        rid = tables.region_ids(D, H, W, win, sh)
        mask = tables.attn_mask_from_regions(rid)
        shape = tuple(g[f"mask_shape_{i}"])
        ref = np.unpackbits(g[f"mask_bits_{i}"])[: int(np.prod(shape))].reshape(shape).astype(bool)
        assert np.array_equal(mask != 0, ref)
        gather = g[f"gather_{i}"].astype(np.int64)
        assert np.array_equal(tables.window_gather_index(2, D, H, W, win, sh), gather)
        if D % win[0] == 0 and H % win[1] == 0 and W % win[2] == 0:
            # the row map of the LN kernels is the inverse of the reference's roll + window_partition permutation
            rmap = tables.window_row_map(D, H, W, win, sh)
            clip0 = gather.reshape(-1)[: D * H * W]
            assert np.array_equal(rmap[clip0], np.arange(D * H * W))


def _kernel_constants():
    src = open(os.path.join(ROOT, "clover_b200", "csrc", "attention_w7.cu")).read()
    m = re.search(r"constexpr int W7_PH = (\d+), W7_PW = (\d+);", src)
    assert m, "staged-table strides not found in attention_w7.cu"
    return int(m.group(1)), int(m.group(2))


@pytest.mark.parametrize("wd", [2, 4, 6, 8])
def test_w7_staged_bias_table_layout(golden_dir, wd):
    """The (wd,7,7) window-attention kernels look the relative-position bias up in a re-laid-out copy of the table
    (attn_w7_table_kernel / w7_code in csrc/attention_w7.cu: strides W7_PH / W7_PW instead of the reference's 169 / 13).
    Restated here from the kernel's formulas with the constants read from the source: every (i, j) lookup must land on the
    entry the reference's relative_position_index names (swin_transformer_3d.py:345-359), inside the staged buffer, and the 32
    lanes of any warp (32 consecutive rows) must hit 32 different shared-memory banks."""
    PH, PW = _kernel_constants()
    cfg_wd, SH, SW = 8, 169, 13
    N = 49 * wd
    ref = np.load(os.path.join(golden_dir, "tables.npz"))["rel_index_877"].astype(np.int64)[:N, :N]
    ld = (2 * (wd - 1) * PH + 12 * PW + 13 + 3) & ~3                    # w7_table_ld
    x = np.arange(ld)
    zz, rem = x // PH, x % PH
    yy, xx = rem // PW, rem % PW
    used = (zz <= 2 * (wd - 1)) & (yy <= 12) & (xx <= 12)
    staged = np.where(used, (zz - (wd - 1) + cfg_wd - 1) * SH + yy * SW + xx, -1)   # attn_w7_table_kernel
    t = np.arange(N)
    pcode = (t // 49) * PH + ((t % 49) // 7) * PW + t % 7              # w7_code / per-thread base
    addr = pcode[:, None] - pcode[None, :] + (wd - 1) * PH + 6 * PW + 6  # a.code_off
    assert addr.min() >= 0 and addr.max() < ld
    assert np.array_equal(staged[addr], ref)
    for r0 in range(N - 31):
        assert len(set((pcode[r0:r0 + 32] % 32).tolist())) == 32


def test_tunable_names_match_enum():
    """clv_set_tunable resolves names by position: the name table in runtime.cu must list exactly the enumerators of
    common.cuh's Tunable enum, in order, and every name must be accepted by the built library."""
    csrc = os.path.join(ROOT, "clover_b200", "csrc")
    enum = re.search(r"enum Tunable \{([^}]*)\}", open(os.path.join(csrc, "common.cuh")).read()).group(1)
    ids = [e.split("=")[0].strip() for e in enum.split(",") if e.strip()]
    assert ids[-1] == "TUNE_COUNT"
    names = re.search(r"TUNE_NAMES\[TUNE_COUNT\] = \{([^}]*)\}", open(os.path.join(csrc, "runtime.cu")).read()).group(1)
    names = re.findall(r'"([a-z0-9_]+)"', names)
    assert [i[len("TUNE_"):].lower() for i in ids[:-1]] == names
    from clover_b200 import _lib
    lib = _lib.load()
    for n in names:
        assert lib.clv_set_tunable(n.encode(), -1) == 0
    assert lib.clv_set_tunable(b"no_such_knob", 1) != 0


def test_library_exports_every_declared_symbol():
    from clover_b200 import _lib
    header = open(os.path.join(ROOT, "include", "clover_b200.h")).read()
    declared = set(re.findall(r"\b(clv_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} has not been built (run __graft_entry__.build())")
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.clv_version() >= 100
    assert lib.clv_launch_count() == 0
    assert lib.clv_nce_workspace_floats(3, 4, 8) > 0


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from clover_b200 import ops
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, a, torch.zeros(128, 128))


def test_dropout_stream_host_matches_numpy_restatement():
    """Integer work, bit-exact: the library's host evaluation of the dropout stream (the same inline function the kernels
    call) against tests/rng_ref.py, incl. 64-bit wrap-around of seed * phi + idx, and the keep thresholds."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import rng_ref
    from clover_b200 import _lib
    lib = _lib.load()
    for seed in (0, 1, 12345, 2 ** 63 + 7, 2 ** 64 - 1):
        idx = np.concatenate([np.arange(64, dtype=np.uint64), np.array([2 ** 32 - 1, 2 ** 32, 2 ** 40 + 3, 2 ** 64 - 5], dtype=np.uint64)])
        want = rng_ref.rand_u32(seed, idx)
        got = np.array([lib.clv_rand_u32(seed, int(i)) for i in idx], dtype=np.uint32)
        assert np.array_equal(got, want), seed
    for p in (0.0, 0.1, 0.3, 0.5, 0.999):
        assert lib.clv_dropout_threshold(p) == int(rng_ref.threshold(p)), p
    # the stream is uniform enough for dropout: keep rate within 1% of 1 - p over 1e5 draws
    for p in (0.1, 0.5):
        assert abs(rng_ref.keep_mask(100000, p, 99, 4096).mean() - (1 - p)) < 0.01


def test_rng_stream_offsets_are_disjoint_and_logged():
    from clover_b200 import rng
    rng.manual_seed(5)
    rng.LOG = []
    try:
        a = rng.next_stream(10, "x", (10,), 0.1)
        b = rng.next_stream(7, "y", (7,), 0.2)
        c = rng.next_stream(4, "z", (4,), 0.3)
    finally:
        log, rng.LOG = rng.LOG, None
    assert a == (5, 0) and b == (5, 12) and c == (5, 20)
    assert [e["kind"] for e in log] == ["x", "y", "z"] and log[1]["offset"] == 12


def test_param_groups_follow_the_shipped_paramwise_cfg():
    """mmcv DefaultOptimizerConstructor semantics for configs/exp_local/pretrain_webvid_cc3m.py:129-134 and
    finetune_msrvttQA.py:90-97: no decay on norms / biases / bias tables, qa_head lr x10."""
    import torch
    from clover_b200 import registry
    from clover_b200.configs import finetune_cfg
    from clover_b200.optim import param_groups_from_cfg
    registry.register_all()
    m = registry.build_model(finetune_cfg("video_qa", embed=32, depths=(2, 2), heads=(1, 2), img_in=64, hidden=128, vocab=1000,
                                          text_layers=1, fusion_layers=1, frames_half=8, num_labels=50, num_attention_heads=2,
                                          intermediate_size=256, max_position_embeddings=64, vocab_size=1000))
    cfg = dict(norm_decay_mult=0.0, bias_decay_mult=0.0,
               custom_keys={"relative_position_bias_table": dict(decay_mult=0.0), "qa_head": dict(lr_mult=10)})
    g = {x["name"]: x for x in param_groups_from_cfg(m, 1e-4, 0.05, cfg)}
    assert g["backbone.layers.0.blocks.0.attn.relative_position_bias_table"]["weight_decay"] == 0.0
    assert g["backbone.layers.0.blocks.0.norm1.weight"]["weight_decay"] == 0.0
    assert g["backbone.layers.0.blocks.0.attn.qkv.bias"]["weight_decay"] == 0.0
    assert g["backbone.layers.0.blocks.0.attn.qkv.weight"]["weight_decay"] == 0.05
    assert abs(g["qa_head.vqa_classifier.1.weight"]["lr"] - 1e-3) < 1e-12 and g["qa_head.vqa_classifier.1.weight"]["weight_decay"] == 0.05
    assert g["text_backbone.bert.embeddings.LayerNorm.weight"]["weight_decay"] == 0.0
    assert len(g) == sum(1 for p in m.parameters() if p.requires_grad)


def test_gradient_sink_bookkeeping_cpu():
    """Host logic of the gradient sinks (clover_b200.functional): a stashed buffer is handed out once as a fresh alias, shapes and
    32-byte alignment are checked, accumulating sinks are cleared at stash time once their producer asked for zeros."""
    import torch
    from clover_b200 import functional as Fn
    Fn.clear_grad_sinks()
    p = torch.nn.Parameter(torch.zeros(8, 16))
    q = torch.nn.Parameter(torch.zeros(16))
    gp, gq = torch.ones(8, 16), torch.full((16,), 3.0)
    Fn.stash_grad_sinks([(p, gp), (q, gq)])
    a = Fn._sink(Fn._pkey(p), p.shape)
    assert a is not gp and a.data_ptr() == gp.data_ptr() and torch.equal(a, gp)          # alias, not the stashed object
    assert Fn._sink(Fn._pkey(p), p.shape) is None                                        # consumed: a second use allocates
    z = Fn._sink(Fn._pkey(q), q.shape, zero=True)
    assert z.data_ptr() == gq.data_ptr() and float(z.abs().sum()) == 0.0                 # first zero request clears in place
    gq.fill_(5.0)
    Fn.stash_grad_sinks([(p, gp), (q, gq)])                                              # ... later stashes clear it in bulk
    assert float(gq.abs().sum()) == 0.0 and float(gp.sum()) == 128.0
    assert Fn._sink(Fn._pkey(p), (4, 32)) is None                                        # shape mismatch -> no sink
    frozen = torch.nn.Parameter(torch.zeros(4), requires_grad=False)
    assert Fn._pkey(frozen) is None and Fn._sink(None, (4,)) is None
    Fn.stash_grad_sinks([(q, torch.zeros(17)[1:])])                                      # 4-byte-offset view: not 32-byte aligned
    assert Fn._sink(Fn._pkey(q), q.shape) is None
    Fn.clear_grad_sinks()


def test_drop_path_predraw_cpu():
    """rng.predraw_drop_path: one uniform draw serves the following drop_path_scales calls in order; values are 0 or 1 / keep;
    a call that does not match the pre-drawn sequence falls back to an individual draw."""
    import torch
    from clover_b200 import rng
    torch.manual_seed(3)
    ps = [0.0, 0.1, 0.1, 0.3, 0.3]
    rng.predraw_drop_path(64, ps, "cpu")
    outs = [rng.drop_path_scales(64, p, "cpu") for p in ps]
    for p, s in zip(ps, outs):
        keep = 1.0 - p
        assert s.shape == (64,) and s.is_contiguous()
        assert bool(((s == 0) | ((s - 1.0 / keep).abs() < 1e-6)).all())
    assert bool((outs[0] == 1).all())
    assert 0.4 < float((outs[3] > 0).float().mean()) <= 1.0
    rng.predraw_drop_path(64, [0.2, 0.2], "cpu")
    s = rng.drop_path_scales(32, 0.2, "cpu")                                             # other batch size: individual draw
    assert s.shape == (32,) and not rng._PREDRAWN
