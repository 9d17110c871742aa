"""TEST INFRASTRUCTURE ONLY -- import shim that executes the UNMODIFIED reference modules.

The reference (LeeYN-43/Clover, mounted read-only at /root/reference) cannot be imported as
a package in this image: mmcv, timm, decord, spacy are absent and `transformers` is 5.x while
the reference pins 4.6.1 (install.sh:27).  This shim stubs exactly the third-party names the
hot-path files import, loads each reference file *where it lies* under its real dotted name
(importlib, no copy), and patches `transformers` back to 4.6.1 semantics (random-init instead of
`from_pretrained`, additive -10000 padding mask, eager attention).

It only exists to (a) generate the golden vectors under tests/golden/ (oracle/make_golden.py)
and (b) validate oracle/clover_oracle.py in this container.  /root/reference does not exist on
the GPU box, so nothing in tests -m gpu / smoke() / bench.py may import this file.
Recipe follows SURVEY.md section 8(c).
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("CLOVER_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "mmaction/models/backbones/swin_transformer_3d.py"))


class _Registry:
    """Minimal mmcv.utils.Registry work-alike (register_module / build / get / __contains__)."""

    def __init__(self, name, parent=None, **kw):
        self.name = name
        self._module_dict = {}
        self.parent = parent

    def __contains__(self, key):
        return key in self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        name = name or cls.__name__
        if not force and name in self._module_dict:
            raise KeyError(f"{name} is already registered in {self.name}")
        self._module_dict[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls
        return deco

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop("type")
        cls = self._module_dict[typ] if isinstance(typ, str) else typ
        return cls(**args)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    if "." not in name or True:
        m.__path__ = []  # behave as a package so sub-imports resolve through sys.modules
    sys.modules[name] = m
    return m


def _load(dotted, relpath):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(dotted, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = m
    spec.loader.exec_module(m)
    parent, _, leaf = dotted.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


_LOADED = None
# golden-vector generation shrinks BERT (hidden/heads/intermediate/vocab) through this dict;
# empty = bert-base-uncased defaults, which BertConfig() reproduces.
BERT_OVERRIDES = {}


def load_reference():
    """Returns a namespace with the reference's hot-path modules (unmodified source)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    import torch
    import torch.nn as nn
    import transformers  # noqa: F401  (must be imported before the timm stub exists)
    from transformers import BertConfig, BertModel, BertForPreTraining, BertForMaskedLM
    import logging

    def digit_version(v):
        return tuple(int(x) for x in str(v).split("+")[0].split(".")[:3] if x.isdigit())

    # ---- mmcv stubs -------------------------------------------------------------------------
    mmcv_models = _Registry("mmcv_models")
    _mod("mmcv", __version__="1.3.18", digit_version=digit_version)
    _mod("mmcv.utils", Registry=_Registry, TORCH_VERSION=torch.__version__, digit_version=digit_version,
         print_log=lambda *a, **k: None, get_logger=lambda name, **k: logging.getLogger(name),
         _BatchNorm=nn.modules.batchnorm._BatchNorm, _InstanceNorm=nn.modules.instancenorm._InstanceNorm,
         build_from_cfg=lambda cfg, reg, default_args=None: reg.build(cfg, default_args))
    _mod("mmcv.cnn", MODELS=mmcv_models)

    def get_dist_info():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _load_state_dict(module, state_dict, strict=False, logger=None):
        return module.load_state_dict(state_dict, strict=strict)

    def _load_checkpoint(model, filename, map_location="cpu", strict=False, logger=None):
        ck = torch.load(filename, map_location=map_location)
        sd = ck.get("state_dict", ck)
        model.load_state_dict(sd, strict=strict)
        return ck

    _mod("mmcv.runner", get_dist_info=get_dist_info, load_checkpoint=_load_checkpoint,
         load_state_dict=_load_state_dict)
    _mod("mmcv.runner.dist_utils", allreduce_grads=lambda *a, **k: None)

    # ---- timm stubs (DropPath: per-sample Bernoulli keep mask scaled by 1/keep) -------------
    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            mask = x.new_empty(shape).bernoulli_(keep)
            return x * mask / keep

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=DropPath, trunc_normal_=trunc_normal_)

    # ---- empty mmaction package skeleton ----------------------------------------------------
    def import_module_error_func(name):
        def deco(fn):
            return fn
        return deco

    for pkg in ["mmaction", "mmaction.core", "mmaction.core.hooks", "mmaction.models", "mmaction.models.utils",
                "mmaction.models.backbones", "mmaction.models.heads", "mmaction.models.losses",
                "mmaction.models.recognizers"]:
        _mod(pkg)
    _mod("mmaction.utils", get_root_logger=lambda *a, **k: logging.getLogger("mmaction"),
         import_module_error_func=import_module_error_func)

    ns = types.SimpleNamespace()
    ns.fp16_utils = _load("mmaction.core.hooks.fp16_utils", "mmaction/core/hooks/fp16_utils.py")
    sys.modules["mmcv.runner"].force_fp32 = ns.fp16_utils.force_fp32
    sys.modules["mmcv.runner"].auto_fp16 = ns.fp16_utils.auto_fp16
    ns.builder = _load("mmaction.models.builder", "mmaction/models/builder.py")
    ns.gather_loss = _load("mmaction.models.utils.gather_loss", "mmaction/models/utils/gather_loss.py")

    # ---- transformers patches: 4.6.1 semantics, offline random init -------------------------
    def _cfg_from_pretrained(name=None, **kw):
        kw.pop("config", None)
        kw.update(BERT_OVERRIDES)
        cfg = BertConfig(**kw)
        cfg._attn_implementation = "eager"
        return cfg

    BertConfig.from_pretrained = staticmethod(_cfg_from_pretrained)

    def _model_from_pretrained(cls):
        def f(name=None, config=None, **kw):
            if config is None:
                config = _cfg_from_pretrained()
            config._attn_implementation = "eager"
            return cls(config)
        return classmethod(lambda c, *a, **k: f(*a, **k))

    for cls in (BertModel, BertForPreTraining, BertForMaskedLM):
        cls.from_pretrained = _model_from_pretrained(cls)

    def _ext_mask(self, attention_mask, input_shape=None, device=None, dtype=None):
        # transformers 4.6.1 modeling_utils.get_extended_attention_mask: (1 - m) * -10000.0
        m = attention_mask[:, None, None, :].to(torch.float32 if dtype is None or not isinstance(dtype, torch.dtype) else dtype)
        return (1.0 - m) * -10000.0

    transformers.modeling_utils.ModuleUtilsMixin.get_extended_attention_mask = _ext_mask

    def _create_attention_masks(self, attention_mask, encoder_attention_mask, embedding_output,
                                encoder_hidden_states, past_key_values):
        # 4.6.1 BertModel.forward: extended mask = (1 - mask[:, None, None, :]) * -10000.0
        if attention_mask is None:
            attention_mask = torch.ones(embedding_output.shape[:2], device=embedding_output.device)
        ext = (1.0 - attention_mask[:, None, None, :].to(embedding_output.dtype)) * -10000.0
        return ext, None

    BertModel._create_attention_masks = _create_attention_masks

    # 4.6.1 BertEncoder.forward(..., output_attentions=True) returns every layer's attention probabilities under
    # 'attentions' (read by finetune.py:192); 5.5's encoder drops the flag, so collect them with forward hooks.
    from transformers.models.bert.modeling_bert import BertEncoder
    _enc_forward = BertEncoder.forward

    def _enc_forward_461(self, hidden_states, attention_mask=None, *a, output_attentions=False, **kw):
        if not output_attentions:
            return _enc_forward(self, hidden_states, attention_mask, *a, **kw)
        probs = []
        hooks = [l.attention.self.register_forward_hook(lambda m, i, o: probs.append(o[1])) for l in self.layer]
        try:
            out = _enc_forward(self, hidden_states, attention_mask, *a, **kw)
        finally:
            for h in hooks:
                h.remove()
        out["attentions"] = tuple(probs)
        return out

    BertEncoder.forward = _enc_forward_461

    ns.swin = _load("mmaction.models.backbones.swin_transformer_3d", "mmaction/models/backbones/swin_transformer_3d.py")
    ns.bert = _load("mmaction.models.backbones.bert_from_hugface", "mmaction/models/backbones/bert_from_hugface.py")
    ns.cross = _load("mmaction.models.backbones.cross_transformer", "mmaction/models/backbones/cross_transformer.py")
    ns.ssl_head = _load("mmaction.models.heads.ssl_head", "mmaction/models/heads/ssl_head.py")
    ns.mlm_head = _load("mmaction.models.heads.mlm_itm_head", "mmaction/models/heads/mlm_itm_head.py")
    ns.qa_head = _load("mmaction.models.heads.qa_head", "mmaction/models/heads/qa_head.py")
    ns.loss_base = _load("mmaction.models.losses.base", "mmaction/models/losses/base.py")
    ns.ce_loss = _load("mmaction.models.losses.cross_entropy_loss", "mmaction/models/losses/cross_entropy_loss.py")
    ns.focal_loss = _load("mmaction.models.losses.focal_loss", "mmaction/models/losses/focal_loss.py")
    ns.contrastive = _load("mmaction.models.losses.contrastive_loss", "mmaction/models/losses/contrastive_loss.py")
    ns.rec_base = _load("mmaction.models.recognizers.base", "mmaction/models/recognizers/base.py")
    ns.pretrain = _load("mmaction.models.recognizers.multimodal_transformer_pretrain",
                        "mmaction/models/recognizers/multimodal_transformer_pretrain.py")
    ns.finetune = _load("mmaction.models.recognizers.multimodal_transformer_finetune",
                        "mmaction/models/recognizers/multimodal_transformer_finetune.py")
    ns.MODELS = ns.builder.MODELS

    # App. D1 workaround (documented in SURVEY.md): NCEHeadForVision.forward begins with
    # img.mean(dim=1) (ssl_head.py:210) but pretrain.py:148-149 feeds it (B,768).  The oracle
    # decision is "mean over a singleton": accept (B,C) by unsqueezing.
    _orig_v_forward = ns.ssl_head.NCEHeadForVision.forward

    def _v_forward(self, img):
        if img.dim() == 2:
            img = img.unsqueeze(1)
        return _orig_v_forward(self, img)

    ns.ssl_head.NCEHeadForVision.forward = _v_forward
    _LOADED = ns
    return ns


def ensure_gloo_group():
    """The reference's losses call dist.all_gather unconditionally (gather_loss.py:49)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
