"""Drop-in registration.

The reference resolves every model component by its class-name string through ONE registry
(mmaction/models/builder.py:8-14: MODELS aliased as BACKBONES/HEADS/LOSSES/RECOGNIZERS).
``register_all()`` force-registers the clover_b200 classes under the reference's own names into
that registry when mmaction is importable, so the configs under configs/_base_ and
configs/exp_local select them unchanged; in environments without mmcv (this image) an
API-compatible local Registry is used instead.
"""
import inspect


class Registry:
    """Minimal mmcv.utils.Registry work-alike: register_module(name, force, module) / build / get."""

    def __init__(self, name, parent=None):
        self.name = name
        self._module_dict = {}

    def __contains__(self, key):
        return key in self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        if not inspect.isclass(cls):
            raise TypeError(f"module must be a class, got {type(cls)}")
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
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise KeyError(f"cfg must be a dict with a 'type' key, got {cfg!r}")
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**args)


MODELS = Registry("models")
BACKBONES = HEADS = LOSSES = RECOGNIZERS = MODELS
MODULE_HOOKS = Registry("module_hooks")


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_model(cfg, train_cfg=None, test_cfg=None):
    return RECOGNIZERS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg)
                             if (train_cfg is not None or test_cfg is not None) else None)


def plugin_classes():
    from . import fusion, heads, losses, recognizers, swin, text
    return {
        "SwinTransformer3D": swin.SwinTransformer3D,
        "BertFromPretrained": text.BertFromPretrained,
        "CrossModalTransformerFromPretrained": fusion.CrossModalTransformerFromPretrained,
        "NCEHeadForMM": heads.NCEHeadForMM,
        "NCEHeadForVision": heads.NCEHeadForVision,
        "NCEHeadForText": heads.NCEHeadForText,
        "MLMHead": heads.MLMHead,
        "ITMHead": heads.ITMHead,
        "QA_OE_Head": heads.QA_OE_Head,
        "QA_MC_head": heads.QA_MC_head,
        "ExclusiveNCEwithRankingLoss": losses.ExclusiveNCEwithRankingLoss,
        "NormSoftmaxLoss": losses.NormSoftmaxLoss,
        "SoftmaxFocalLossMultiClass": losses.SoftmaxFocalLossMultiClass,
        "CrossEntropyLoss": losses.CrossEntropyLoss,
        "CloverPretrain": recognizers.CloverPretrain,
        "CloverFinetune": recognizers.CloverFinetune,
    }


def register_all(target=None, force=True):
    """Register every plugin class under the reference's names.  ``target`` defaults to mmaction's
    MODELS registry when importable, else the local one.  Returns the registry used."""
    if target is None:
        try:
            from mmaction.models.builder import MODELS as target  # the reference's registry
        except Exception:
            target = MODELS
    for name, cls in plugin_classes().items():
        target.register_module(name=name, force=force, module=cls)
    if target is not MODELS:
        for name, cls in plugin_classes().items():
            MODELS.register_module(name=name, force=True, module=cls)
    # the GPUNormalize module hook lives in its own registry (mmaction/utils/module_hooks.py:4,35)
    from .swin import GPUNormalize
    try:
        from mmaction.utils.module_hooks import MODULE_HOOKS as hooks
        hooks.register_module(name="GPUNormalize", force=force, module=GPUNormalize)
    except Exception:
        pass
    MODULE_HOOKS.register_module(name="GPUNormalize", force=True, module=GPUNormalize)
    return target
