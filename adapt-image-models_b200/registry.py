"""The drop-in boundary: ``BACKBONES.build(dict(type='ViT_CLIP', ...))`` (mmaction/models/builder.py:8-9,27-29).

When mmaction + mmcv are importable, ``ViT_CLIP`` is registered into the real
``mmaction.models.builder.BACKBONES`` (``force=True``, replacing the stock class), so
``build_model(cfg.model)`` picks it up with no config change.  Otherwise a minimal registry with the
same ``register_module()`` / ``build(cfg)`` contract is used (mmcv is not installed in this image).
"""
from __future__ import annotations


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError("cfg must be a dict with the key 'type'")
        args = dict(cfg)
        typ = args.pop("type")
        cls = typ if isinstance(typ, type) else self.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**args)


class _DualRegistry(Registry):
    """Registers locally and, when available, into mmaction's own BACKBONES."""

    def __init__(self, name):
        super().__init__(name)
        self.mm = None
        try:
            from mmaction.models.builder import BACKBONES as _MM  # type: ignore
            if hasattr(_MM, "register_module") and hasattr(_MM, "build"):
                self.mm = _MM
        except Exception:
            self.mm = None

    def register_module(self, name=None, force=True, module=None):
        local = super().register_module(name=name, force=True, module=module)

        def _reg(cls):
            cls = local(cls) if module is None else cls
            if self.mm is not None:
                try:
                    self.mm.register_module(name=name, force=True, module=cls)
                except TypeError:
                    self.mm.register_module(force=True)(cls)
            return cls
        if module is not None:
            return _reg(module)
        return _reg


BACKBONES = _DualRegistry("backbone")


def build_backbone(cfg):
    """mmaction/models/builder.py:27-29"""
    return BACKBONES.build(cfg)
