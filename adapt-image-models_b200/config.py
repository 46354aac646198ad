"""Minimal reader of the reference's Python config files (``configs/recognition/vit/vitclip_*.py``).

The reference parses them with ``mmcv.Config.fromfile`` (tools/train.py:81): execute the file, load every path in
``_base_`` recursively, and merge the child's dicts *into* the base dicts key by key (``_delete_=True`` in a child
dict replaces instead of merging).  mmcv is not installed in this image, so the boundary carries this ~40-line
restatement of exactly those rules; ``backbone_cfg(path)`` returns the ``model.backbone`` dict that
``BACKBONES.build`` receives, so the in-tree configs drop in unchanged.
"""
from __future__ import annotations

import os
from typing import Any, Dict

DELETE_KEY = "_delete_"
BASE_KEY = "_base_"


def _merge(child: Dict[str, Any], base: Dict[str, Any]) -> Dict[str, Any]:
    """mmcv Config._merge_a_into_b: dict values merge recursively unless the child dict carries _delete_=True."""
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get(DELETE_KEY, False):
            out[k] = _merge(v, out[k])
        elif isinstance(v, dict):
            out[k] = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
        else:
            out[k] = v
    return out


def load_config(path: str) -> Dict[str, Any]:
    """Execute a config file and resolve its ``_base_`` chain; returns the merged top-level dict."""
    path = os.path.abspath(path)
    ns: Dict[str, Any] = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)          # config files are plain Python assignments
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not callable(v) and not isinstance(v, type(os))}
    bases = cfg.pop(BASE_KEY, [])
    if isinstance(bases, str):
        bases = [bases]
    merged: Dict[str, Any] = {}
    for b in bases:
        sub = load_config(os.path.join(os.path.dirname(path), b))
        dup = set(sub) & set(merged)
        if dup:
            raise KeyError(f"duplicate keys in the _base_ files of {path}: {sorted(dup)}")   # as mmcv does
        merged.update(sub)
    return _merge(cfg, merged)


def backbone_cfg(path: str) -> Dict[str, Any]:
    """The ``model.backbone`` dict of a recognizer config: what ``build_backbone`` is called with (recognizers/base.py:75)."""
    return dict(load_config(path)["model"]["backbone"])
