"""aimb200 — B200-native (sm_100a) implementation of the AIM ``ViT_CLIP`` backbone hot path.

The directory is named ``adapt-image-models_b200`` (not importable as such); import it as ``aimb200``
(the sibling ``aimb200/`` shim points its ``__path__`` here).
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
from .registry import BACKBONES, build_backbone  # noqa: E402,F401
from .backbone import ViT_CLIP, AIM, ViT_ImageNet  # noqa: E402,F401
from .config import load_config, backbone_cfg  # noqa: E402,F401
from .parallel import GradSync, shard_indices, gather_scores  # noqa: E402,F401
from .graphs import GraphedStep  # noqa: E402,F401
from .optim import FlatAdamW  # noqa: E402,F401
from .recognizer import Recognizer3D, I3DHead, trainable_state_dict, load_checkpoint  # noqa: E402,F401

__all__ += ["BACKBONES", "build_backbone", "ViT_CLIP", "AIM", "ViT_ImageNet", "load_config", "backbone_cfg", "GradSync", "shard_indices", "gather_scores", "GraphedStep", "FlatAdamW", "Recognizer3D", "I3DHead", "trainable_state_dict", "load_checkpoint"]
