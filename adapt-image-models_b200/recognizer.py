"""Device-side recognizer glue either side of the backbone.

`Recognizer3D` mirrors the parts of `mmaction/models/recognizers/recognizer3d.py:12-85`,
`recognizers/base.py:160-194,211-244` and `heads/i3d_head.py:53-73` / `heads/base.py:68-108` that sit directly on
the hot path — view flattening, I3D head (T-mean, dropout, FC), cross-entropy, top-k accuracy, multi-view
`average_clip` — but keeps everything on the device: no `.cpu().numpy()` top-k (heads/base.py:90) and no per-iteration
`.item()` (recognizers/base.py:242).  It is thin torch glue around `ViT_CLIP`; the heavy lifting stays in the CUDA
library.  Checkpoint helpers keep `state_dict` compatible with mmcv checkpoints (`mmcv_custom/runner/checkpoint.py:39-42`).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .backbone import ViT_CLIP, _is_trainable_name


class I3DHead(nn.Module):
    """heads/i3d_head.py:9-73: AdaptiveAvgPool3d(1) -> Dropout -> Linear; init normal(std=0.01)."""

    def __init__(self, num_classes: int, in_channels: int, dropout_ratio: float = 0.5, init_std: float = 0.01):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout_ratio) if dropout_ratio else None
        self.fc_cls = nn.Linear(in_channels, num_classes)
        nn.init.normal_(self.fc_cls.weight, std=init_std)
        nn.init.constant_(self.fc_cls.bias, 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.mean(dim=(2, 3, 4))                  # [N, C, T, 1, 1] -> [N, C]
        if self.dropout is not None:
            x = self.dropout(x)
        return self.fc_cls(x)


def top_k_accuracy_device(scores: torch.Tensor, labels: torch.Tensor, topk=(1, 5)):
    """core/evaluation/accuracy.py:90-109 semantics, on the device, returning 0-dim tensors (no host sync)."""
    kmax = min(max(topk), scores.shape[1])
    pred = scores.topk(kmax, dim=1).indices
    hit = pred.eq(labels.view(-1, 1))
    return [hit[:, :min(k, kmax)].any(dim=1).float().mean() for k in topk]


class Recognizer3D(nn.Module):
    def __init__(self, backbone: Dict, cls_head: Dict, test_cfg: Optional[Dict] = None):
        super().__init__()
        backbone = dict(backbone)
        backbone.pop("type", None)
        self.backbone = ViT_CLIP(**backbone)
        head = dict(cls_head)
        head.pop("type", None)
        head.pop("spatial_type", None)
        self.cls_head = I3DHead(**head)
        self.average_clips = (test_cfg or {}).get("average_clips", "prob")
        self.max_testing_views = (test_cfg or {}).get("max_testing_views")
        self.backbone.init_weights()

    # recognizer3d.py:12-29
    def forward_train(self, imgs: torch.Tensor, labels: torch.Tensor) -> Dict[str, torch.Tensor]:
        imgs = imgs.reshape((-1,) + imgs.shape[2:])           # [B, views, C, T, H, W] -> [B*views, C, T, H, W]
        cls_score = self.cls_head(self.backbone(imgs))
        labels = labels.reshape(-1)
        top1, top5 = top_k_accuracy_device(cls_score.detach(), labels)
        return {"loss_cls": F.cross_entropy(cls_score, labels), "top1_acc": top1, "top5_acc": top5}

    # recognizers/base.py:160-194
    def average_clip(self, cls_score: torch.Tensor, num_segs: int) -> torch.Tensor:
        if self.average_clips not in ("score", "prob", None):
            raise ValueError(f"{self.average_clips} is not supported. Currently supported ones are ['score', 'prob', None]")
        if self.average_clips is None:
            return cls_score
        cls_score = cls_score.view(cls_score.shape[0] // num_segs, num_segs, -1)
        if self.average_clips == "prob":
            return F.softmax(cls_score, dim=2).mean(dim=1)
        return cls_score.mean(dim=1)

    # recognizer3d.py:31-85
    @torch.no_grad()
    def forward_test(self, imgs: torch.Tensor) -> torch.Tensor:
        num_segs = imgs.shape[1]
        imgs = imgs.reshape((-1,) + imgs.shape[2:])
        step = self.max_testing_views or imgs.shape[0]
        scores = [self.cls_head(self.backbone(imgs[i:i + step])) for i in range(0, imgs.shape[0], step)]
        return self.average_clip(torch.cat(scores), num_segs)

    @torch.no_grad()
    def forward_test_sharded(self, imgs: torch.Tensor, n_videos: int, shard: str = "videos", group=None) -> torch.Tensor:
        """Multi-GPU test (apis/test.py:54-97 + 159-199).  `imgs` is THIS rank's share:
        shard='videos': [k, views, C, T, H, W] for the videos parallel.shard_indices(n_videos, rank, world); every rank
                        averages its own views and the [k, classes] scores are gathered.
        shard='views' : fewer videos than GPUs - the n_videos * views (video, view) pairs are dealt out the same way,
                        imgs is [k, C, T, H, W]; the raw class scores are gathered first and `average_clip` runs on the
                        gathered [n_videos * views, classes] (softmax-mean needs all views of a video).
        Returns [n_videos, classes] on every rank."""
        from .parallel import gather_scores
        if shard == "videos":
            return gather_scores(self.forward_test(imgs), n_videos, group)
        if shard != "views":
            raise ValueError("shard must be 'videos' or 'views'")
        num_segs = self._views_per_video
        step = self.max_testing_views or imgs.shape[0]
        raw = torch.cat([self.cls_head(self.backbone(imgs[i:i + step])) for i in range(0, imgs.shape[0], step)])
        return self.average_clip(gather_scores(raw, n_videos * num_segs, group), num_segs)

    _views_per_video = 3          # 3-view testing of the K400 configs (vitclip_base_k400.py: ThreeCrop); set per dataset

    def forward(self, imgs, label=None, return_loss=True):
        if return_loss:
            if label is None:
                raise ValueError("Label should not be None.")
            return self.forward_train(imgs, label)
        return self.forward_test(imgs)


# ------------------------------------------------------------------------------------------------ checkpoints (f3)
def trainable_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """Adapters-only checkpoint: the ~11 M (ViT-B) / ~38 M (ViT-L) trained tensors + the head; frozen CLIP weights are
    reproducible from the CLIP checkpoint and need not be stored 30 epochs x N times."""
    out = {}
    for k, v in model.state_dict().items():
        if _is_trainable_name(k) or "cls_head" in k:
            out[k] = v.detach().cpu().clone()
    return out


def load_checkpoint(model: nn.Module, ckpt, strict: bool = False):
    """Accepts an mmcv checkpoint ({'meta','state_dict','optimizer'}, mmcv_custom/runner/checkpoint.py:39-42), a bare
    state_dict, a CLIP `visual.*` dict or an adapters-only dict; strips `module.` / `visual.` prefixes."""
    sd = torch.load(ckpt, map_location="cpu") if isinstance(ckpt, str) else ckpt
    sd = sd.get("state_dict", sd)
    fixed = {}
    own = set(model.state_dict().keys())
    for k, v in sd.items():
        for pfx in ("module.", "visual."):
            if k.startswith(pfx):
                k = k[len(pfx):]
        if k == "proj":                                   # CLIP's projection is dropped (vit_clip.py:373-375)
            continue
        if k not in own and ("backbone." + k) in own:
            k = "backbone." + k
        elif k not in own and k.startswith("backbone.") and k[len("backbone."):] in own:
            k = k[len("backbone."):]
        fixed[k] = v
    return model.load_state_dict(fixed, strict=strict)
