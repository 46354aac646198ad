"""Writes tests/golden/vitclip_configs.json: the `model` dict of every configs/recognition/vit/vitclip_*.py (and the two
AIM/*.py configs of the upstream-math class) of the reference, parsed by aimb200.config (exec + recursive _base_ merge, the
rules of mmcv.Config.fromfile).  Run in the build container, where /root/reference is mounted:
    python tests/golden/make_config_fixture.py
tests/test_boundary.py re-parses the real files whenever the reference tree is present and compares with this fixture."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import aimb200  # noqa: E402

REF = os.environ.get("AIM_REFERENCE", "/root/reference")


def collect(ref=REF):
    out = {}
    files = sorted(glob.glob(os.path.join(ref, "configs/recognition/vit/vitclip_*.py")))
    files += [os.path.join(ref, "configs/recognition/vit/AIM", f) for f in ("AIM_base_diving48.py", "AIM_base_hmdb51.py")]
    for f in files:
        m = aimb200.load_config(f)["model"]
        out[os.path.relpath(f, ref)] = {"backbone": m["backbone"], "cls_head": m["cls_head"], "test_cfg": m.get("test_cfg")}
    return out


def collect_imagenet(ref=REF):
    """configs/recognition/vit/vit_imagenet_*.py (type='ViT_ImageNet')."""
    out = {}
    for f in sorted(glob.glob(os.path.join(ref, "configs/recognition/vit/vit_imagenet_*.py"))):
        m = aimb200.load_config(f)["model"]
        out[os.path.relpath(f, ref)] = {"backbone": m["backbone"], "cls_head": m["cls_head"], "test_cfg": m.get("test_cfg")}
    return out


if __name__ == "__main__":
    di = collect_imagenet()
    pi = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vit_imagenet_configs.json")
    json.dump(di, open(pi, "w"), indent=1, sort_keys=True)
    print(f"{len(di)} configs -> {pi}")
    d = collect()
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vitclip_configs.json")
    json.dump(d, open(p, "w"), indent=1, sort_keys=True)
    print(f"{len(d)} configs -> {p}")
