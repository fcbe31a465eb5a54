"""``sys.modules`` stubs so the reference's OWN python source imports unchanged here.

Oracle / test infrastructure only.  Works only where ``/root/reference`` exists (the build
container); the GPU box never calls this.  Nothing from the reference is copied: the stubs
only supply the third-party names its modules import at module load:

=====================  ==========================================================
stub                   needed by (reference file:line)
=====================  ==========================================================
mobile_cv...fbnet_v2   model_training/model/blocks.py:5   -> oracle.fbnet_c.fbnet
hydra / omegaconf      model_training/utils/hydra.py:5-8, dataset/tracking_dataset.py:8
coloredlogs            model_training/utils/logger.py:4-13,27-32
albumentations         tracker/base_tracker.py:4,69-81, utils/utils.py:6,234-252,
                       dataset/aug.py:6-49,52  (Compose / Normalize / Resize are real)
got10k.datasets        dataset/__init__.py:3
pytorch_toolbelt       dataset/siam_dataset.py:6, utils/torch.py:6
=====================  ==========================================================

albumentations==1.0.0 semantics restated from its published source (package absent, so
this is from the library's documented behaviour, not verifiable offline):
``Normalize``: ``(img.astype(f32) - mean*255) * (1/(std*255))`` in float32;
``Resize``: ``cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)`` (no-op when the size
already matches) and coco boxes scaled proportionally; boxes of zero area are filtered.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("FEAR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model_training", "model", "fear_net.py"))


# --------------------------------------------------------------------------- albumentations
class _Transform:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class Normalize(_Transform):
    def __init__(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), max_pixel_value=255.0, **kw):
        super().__init__()
        self.mean, self.std, self.max_pixel_value = mean, std, max_pixel_value

    def apply(self, img):
        mean = np.array(self.mean, dtype=np.float32)
        mean *= self.max_pixel_value
        std = np.array(self.std, dtype=np.float32)
        std *= self.max_pixel_value
        denominator = np.reciprocal(std, dtype=np.float32)
        img = img.astype(np.float32)
        img -= mean
        img *= denominator
        return img


class Resize(_Transform):
    def __init__(self, height, width, interpolation=None, **kw):
        super().__init__()
        self.height, self.width = height, width

    def apply(self, img):
        import cv2

        if img.shape[0] == self.height and img.shape[1] == self.width:
            return img
        return cv2.resize(img, dsize=(self.width, self.height), interpolation=cv2.INTER_LINEAR)


class Compose:
    def __init__(self, transforms, bbox_params=None, **kw):
        self.transforms = list(transforms)
        self.bbox_params = bbox_params

    def __call__(self, **data):
        img = data["image"]
        rows, cols = img.shape[:2]
        for t in self.transforms:
            img = t.apply(img)
        out = dict(data)
        out["image"] = img
        if "bboxes" in data:
            assert self.bbox_params is None or self.bbox_params.get("format", "coco") == "coco"
            new_rows, new_cols = img.shape[:2]
            boxes = []
            for box in data["bboxes"]:
                x, y, w, h = [float(v) for v in box[:4]]
                # coco -> normalised albumentations format
                nb = [x / cols, y / rows, (x + w) / cols, (y + h) / rows]
                for v in nb:
                    if not 0.0 <= v <= 1.0:
                        raise ValueError(f"Expected bbox in [0,1], got {nb}")
                area = (nb[2] - nb[0]) * cols * (nb[3] - nb[1]) * rows
                if not area:  # filter_bboxes drops zero-area boxes (min_area=0, min_visibility=0)
                    continue
                x0, y0, x1, y1 = nb[0] * new_cols, nb[1] * new_rows, nb[2] * new_cols, nb[3] * new_rows
                boxes.append((x0, y0, x1 - x0, y1 - y0) + tuple(box[4:]))
            out["bboxes"] = boxes
        return out


def _dummy_class(name):
    return type(name, (_Transform,), {})


def _make_albumentations():
    m = types.ModuleType("albumentations")
    m.Compose, m.Normalize, m.Resize = Compose, Normalize, Resize
    m.DualTransform = _dummy_class("DualTransform")
    m.to_tuple = lambda v, *a, **k: tuple(v) if isinstance(v, (tuple, list)) else (-v, v)
    m.__getattr__ = lambda name: _dummy_class(name)  # PEP 562: OneOf, Blur, ... (aug.py:8-49)
    return m


# --------------------------------------------------------------------------- everything else
def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _not_available(*a, **k):
    raise RuntimeError("hydra/omegaconf are stubs in the oracle environment")


def install() -> None:
    """Install the stubs and put the reference on sys.path (idempotent)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from oracle import fbnet_c

    stubs = {
        "mobile_cv": _module("mobile_cv"),
        "mobile_cv.model_zoo": _module("mobile_cv.model_zoo"),
        "mobile_cv.model_zoo.models": _module("mobile_cv.model_zoo.models"),
        "mobile_cv.model_zoo.models.fbnet_v2": _module("mobile_cv.model_zoo.models.fbnet_v2", fbnet=fbnet_c.fbnet),
        "hydra": _module("hydra", main=lambda *a, **k: (lambda f: f)),
        "hydra.utils": _module("hydra.utils", get_original_cwd=os.getcwd, instantiate=_not_available),
        "hydra.initialize": _module("hydra.initialize", initialize=_not_available),
        "hydra.compose": _module("hydra.compose", compose=_not_available),
        "omegaconf": _module("omegaconf", OmegaConf=type("OmegaConf", (), {}), DictConfig=dict),
        "coloredlogs": _module("coloredlogs", DEFAULT_FIELD_STYLES={}, install=lambda **kw: None),
        "albumentations": _make_albumentations(),
        "got10k": _module("got10k"),
        "got10k.datasets": _module(
            "got10k.datasets", VOT=type("VOT", (), {}), GOT10k=type("GOT10k", (), {}), NfS=type("NfS", (), {})
        ),
        "pytorch_toolbelt": _module("pytorch_toolbelt"),
        "pytorch_toolbelt.utils": _module(
            "pytorch_toolbelt.utils", image_to_tensor=_not_available, transfer_weights=_not_available
        ),
    }
    for name, mod in stubs.items():
        sys.modules.setdefault(name, mod)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


REF_MODEL_KWARGS = dict(  # model_training/config/model/fear.yaml:2-17
    backbone="custom_fbnet", growth_factor=1.2, upsample="pixel_shuffle", pretrained=True, num_classes=1,
    num_filters=32, num_channels=3, align=False, img_size=256, stride=2, conv_block="sep_conv", towernum=2,
    mobile=True, max_layer=4, crop_template_features=False,
)
REF_TRACKER_KWARGS = dict(  # model_training/config/tracker/siam_tracker.yaml:2-15 (stride = ${model.stride})
    penalty_k=0.062, window_influence=0.38, lr=0.765, windowing="cosine", total_stride=16, score_size=16,
    ratio=0.94, stride=2, bbox_ratio=0.5, template_bbox_offset=0.2, search_context=2, instance_size=256,
    template_size=128,
)
REF_CKPT = os.path.join(REFERENCE_ROOT, "evaluate", "checkpoints", "FEAR-XS-NoEmbs.ckpt")
REF_VIDEO = os.path.join(REFERENCE_ROOT, "assets", "test.mp4")
REF_INIT_BBOX = [163, 53, 45, 174]  # demo_video.py:45


def build_reference_net():
    """The reference's own FEARNet, strict-loaded from its shipped checkpoint, eval mode, CPU."""
    install()
    from model_training.model.fear_net import FEARNet  # reference source
    from model_training.utils.torch import load_from_lighting  # reference source

    net = FEARNet(**REF_MODEL_KWARGS)
    net = load_from_lighting(net, REF_CKPT, map_location="cpu").eval()
    return net


def build_reference_tracker(net):
    install()
    from model_training.tracker.fear_tracker import FEARTracker  # reference source

    return FEARTracker(model=net, cuda_id="cpu", **REF_TRACKER_KWARGS)
