"""Restatement of mobile_cv's FBNet-V2 ``fbnet_c`` (oracle; test infrastructure only).

The reference imports ``mobile_cv.model_zoo.models.fbnet_v2.fbnet`` (reference
``model_training/model/blocks.py:5,22-25``) from facebookresearch/mobile-vision pinned at
commit 51804a6873ae1029257cf652179c960cceeecc75 (``requirements.txt:8``).  That package is
neither vendored in the reference nor installed here, so its published architecture is
restated below.  What pins it:

* the shipped checkpoint ``evaluate/checkpoints/FEAR-XS-NoEmbs.ckpt`` -- parameter names
  and shapes fix, per block, which of ``pw`` / ``dw`` / ``pwl`` exist, kernel sizes,
  channel counts, ``bias=True`` on every conv and BN after every conv
  (``load_state_dict(strict=True)`` must succeed, reference ``utils/torch.py:20-21``);
* the traced op graph inside the reference's CoreML models:
  ReLU after stem / pw / dw, none after pwl, residual add exactly on blocks with
  stride 1 and Cin == Cout, padding k//2.

Unverifiable residue ("parity unpinned"): BN eps (torch default 1e-5 assumed).
"""
from collections import OrderedDict
from typing import List, NamedTuple, Optional

import torch
import torch.nn as nn


class BlockSpec(NamedTuple):
    name: str
    kind: str  # "conv" (conv-bn-relu), "ir" (inverted residual), "skip" (identity)
    cin: int
    cout: int
    k: int
    stride: int
    expand: int  # expansion ratio (mid = cin * expand); 1 => no pw conv

    @property
    def mid(self) -> int:
        return self.cin * self.expand

    @property
    def residual(self) -> bool:
        return self.kind == "ir" and self.stride == 1 and self.cin == self.cout


# 24 entries of ``fbnet_c.backbone.stages`` in order.
FBNET_C: List[BlockSpec] = [
    BlockSpec("xif0_0", "conv", 3, 16, 3, 2, 1),
    BlockSpec("xif1_0", "ir", 16, 16, 3, 1, 1),
    BlockSpec("xif2_0", "ir", 16, 24, 3, 2, 6),
    BlockSpec("xif2_1", "skip", 24, 24, 0, 1, 1),
    BlockSpec("xif2_2", "ir", 24, 24, 3, 1, 1),
    BlockSpec("xif2_3", "ir", 24, 24, 3, 1, 1),
    BlockSpec("xif3_0", "ir", 24, 32, 5, 2, 6),
    BlockSpec("xif3_1", "ir", 32, 32, 5, 1, 3),
    BlockSpec("xif3_2", "ir", 32, 32, 5, 1, 6),
    BlockSpec("xif3_3", "ir", 32, 32, 3, 1, 6),
    BlockSpec("xif4_0", "ir", 32, 64, 5, 2, 6),
    BlockSpec("xif4_1", "ir", 64, 64, 5, 1, 3),
    BlockSpec("xif4_2", "ir", 64, 64, 5, 1, 6),
    BlockSpec("xif4_3", "ir", 64, 64, 5, 1, 6),
    BlockSpec("xif4_4", "ir", 64, 112, 5, 1, 6),
    BlockSpec("xif4_5", "ir", 112, 112, 5, 1, 6),
    BlockSpec("xif4_6", "ir", 112, 112, 5, 1, 6),
    BlockSpec("xif4_7", "ir", 112, 112, 5, 1, 3),
    # constructed + loaded, never executed by FEAR (max_layer=4 runs stages[:18])
    BlockSpec("xif5_0", "ir", 112, 184, 5, 2, 6),
    BlockSpec("xif5_1", "ir", 184, 184, 5, 1, 6),
    BlockSpec("xif5_2", "ir", 184, 184, 5, 1, 6),
    BlockSpec("xif5_3", "ir", 184, 184, 5, 1, 6),
    BlockSpec("xif5_4", "ir", 184, 352, 3, 1, 6),
    BlockSpec("xif6_0", "conv", 352, 1984, 1, 1, 1),
]

NUM_HOT_BLOCKS = 18  # blocks.py:27-35 slices [0:2],[2:5],[5:9],[9:18]; fear_net.py:59 uses stages[:4]
BN_EPS = 1e-5


class ConvBNRelu(nn.Module):
    """conv(bias=True, pad=k//2) -> BN -> optional ReLU, with sub-module names
    ``conv`` / ``bn`` / ``relu`` as the checkpoint expects."""

    def __init__(self, cin: int, cout: int, k: int, stride: int, groups: int = 1, relu: bool = True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, groups=groups, bias=True)
        self.bn = nn.BatchNorm2d(cout, eps=BN_EPS)
        if relu:
            self.relu = nn.ReLU(inplace=True)
        self._has_relu = relu

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.bn(self.conv(x))
        return self.relu(x) if self._has_relu else x


class IRFBlock(nn.Module):
    """Inverted-residual block: [pw 1x1 + BN + ReLU] -> dw kxk + BN + ReLU -> pwl 1x1 + BN (+ x)."""

    def __init__(self, spec: BlockSpec):
        super().__init__()
        mid = spec.mid
        if spec.expand != 1:
            self.pw = ConvBNRelu(spec.cin, mid, 1, 1)
        self.dw = ConvBNRelu(mid, mid, spec.k, spec.stride, groups=mid)
        self.pwl = ConvBNRelu(mid, spec.cout, 1, 1, relu=False)
        self._has_pw = spec.expand != 1
        self._residual = spec.residual

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.pw(x) if self._has_pw else x
        y = self.pwl(self.dw(y))
        return y + x if self._residual else y


def build_stage(spec: BlockSpec) -> nn.Module:
    if spec.kind == "conv":
        return ConvBNRelu(spec.cin, spec.cout, spec.k, spec.stride)
    if spec.kind == "skip":
        return nn.Identity()
    return IRFBlock(spec)


class _Backbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.stages = nn.Sequential(OrderedDict((s.name, build_stage(s)) for s in FBNET_C))


class _ClsHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(1984, 1000, 1)


class FBNetC(nn.Module):
    """Object with the attribute surface the reference's ``Encoder`` touches:
    ``.backbone.stages`` (sliceable nn.Sequential of 24 named blocks) and ``.head``."""

    def __init__(self):
        super().__init__()
        self.backbone = _Backbone()
        self.head = _ClsHead()


def fbnet(name: str, pretrained: Optional[bool] = False) -> FBNetC:
    """Stand-in for ``mobile_cv.model_zoo.models.fbnet_v2.fbnet`` (no download: there is
    no network; the FEAR checkpoint overwrites every weight anyway)."""
    assert name == "fbnet_c", name
    return FBNetC()
