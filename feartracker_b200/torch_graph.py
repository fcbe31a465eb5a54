"""Differentiable torch graph of FEARNet for ``train()`` mode.

The accelerated product is the eval-mode inference path in libfear_b200.  Training needs what that path does not
have -- BatchNorm batch statistics + running-stat updates and autograd through every layer -- so in ``train()`` mode
``FEARNet.forward`` / ``get_features`` / ``connector`` / ``track`` run this plain PyTorch graph over the SAME
``nn.Parameter``s the library packs (NOT accelerated: cuDNN / eager kernels on whatever device the tensors live on).
That is what lets the reference's training step -- ``FEARLightningModel.forward`` ->
``self.model.forward(inputs)`` (reference model_training/train/fear_lightning_model.py:60-62) followed by
``FEARLoss`` and ``loss.backward()`` -- run on this FEARNet unchanged; gradients land on the parameters, and the next
``eval()`` call re-folds and re-packs them for the library.

Structure restated from the reference: mobile_cv fbnet_c IRF blocks ``[pw] -> dw -> pwl (+x)`` with ReLU after pw / dw
(see oracle/fbnet_c.py header for how that architecture is pinned), ``Encoder`` stage slicing blocks.py:27-35,
``AdjustLayer`` blocks.py:75-88, ``SepConv`` blocks.py:45-72, ``MatrixMobile`` blocks.py:91-105,
``MobileCorrelation`` blocks.py:108-126, ``BoxTower.forward`` blocks.py:174-194, ``FEARNet`` fear_net.py:58-96.
"""
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .constants import TARGET_CLASSIFICATION_KEY, TARGET_REGRESSION_LABEL_KEY

NUM_HOT_STAGES = 18  # fear_net.py:59 with max_layer = 4: Encoder.stages[:4] = fbnet_c stages 0..17


def _conv_bn(m, x: torch.Tensor, relu: bool) -> torch.Tensor:
    y = m.bn(m.conv(x))  # real nn.Conv2d / nn.BatchNorm2d modules: batch statistics in train(), running in eval()
    return F.relu(y) if relu else y


def _irf(m, x: torch.Tensor) -> torch.Tensor:
    y = _conv_bn(m.pw, x, True) if hasattr(m, "pw") else x
    y = _conv_bn(m.dw, y, True)
    y = _conv_bn(m.pwl, y, False)
    residual = m.dw.conv.stride[0] == 1 and m.pwl.conv.out_channels == x.shape[1]
    return y + x if residual else y


def feature_extractor(net, x: torch.Tensor) -> torch.Tensor:
    """FEARNet.feature_extractor (fear_net.py:58-61): fbnet_c stages 0..17."""
    stages = net.encoder.model.backbone.stages
    for i, (_, stage) in enumerate(stages.named_children()):
        if i >= NUM_HOT_STAGES:
            break
        if isinstance(stage, torch.nn.Identity):
            continue
        x = _irf(stage, x) if hasattr(stage, "dw") else _conv_bn(stage, x, True)
    return x


def get_features(net, crop: torch.Tensor) -> torch.Tensor:
    """FEARNet.get_features (fear_net.py:63-66): backbone + AdjustLayer (1x1 conv, BN)."""
    return net.neck.downsample(feature_extractor(net, crop))


def _sep(m, x: torch.Tensor) -> torch.Tensor:
    return m.pointwise(m.depthwise(x))  # SepConv.forward, blocks.py:69-72


def _sep_bn_relu(seq, x: torch.Tensor, start: int = 0) -> torch.Tensor:
    return F.relu(seq[start + 1](_sep(seq[start], x)))


def _correlate(enc_seq, z: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """MobileCorrelation.forward (blocks.py:119-126); z (B|1, C, 64) broadcasts over the search batch."""
    b, c, w, h = x.size()
    s = torch.matmul(z.permute(0, 2, 1), x.view(b, c, -1)).view(b, -1, w, h)
    return _sep_bn_relu(enc_seq, torch.cat([x, s], dim=1))


def box_tower(net, search: torch.Tensor, kernel: torch.Tensor, update: Optional[torch.Tensor] = None
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """BoxTower.forward (blocks.py:174-194) -> (bbox, cls, cls_dw, x_reg)."""
    cm = net.connect_model
    zc = update if update is not None else kernel  # cls_encode(update, search) when a dynamic template is given
    cls_z, cls_x = zc.reshape(zc.size(0), zc.size(1), -1), _sep_bn_relu(cm.cls_encode.matrix11_s, search)
    reg_z, reg_x = kernel.reshape(kernel.size(0), kernel.size(1), -1), _sep_bn_relu(cm.reg_encode.matrix11_s, search)
    cls_dw = _correlate(cm.cls_dw.enc, cls_z, cls_x)
    reg_dw = _correlate(cm.reg_dw.enc, reg_z, reg_x)
    x_reg = reg_dw
    for i in range(0, len(cm.bbox_tower), 3):
        x_reg = _sep_bn_relu(cm.bbox_tower, x_reg, i)
    x = torch.exp(cm.adjust * _sep(cm.bbox_pred, x_reg) + cm.bias)
    c = cls_dw
    for i in range(0, len(cm.cls_tower), 3):
        c = _sep_bn_relu(cm.cls_tower, c, i)
    cls = 0.1 * _sep(cm.cls_pred, c)
    return x, cls, cls_dw, x_reg


def connector(net, template_features: torch.Tensor, search_features: torch.Tensor) -> Dict[str, torch.Tensor]:
    bbox, cls, _, _ = box_tower(net, search_features, template_features)
    return {TARGET_REGRESSION_LABEL_KEY: bbox, TARGET_CLASSIFICATION_KEY: cls}


def forward(net, x: Tuple[torch.Tensor, torch.Tensor]) -> Dict[str, torch.Tensor]:
    template, search = x
    net.size = search.size(0)
    return connector(net, get_features(net, template), get_features(net, search))


def track(net, search: torch.Tensor, template_features: torch.Tensor) -> Dict[str, torch.Tensor]:
    return connector(net, template_features, get_features(net, search))
