"""BN folding and packing of a FEARNet state_dict into the blob libfear_b200 consumes.

All folding is done in float64 and cast to float32 once (several hot-path BN layers have
running_var ~ 0, giving folded scales up to ~3.6e3 -- SURVEY.md section 7.4).  Tensor names and
order come from the library itself (``fear_weight_name``), so the two sides cannot drift.

Folds (reference module -> library tensor):
  conv(bias) -> BN                      w' = w*s ; b' = (b - mean)*s + beta      s = gamma/sqrt(var+eps)
  SepConv(dw bias bd, pw bias bp) -> BN pw' = W*s ; b' = (W.bd + bp - mean)*s + beta ; dw keeps no bias
     (a depthwise bias is spatially uniform, so pushing it through the 1x1 is exact)
  bbox_pred + exp(adjust*x + bias)      pw' = adjust*W ; b' = adjust*(W.bd + bp) + bias   (blocks.py:187)
  cls_pred, 0.1*x                       pw' = 0.1*W   ; b' = 0.1*(W.bd + bp)              (blocks.py:192)
"""
from typing import Dict, Mapping

import numpy as np
import torch

BN_EPS = 1e-5
_BACKBONE = "encoder.model.backbone.stages."


def _f64(t: torch.Tensor) -> np.ndarray:
    return t.detach().to("cpu", torch.float64).numpy()


def _bn_scale_shift(sd: Mapping[str, torch.Tensor], prefix: str):
    g, b = _f64(sd[prefix + ".weight"]), _f64(sd[prefix + ".bias"])
    m, v = _f64(sd[prefix + ".running_mean"]), _f64(sd[prefix + ".running_var"])
    s = g / np.sqrt(v + BN_EPS)
    return s, b - m * s  # y = x*s + shift


def _fold_conv_bn(sd, conv: str, bn: str):
    w = _f64(sd[conv + ".weight"])
    bias = _f64(sd[conv + ".bias"]) if conv + ".bias" in sd else np.zeros(w.shape[0])
    s, shift = _bn_scale_shift(sd, bn)
    return w * s.reshape(-1, 1, 1, 1), bias * s + shift


def _fold_sep_bn(sd, sep: str, bn: str):
    """SepConv(depthwise, pointwise) followed by BN -> (dw weights, folded pw weights, folded bias)."""
    dw = _f64(sd[sep + ".depthwise.weight"])
    W = _f64(sd[sep + ".pointwise.weight"])[:, :, 0, 0]
    bias = np.zeros(W.shape[0])
    if sep + ".pointwise.bias" in sd:
        bias = bias + _f64(sd[sep + ".pointwise.bias"])
    if sep + ".depthwise.bias" in sd:
        bias = bias + W @ _f64(sd[sep + ".depthwise.bias"])
    if bn is not None:
        s, shift = _bn_scale_shift(sd, bn)
        W = W * s[:, None]
        bias = bias * s + shift
    return dw, W, bias


def fold_state_dict(sd: Mapping[str, torch.Tensor], wanted_blocks=None) -> Dict[str, np.ndarray]:
    """state_dict (reference key names, ``model.`` prefix already stripped) -> folded fp64 tensors
    keyed by the library's tensor names.  ``wanted_blocks`` limits the backbone blocks folded
    (the checkpoint also carries the never-executed xif5_*/xif6_0 tail)."""
    out: Dict[str, np.ndarray] = {}
    w, b = _fold_conv_bn(sd, _BACKBONE + "xif0_0.conv", _BACKBONE + "xif0_0.bn")
    out["stem.w"], out["stem.b"] = w, b
    blocks = sorted({k[len(_BACKBONE):].split(".")[0] for k in sd if k.startswith(_BACKBONE)})
    for name in blocks:
        if wanted_blocks is not None and name not in wanted_blocks:
            continue
        p = _BACKBONE + name
        for part in ("pw", "dw", "pwl"):
            if f"{p}.{part}.conv.weight" in sd:
                w, b = _fold_conv_bn(sd, f"{p}.{part}.conv", f"{p}.{part}.bn")
                out[f"{name}.{part}.w"], out[f"{name}.{part}.b"] = w, b
    w, b = _fold_conv_bn(sd, "neck.downsample.0", "neck.downsample.1")
    out["neck.w"], out["neck.b"] = w, b
    cm = "connect_model."
    for br in ("cls", "reg"):
        dw, W, bias = _fold_sep_bn(sd, f"{cm}{br}_encode.matrix11_s.0", f"{cm}{br}_encode.matrix11_s.1")
        out[f"{br}_encode.dw.w"], out[f"{br}_encode.pw.w"], out[f"{br}_encode.pw.b"] = dw, W, bias
        dw, W, bias = _fold_sep_bn(sd, f"{cm}{br}_dw.enc.0", f"{cm}{br}_dw.enc.1")
        out[f"{br}_dw.dw.w"], out[f"{br}_dw.pw.w"], out[f"{br}_dw.pw.b"] = dw, W, bias
    for tw in ("bbox_tower", "cls_tower"):
        for i, seq in enumerate((0, 3)):
            dw, W, bias = _fold_sep_bn(sd, f"{cm}{tw}.{seq}", f"{cm}{tw}.{seq + 1}")
            out[f"{tw}.{i}.dw.w"], out[f"{tw}.{i}.pw.w"], out[f"{tw}.{i}.pw.b"] = dw, W, bias
    adjust = float(_f64(sd[cm + "adjust"]).reshape(-1)[0])
    bias4 = _f64(sd[cm + "bias"]).reshape(-1)
    dw, W, bias = _fold_sep_bn(sd, cm + "bbox_pred", None)
    out["bbox_pred.dw.w"], out["bbox_pred.pw.w"], out["bbox_pred.pw.b"] = dw, adjust * W, adjust * bias + bias4
    dw, W, bias = _fold_sep_bn(sd, cm + "cls_pred", None)
    out["cls_pred.dw.w"], out["cls_pred.pw.w"], out["cls_pred.pw.b"] = dw, 0.1 * W, 0.1 * bias
    return out


def pack(sd: Mapping[str, torch.Tensor], table) -> (np.ndarray, np.ndarray):
    """Fold and lay tensors out in the library's order.  Returns (blob float32, offsets uint64[n+1])."""
    folded = fold_state_dict(sd, wanted_blocks={name.split(".")[0] for name, _ in table})
    offsets = np.zeros(len(table) + 1, dtype=np.uint64)
    chunks = []
    for i, (name, numel) in enumerate(table):
        if name not in folded:
            raise KeyError(f"state_dict lacks the tensors for library weight '{name}'")
        a = np.ascontiguousarray(folded[name], dtype=np.float64).reshape(-1)
        if a.size != numel:
            raise ValueError(f"weight '{name}': library expects {numel} elements, state_dict gives {a.size}")
        if not np.all(np.isfinite(a)):
            raise ValueError(f"weight '{name}' is not finite after BN folding")
        chunks.append(a.astype(np.float32))
        offsets[i + 1] = offsets[i] + np.uint64(numel)
    return np.concatenate(chunks), offsets
