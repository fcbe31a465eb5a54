"""CPU: BN folding / packing (feartracker_b200.weights) is exact -- a plain-torch pipeline fed with
the FOLDED tensors (in the layouts the library receives) reproduces the fp64 oracle."""
import numpy as np
import torch
import torch.nn.functional as F

from feartracker_b200 import _lib, build, weights
from oracle import fear_oracle as fo
from oracle.fbnet_c import FBNET_C, NUM_HOT_BLOCKS
from tests.helpers import golden

R, C = fo.TARGET_REGRESSION_LABEL_KEY, fo.TARGET_CLASSIFICATION_KEY


def _t(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float64))


def _pw(x, w, b, relu):
    y = F.conv2d(x, _t(w).reshape(w.shape[0], -1, 1, 1), _t(b))
    return F.relu(y) if relu else y


def _dw(x, w, b, stride, relu):
    w = _t(w).reshape(w.shape[0], 1, w.shape[-2], w.shape[-1])
    y = F.conv2d(x, w, None if b is None else _t(b), stride, w.shape[-1] // 2, 1, w.shape[0])
    return F.relu(y) if relu else y


def folded_features(fw, x):
    y = F.relu(F.conv2d(x, _t(fw["stem.w"]), _t(fw["stem.b"]), 2, 1))
    for spec in FBNET_C[1:NUM_HOT_BLOCKS]:
        if spec.kind != "ir":
            continue
        n, inp = spec.name, y
        if spec.expand != 1:
            y = _pw(y, fw[n + ".pw.w"], fw[n + ".pw.b"], True)
        y = _dw(y, fw[n + ".dw.w"], fw[n + ".dw.b"], spec.stride, True)
        y = _pw(y, fw[n + ".pwl.w"], fw[n + ".pwl.b"], False)
        if spec.residual:
            y = y + inp
    return _pw(y, fw["neck.w"], fw["neck.b"], False)


def folded_head(fw, zf, xf):
    outs = {}
    for br in ("cls", "reg"):
        x = _pw(_dw(xf, fw[f"{br}_encode.dw.w"], None, 1, False), fw[f"{br}_encode.pw.w"], fw[f"{br}_encode.pw.b"], True)
        cat = fo.pixelwise_correlation(zf.reshape(zf.size(0), zf.size(1), -1), x)
        outs[br] = _pw(_dw(cat, fw[f"{br}_dw.dw.w"], None, 1, False), fw[f"{br}_dw.pw.w"], fw[f"{br}_dw.pw.b"], True)
    res = {}
    for tw, br, pred, key in (("bbox_tower", "reg", "bbox_pred", R), ("cls_tower", "cls", "cls_pred", C)):
        x = outs[br]
        for i in range(2):
            x = _pw(_dw(x, fw[f"{tw}.{i}.dw.w"], None, 1, False), fw[f"{tw}.{i}.pw.w"], fw[f"{tw}.{i}.pw.b"], True)
        x = _pw(_dw(x, fw[f"{pred}.dw.w"], None, 1, False), fw[f"{pred}.pw.w"], fw[f"{pred}.pw.b"], False)
        res[key] = torch.exp(x) if pred == "bbox_pred" else x
    return res


def test_folded_pipeline_matches_oracle_fp64(state_dict):
    fw = weights.fold_state_dict(state_dict)
    sd64 = fo.to_dtype(state_dict, torch.float64)
    zt, xt, _, _ = fo.synthetic_crops(4)
    zt, xt = zt[:2], xt[:2]
    zf = folded_features(fw, zt.double())
    xf = folded_features(fw, xt.double())
    np.testing.assert_allclose(zf.numpy(), fo.get_features(sd64, zt.double()).numpy(), rtol=1e-9, atol=1e-10)
    mine = folded_head(fw, zf, xf)
    ref = fo.forward(sd64, zt.double(), xt.double())
    np.testing.assert_allclose(mine[R].numpy(), ref[R].numpy(), rtol=1e-9)
    np.testing.assert_allclose(mine[C].numpy(), ref[C].numpy(), rtol=1e-8, atol=1e-10)
    g = golden("synthetic_b4.npz")
    np.testing.assert_allclose(mine[R].numpy(), g["reg64"][:2], rtol=1e-8)


def test_pack_follows_library_table(state_dict):
    build.build()
    table = _lib.weight_table()
    blob, offsets = weights.pack(state_dict, table)
    assert blob.dtype == np.float32 and offsets.dtype == np.uint64
    assert int(offsets[-1]) == blob.size == sum(n for _, n in table)
    fw = weights.fold_state_dict(state_dict)
    i = [n for n, _ in table].index("xif4_7.pwl.w")
    np.testing.assert_array_equal(blob[int(offsets[i]):int(offsets[i + 1])],
                                  fw["xif4_7.pwl.w"].astype(np.float32).reshape(-1))
    assert np.abs(blob).max() < 1e5 and np.isfinite(blob).all()
