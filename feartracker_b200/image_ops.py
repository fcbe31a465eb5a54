"""Host-side crop / box helpers of the tracking loop (the CPU half of FEARTracker).

Behavioural mirror of the helpers the reference tracker calls (reference
model_training/utils/utils.py:29-71,202-253 and base_tracker.py:69-103): integer context
boxes, constant-colour padding, cv2 bilinear resize, ImageNet normalisation in float32.
``albumentations`` is not required: its Resize is ``cv2.resize(INTER_LINEAR)`` and its Normalize
is ``(img - mean*255) * (1/(std*255))`` in float32.
"""
from typing import Optional, Sequence, Tuple

import cv2
import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_MEAN255 = np.array(IMAGENET_MEAN, dtype=np.float32) * np.float32(255.0)
_INV_STD255 = np.reciprocal(np.array(IMAGENET_STD, dtype=np.float32) * np.float32(255.0), dtype=np.float32)


def context_box(bbox: Sequence[float], offset: float) -> np.ndarray:
    """Grow [x, y, w, h] by ``offset`` * side on every side, truncated to int32."""
    x, y, w, h = bbox
    return np.array([x - w * offset, y - h * offset, w * (1.0 + 2.0 * offset), h * (1.0 + 2.0 * offset)]).astype(
        "int32")


def trim_box(bbox: Sequence[float], img_shape: Sequence[int]) -> np.ndarray:
    """Clip [x, y, w, h] to an image of shape (h, w, ...)."""
    x1, y1, w, h = bbox
    x1, y1 = min(max(0, x1), img_shape[1]), min(max(0, y1), img_shape[0])
    x2, y2 = min(max(0, x1 + w), img_shape[1]), min(max(0, y1 + h), img_shape[0])
    return np.array([x1, y1, x2 - x1, y2 - y1]).astype("int32")


def clamp_bbox(bbox: Sequence[float], shape: Sequence[int], min_side: int = 3) -> np.ndarray:
    x, y, w, h = trim_box(bbox, shape)
    img_h, img_w = shape[0], shape[1]
    if w < min_side:
        w = min_side
        x -= max(0, x + w - img_w)
    if h < min_side:
        h = min_side
        y -= max(0, y + h - img_h)
    return np.array([x, y, w, h])


def extended_crop(image: np.ndarray, bbox: Sequence[float], crop_size: int, offset: float,
                  padding_value: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Square-resized context crop around ``bbox``.

    Returns (crop uint8 (crop_size, crop_size, 3), bbox inside the crop [x,y,w,h] float,
    context box in frame coordinates int32 [x,y,w,h])."""
    if padding_value is None:
        padding_value = np.mean(image, axis=(0, 1))
    ctx = context_box(bbox, offset)
    img_h, img_w = image.shape[:2]
    left, top = max(-ctx[0], 0), max(-ctx[1], 0)
    right, bottom = max(ctx[0] + ctx[2] - img_w, 0), max(ctx[1] + ctx[3] - img_h, 0)
    inner = image[ctx[1] + top: ctx[1] + ctx[3] - bottom, ctx[0] + left: ctx[0] + ctx[2] - right]
    padded = cv2.copyMakeBorder(inner, top, bottom, left, right, cv2.BORDER_CONSTANT, value=padding_value)
    rows, cols = padded.shape[:2]
    box = trim_box([bbox[0] - ctx[0], bbox[1] - ctx[1], bbox[2], bbox[3]], (rows, cols))
    if box[2] * box[3] == 0:
        raise IndexError("target box has zero area inside its context crop")
    crop = padded if (rows, cols) == (crop_size, crop_size) else cv2.resize(
        padded, dsize=(crop_size, crop_size), interpolation=cv2.INTER_LINEAR)
    x0, y0 = float(box[0]) / cols * crop_size, float(box[1]) / rows * crop_size
    x1, y1 = float(box[0] + box[2]) / cols * crop_size, float(box[1] + box[3]) / rows * crop_size
    return crop, np.array([x0, y0, x1 - x0, y1 - y0]), ctx


def normalize(image: np.ndarray) -> np.ndarray:
    """uint8 HWC -> float32 HWC, ImageNet statistics, same float32 operation order as the reference."""
    out = image.astype(np.float32)
    out -= _MEAN255
    out *= _INV_STD255
    return out


def rescale_bbox(bbox: Sequence[float], context: Sequence[float], instance_size: int) -> list:
    """Map a box from the 256x256 search crop back to frame pixels (python round, sides >= 3)."""
    sx, sy = context[2] / instance_size, context[3] / instance_size
    out = [round(bbox[0] * sx + context[0]), round(bbox[1] * sy + context[1]),
           max(3, round(bbox[2] * sx)), max(3, round(bbox[3] * sy))]
    return [int(v) for v in out]


# ---------------------------------------------------------------------------------------------- device crop
RESIZE_COEF_SCALE = np.float32(2048.0)  # cv::INTER_RESIZE_COEF_BITS = 11


def _axis_table(src: int, dst: int, clamp: bool):
    """Source offset + two fixed-point coefficients per destination index, as cv::resize (INTER_LINEAR, 8-bit)
    computes them: float32 position (dst + 0.5) * scale - 0.5, floor, fraction; along x the offset is clamped into
    the row and the fraction zeroed, along y only the ROW INDEX is clamped later (the kernel does that)."""
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * (float(src) / float(dst)) - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        low, high = s < 0, s >= src - 1
        f = np.where(low | high, np.float32(0.0), f)
        s = np.where(low, 0, np.where(high, src - 1, s)).astype(np.int32)
    a0 = np.rint((np.float32(1.0) - f) * RESIZE_COEF_SCALE).astype(np.int32)
    a1 = np.rint(f * RESIZE_COEF_SCALE).astype(np.int32)
    return s, a0, a1


def resize_tables(src_w: int, src_h: int, dst: int) -> np.ndarray:
    """(6 * dst,) int32: xofs, xa0, xa1, yofs, ya0, ya1 -- the tail of fear_crop_resize_u8's parameter block."""
    return np.concatenate(_axis_table(src_w, dst, True) + _axis_table(src_h, dst, False)).astype(np.int32)


def crop_params(bbox: Sequence[float], crop_size: int, offset: float, padding_value: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Parameter block of fear_crop_resize_u8 for the context crop around ``bbox`` plus the two values
    ``extended_crop`` returns besides the image: the box inside the crop and the context box."""
    ctx = context_box(bbox, offset)
    cols, rows = int(ctx[2]), int(ctx[3])
    box = trim_box([bbox[0] - ctx[0], bbox[1] - ctx[1], bbox[2], bbox[3]], (rows, cols))
    if box[2] * box[3] == 0:
        raise IndexError("target box has zero area inside its context crop")
    pad = np.clip(np.rint(np.asarray(padding_value, dtype=np.float64)), 0, 255).astype(np.int32)  # cv::saturate_cast
    head = np.array([ctx[0], ctx[1], cols, rows, pad[0], pad[1], pad[2], 0], dtype=np.int32)
    params = np.concatenate([head, resize_tables(cols, rows, crop_size)])
    x0, y0 = float(box[0]) / cols * crop_size, float(box[1]) / rows * crop_size
    x1, y1 = float(box[0] + box[2]) / cols * crop_size, float(box[1] + box[3]) / rows * crop_size
    return params, np.array([x0, y0, x1 - x0, y1 - y0]), ctx


def crop_resize_reference(frame: np.ndarray, params: np.ndarray, crop_size: int) -> np.ndarray:
    """numpy model of crop_resize_u8_kernel (same integer arithmetic); used by the CPU tests to pin the kernel's
    formula to cv2.resize without a GPU."""
    cx, cy, cw, ch = (int(v) for v in params[:4])
    pad = params[4:7].astype(np.int64)
    t = params[8:].reshape(6, crop_size).astype(np.int64)
    xo, a0, a1, yo, b0, b1 = t
    h, w = frame.shape[:2]

    def px(ys, xs):
        fy, fx = cy + ys[:, None], cx + xs[None, :]
        inside = (fy >= 0) & (fy < h) & (fx >= 0) & (fx < w)
        vals = frame[np.clip(fy, 0, h - 1), np.clip(fx, 0, w - 1)].astype(np.int64)
        return np.where(inside[..., None], vals, pad[None, None, :])

    x1 = np.minimum(xo + 1, cw - 1)
    y0, y1 = np.clip(yo, 0, ch - 1), np.clip(yo + 1, 0, ch - 1)
    s0 = px(y0, xo) * a0[None, :, None] + px(y0, x1) * a1[None, :, None]
    s1 = px(y1, xo) * a0[None, :, None] + px(y1, x1) * a1[None, :, None]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
