"""CPU: host-side logic of the product (API surface, crops, sharding) against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

import feartracker_b200 as fb
from feartracker_b200 import image_ops, sharding
from oracle import fear_oracle as fo
from tests.helpers import GOLDEN, load_full_state


def test_fearnet_state_dict_is_checkpoint_compatible():
    net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS)
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        keys = json.load(f)
    sd = net.state_dict()
    assert set(sd) == set(keys) and len(sd) == 520
    for k, (shape, dtype) in keys.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == "torch." + dtype, k
    net.load_state_dict(load_full_state(), strict=True)
    assert net.search_size == 256 and net.max_layer == 4 and net.grid_x.dtype == torch.float64
    assert net.encoder.encoder_channels["layer1"] == 112 and len(net.encoder.stages) == 5


def test_fearnet_constructor_contract():
    with pytest.raises(AssertionError):
        fb.FEARNet(backbone="x", img_size=256, max_layer=5)
    with pytest.raises(NotImplementedError):
        fb.FEARNet(backbone="x", img_size=256)  # reference defaults (towernum=4, max_layer=3) are not FEAR-XS
    fb.FEARNet(backbone="custom_fbnet", img_size=256, towernum=2, max_layer=4, growth_factor=1.2, num_filters=32)


def test_no_cpu_and_no_training_path():
    net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        net((torch.zeros(1, 3, 128, 128), torch.zeros(1, 3, 256, 256)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.get_features(torch.zeros(1, 3, 128, 128))
    net.train()
    with pytest.raises(NotImplementedError):
        net.track(torch.zeros(1, 3, 256, 256), torch.zeros(1, 256, 8, 8))
    with pytest.raises(RuntimeError):
        net.encoder(torch.zeros(1, 3, 32, 32))  # parameter containers never compute


def test_crop_helpers_match_oracle():
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (256, 480, 3), dtype=np.uint8)
    mean = np.mean(frame, axis=(0, 1))
    boxes = [[163, 53, 45, 174], [0, 0, 30, 40], [450, 230, 30, 26], [200, 100, 3, 3], [-5, -7, 50, 60],
             [10, 200, 400, 56]]
    for box in boxes:
        box = image_ops.clamp_bbox(box, frame.shape)
        assert list(box) == list(fo.clamp_bbox(box, frame.shape))
        for size, off, pad in ((128, 0.2, None), (256, 2, mean)):
            a = image_ops.extended_crop(frame, box, size, off, pad)
            b = fo.get_extended_crop(frame, box, size, off, pad)
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_allclose(a[1], b[1], rtol=0, atol=1e-12)
            np.testing.assert_array_equal(a[2], b[2])
    crop = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
    np.testing.assert_array_equal(image_ops.normalize(crop), fo.normalize_image(crop))
    for pred in ([106.2, 88.3, 44.6, 49.1], [0.4999, 255.5, 1.0, 2.5]):
        ctx = [100, 20, 225, 225]
        assert image_ops.rescale_bbox(np.array(pred), ctx, 256) == fo.rescale_bbox(np.array(pred), ctx, 256)


def test_tracker_constructs_without_gpu_and_fails_loudly():
    net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS).eval()
    trk = fb.FEARTracker(net, cuda_id="cpu", **fb.FEAR_XS_TRACKER_KWARGS)
    assert trk.window.shape == (16, 16) and trk.box_coder.grid_x.shape == (1, 16, 16)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            trk.initialize(np.zeros((64, 64, 3), np.uint8), np.array([10, 10, 20, 20]))


def test_shard_range_partitions():
    for total, world in ((2048, 8), (10, 4), (3, 8), (256, 1)):
        spans = [sharding.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
