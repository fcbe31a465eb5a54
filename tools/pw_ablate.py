"""Profiling build only (FEAR_NVCC_FLAGS=-DFEAR_PW_ABLATE): the bench step with roles of pw_tc_kernel switched off, per-stage
device time.  mask bits: 1 = no MMAs, 2 = no split / depthwise compute, 4 = no epilogue, 8 = no weight loads, 16 = no
activation loads.  Results are garbage; only the timing matters."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feartracker_b200 as fb  # noqa: E402
from bench import load_state  # noqa: E402
from feartracker_b200 import _lib  # noqa: E402
from oracle import fear_oracle as fo  # noqa: E402

B = 256
net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS)
net.load_state_dict(load_state(), strict=True)
net = net.cuda().eval()
net.reserve(B)
lib = _lib.load()
lib.fear_debug_pw_ablate.argtypes = [ctypes.c_int]
zt, xt, _, _ = fo.synthetic_crops(B)
x = xt.cuda()
zf = net.get_features(zt.cuda())
masks = [int(m) for m in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 3, 24, 31]
for mask in masks:
    lib.fear_debug_pw_ablate(mask)
    for _ in range(3):
        net.track_boxes(x, zf)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        net.track_boxes(x, zf)
    b.record()
    torch.cuda.synchronize()
    total = a.elapsed_time(b) / 10
    net.profile(True)
    for _ in range(5):
        net.track_boxes(x, zf)
    torch.cuda.synchronize()
    st = net.stage_times()
    net.profile(False)
    print(f"mask {mask:2d}: step {total:6.3f} ms | " + "  ".join(f"{k} {v[0] / 5:.3f}" for k, v in st.items() if v[0] / 5 > 0.03), flush=True)
lib.fear_debug_pw_ablate(0)
