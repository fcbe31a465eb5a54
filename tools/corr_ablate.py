"""Profiling build only (FEAR_NVCC_FLAGS=-DFEAR_CORR_ABLATE): time corr_ts_kernel with pipeline roles switched off, to see
which role bounds the kernel.  mask bits: 1 = no MMAs, 2 = no convert, 4 = no epilogue, 8 = x tiles from L2 (no HBM stream)."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feartracker_b200 import _lib  # noqa: E402

lib = _lib.init(0)
lib.fear_debug_corr_ablate.argtypes = [ctypes.c_int]
B = 512
zt = torch.randn(B, 64, 256, device="cuda")
cats = [torch.randn(B, 256, 320, device="cuda") for _ in range(3)]
st = torch.cuda.current_stream().cuda_stream
out = {}
for form in ("ts",):
    for mask in [0, 1, 2, 4, 6, 7, 8, 14]:
        lib.fear_debug_corr_ablate(mask)
        for c in cats:
            lib.fear_corr_nhwc_f32(zt.data_ptr(), B, c.data_ptr(), B, st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(30):
            lib.fear_corr_nhwc_f32(zt.data_ptr(), B, cats[i % 3].data_ptr(), B, st)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / 30
        out[f"mask{mask}"] = round(us, 2)
        print(f"{form} mask={mask:2d}: {us:7.2f} us  ({393216 * B / us * 1e-3:6.0f} GB/s algorithmic)", flush=True)
lib.fear_debug_corr_ablate(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/corr_ablate.json", "w"))
